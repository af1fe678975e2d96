"""Host-side tile-grid arithmetic. Kept in numpy with the reference's exact expressions so that tile counts and
start indices are bit-identical (uint16 truncation of np.linspace included)."""
import numpy as np


def linspace_starts(extent, tile, count):
    """np.linspace(0, extent - tile, count).astype('uint16') — unet/predict.py:171-172."""
    return np.linspace(0, extent - tile, count).astype('uint16')


def grid_2d(height, width, resize_dim, add_tile):
    """(N_x, N_y, X_start, Y_start) of unet/predict.py:154-155,171-172 and siam_unet/predict.py:92-95,183-184.
    Note the reference's naming: 'x' runs along image rows (axis 1 of the stack), 'y' along columns."""
    n_x = int(np.ceil(height / resize_dim[0])) + add_tile
    n_y = int(np.ceil(width / resize_dim[1])) + add_tile
    return n_x, n_y, linspace_starts(height, resize_dim[0], n_x), linspace_starts(width, resize_dim[1], n_y)


def grid_3d(vol_shape, resize_dim, add_patch):
    """(N_z, N_x, N_y, Z_start, X_start, Y_start) of unet3d/predict.py:121-126,140-142 — including the
    as-written quirk that N_x is incremented when N_z > 1 and again when N_x > 1."""
    n_z = int(np.ceil(vol_shape[0] / resize_dim[0])) + add_patch
    n_x = int(np.ceil(vol_shape[1] / resize_dim[1])) + add_patch
    n_y = int(np.ceil(vol_shape[2] / resize_dim[2])) + add_patch
    if n_z > 1:
        n_x += add_patch
    if n_x > 1:
        n_x += add_patch
    if n_y > 1:
        n_y += add_patch
    return (n_z, n_x, n_y, linspace_starts(vol_shape[0], resize_dim[0], n_z),
            linspace_starts(vol_shape[1], resize_dim[1], n_x), linspace_starts(vol_shape[2], resize_dim[2], n_y))


def strided_starts(extent, patch, overlap_factor):
    """Patch starts along one axis of multi_output_unet3d/predict.py:134-147: uniform stride
    int(patch * (1 - overlap)), plus a final start flush with the far edge when the tail is uncovered."""
    stride = max(1, int(patch * (1 - overlap_factor)))
    starts = list(range(0, max(extent - patch + 1, 1), stride))
    if starts[-1] + patch < extent:
        starts.append(extent - patch)
    return starts


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [start, stop) slice of n_items for `rank` (frames / z-rows of patches / volumes)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
