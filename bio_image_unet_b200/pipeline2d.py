"""Device pipeline shared by the 2D predictors: normalise -> split -> batched forward -> stitch.

Reference: unet/predict.py:115-229 (and the per-pair variant siam_unet/predict.py:125-240). All heavy steps run
as CUDA kernels through the C-ABI; this module only sequences them and does the (tiny) index arithmetic.
"""
import numpy as np
import torch

from . import engine as E
from . import tiling


def to_device_stack(frames_np, device):
    """(F, H, W) uint8 / uint16 / float32 host array (numpy, or a torch tensor - ideally pinned) -> device tensor."""
    if torch.is_tensor(frames_np):
        if frames_np.dtype not in (torch.uint8, torch.uint16, torch.float32):
            raise TypeError(f'bio_image_unet_b200 normalises uint8 / uint16 / float32 stacks on the device; got {frames_np.dtype}')
        return frames_np.contiguous().to(device, non_blocking=True)
    if frames_np.dtype not in (np.uint8, np.uint16, np.float32):
        raise TypeError(f'bio_image_unet_b200 normalises uint8 / uint16 / float32 stacks on the device; got '
                        f'{frames_np.dtype}. Convert the stack (e.g. to uint16 or float32) before calling Predict.')
    host = torch.from_numpy(np.ascontiguousarray(frames_np))
    try:
        host = host.pin_memory()
    except RuntimeError:
        pass
    return host.to(device, non_blocking=True)


class PinnedOut:
    """Reusable pinned host buffer for the final D2H copy of a result (grown on demand)."""

    def __init__(self):
        self.buf = None

    def view(self, shape, dtype):
        """Pinned host tensor of the given shape (the buffer is reused by the next call)."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        return self.buf[:n].view(dtype).view(*shape)

    def fetch(self, t):
        n = t.numel() * t.element_size()
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        host = self.buf[:n].view(t.dtype).view(t.shape)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()


class Normalizer2D:
    """Percentile normalisation of a (F, H, W) integer stack to uint8 on the device
    (unet/predict.py:122-150). ``stats_reduce`` lets the multi-GPU driver all-reduce the stack-wide histogram."""

    def __init__(self, mode, clip_threshold, invert, stats_reduce=None, first_frame_hist=None):
        if mode not in ('single', 'first', 'all'):
            raise ValueError(f'normalization_mode {mode} not valid!')
        self.mode, self.clip, self.invert = mode, clip_threshold, invert
        self.stats_reduce = stats_reduce
        self.first_frame_hist = first_frame_hist
        self.params = None

    def __call__(self, frames_dev):
        f = frames_dev.shape[0]
        hist = E.histogram(frames_dev)
        if self.mode == 'single':
            lut, self.params = E.norm_lut(hist, hist, f, self.clip[0], self.clip[1], self.invert)
        else:
            total = E.hist_sum(hist)
            if self.stats_reduce is not None:
                total = self.stats_reduce(total)
            if self.mode == 'all':
                bounds = total
            else:
                bounds = self.first_frame_hist(hist) if self.first_frame_hist is not None else hist[0:1].contiguous()
            lut, self.params = E.norm_lut(bounds, total, 1, self.clip[0], self.clip[1], self.invert)
        return E.apply_lut(frames_dev, lut)


def run_tiles(eng, tiles, tile_batch, prev_tiles=None, want_val=False):
    """Forward all tiles in batches of the planned size (the tail batch is zero-padded)."""
    n = tiles.shape[0]
    outs_u8, outs_val = [], []
    for s in range(0, n, tile_batch):
        t = tiles[s:s + tile_batch]
        p = None if prev_tiles is None else prev_tiles[s:s + tile_batch]
        cnt = t.shape[0]
        if cnt < tile_batch:
            pad = torch.zeros((tile_batch - cnt, *t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat((t, pad))
            if p is not None:
                p = torch.cat((p, pad))
        val, u8 = eng.forward(t.contiguous(), None if p is None else p.contiguous(), want_val=want_val, want_u8=True)
        outs_u8.append(u8[:cnt])
        if want_val:
            outs_val.append(val[:cnt])
    u8 = outs_u8[0] if len(outs_u8) == 1 else torch.cat(outs_u8)
    val = None if not want_val else (outs_val[0] if len(outs_val) == 1 else torch.cat(outs_val))
    return u8, val


def even_batch(total_tiles, budget_batch):
    """Tile batch for a job of `total_tiles`: as few forwards as the workspace budget allows, all of (almost) the same
    size - 288 tiles with room for 160 run as 2 x 144, not 160 + 128 padded to 160."""
    total_tiles, budget_batch = max(1, int(total_tiles)), max(1, int(budget_batch))
    n_fwd = -(-total_tiles // budget_batch)
    return -(-total_tiles // n_fwd)


class BatchPlanner:
    """Plan policy shared by the Session classes: the plan grows on demand, shrinks only when a job is at most half
    the planned batch (padding a 25-tile image to a 200-tile batch would cost 8x), and is never re-made for the
    shorter tail chunk of a movie (run_tiles pads that one)."""

    def __init__(self, engine, workspace_bytes):
        self.engine, self.workspace_bytes = engine, workspace_bytes
        self.tile, self.budget_batch, self.tile_batch = None, None, None
        self.plan_kwargs = {}                      # e.g. siam_shared=tiles per frame (Engine.plan)

    def budget(self, tile):
        """Largest tile batch the workspace budget holds for this tile size."""
        tile = tuple(int(v) for v in tile)
        key = (tile, tuple(sorted(self.plan_kwargs.items())))
        if self.tile != key:
            per_tile = self.engine.plan(1, tile)
            self.budget_batch = int(max(1, self.workspace_bytes // max(per_tile, 1)))
            self.tile, self.tile_batch = key, None
        return self.budget_batch

    def ensure(self, tile, total_tiles):
        tile = tuple(int(v) for v in tile)
        self.budget(tile)
        target = even_batch(total_tiles, self.budget_batch)
        if self.tile_batch is None or target > self.tile_batch or 2 * target <= self.tile_batch:
            self.engine.plan(target, tile, **self.plan_kwargs)
            self.tile_batch = target
        return self.tile_batch


def run_tiles_shared(eng, tiles_u, tile_batch, n_per):
    """Siam_UNet with the shared twin encoder: tiles_u = [tiles of the previous frame of the first pair | tiles of the
    current frames] (n_pairs + n_per tiles); pair j = (previous tile j, current tile j + n_per). Batches of `tile_batch`
    pairs take the contiguous slice of tile_batch + n_per unique tiles they need (the tail batch is zero-padded)."""
    n_pairs = tiles_u.shape[0] - n_per
    outs = []
    for s in range(0, n_pairs, tile_batch):
        cnt = min(tile_batch, n_pairs - s)
        t = tiles_u[s:s + cnt + n_per]
        if cnt < tile_batch:
            t = torch.cat((t, torch.zeros((tile_batch - cnt, *t.shape[1:]), dtype=t.dtype, device=t.device)))
        _, u8 = eng.forward(t.contiguous(), None, want_val=False, want_u8=True)
        outs.append(u8[:cnt])
    return outs[0] if len(outs) == 1 else torch.cat(outs)


def pick_tile_batch(eng, tile, total_tiles, budget_bytes):
    per_tile = eng.plan(1, tile)
    batch = int(max(1, min(total_tiles, budget_bytes // max(per_tile, 1))))
    eng.plan(batch, tile)
    return batch


def check_starts(starts, tile, extent):
    """The reference wraps negative uint16 starts when an undersized image is split into more than one tile and
    then fails on the slice assignment; report that as the same ValueError class up front."""
    for s in starts:
        if int(s) + tile > max(extent, tile):
            raise ValueError(f'could not broadcast tile: start {int(s)} + tile {tile} exceeds extent {extent} '
                             f'(image smaller than resize_dim with add_tile > 0)')


def predict_frames_2d(eng, norm_u8, resize_dim, add_tile, out_channels, tile_batch, pad_mode=0, raw=None, lut=None):
    """norm_u8: (F, H, W) uint8 device tensor (already normalised) - or None with `raw` (the uint8 / uint16 stack) and
    `lut` (its normalisation tables): the normalisation is then fused into the tile gather and the normalised stack
    is never written. Returns ((F, C, H, W) uint8 device tensor, grid, tiles, result_tiles)."""
    f, h, w = (norm_u8 if norm_u8 is not None else raw).shape
    th, tw = resize_dim
    n_x, n_y, xs, ys = tiling.grid_2d(h, w, resize_dim, add_tile)
    check_starts(xs, th, h)
    check_starts(ys, tw, w)
    if norm_u8 is not None:
        tiles = E.gather_tiles(norm_u8.view(f, 1, h, w), [0], xs, ys, (1, th, tw), pad_mode)  # (F*N, 1, th, tw)
    else:
        tiles = E.gather_tiles_lut(raw.view(f, 1, h, w), lut, [0], xs, ys, (1, th, tw), pad_mode)
    res_u8, _ = run_tiles(eng, tiles, tile_batch)
    out = E.stitch_mean_u8(res_u8, f, out_channels, (h, w), xs, ys, (th, tw))
    return out, (n_x, n_y, xs, ys), tiles, res_u8
