"""Tiled Siamese U-Net prediction of TIFF movies on B200 (reference: siam_unet/predict.py:15-240).

The reference walks the movie frame by frame: pair = (previous, current) with frame 0 paired with frame 1,
normalise the pair, split both frames into tiles, forward each tile pair, stitch, append to the TIFF writer. Here
pairs are processed in chunks: every kernel runs once over all pairs of a chunk, and in 'single' mode each frame
is normalised once (its result is the same in both pairs it belongs to).
"""
import os
from typing import Union

import numpy as np
import torch

from .. import pipeline2d as P
from .. import tiff
from ..dist import DistContext
from ..engine import Engine
from ..progress import ProgressNotifier
from ..utils import get_device


def _take(frames, sel):
    """frames[sel] for uint8 / uint16 device stacks (torch has no CUDA index kernel for uint16)."""
    if frames.dtype == torch.uint16:
        return frames.view(torch.int16)[sel].contiguous().view(torch.uint16)
    return frames[sel].contiguous()


class Predict:
    """Prediction of tif-movies with Siamese U-Net (constructor surface of siam_unet/predict.py:53-56).

    tif_file : str (an ndarray (T, H, W) is also accepted); resize_dim=None processes whole frames.
    Engine-only keyword arguments: precision, workspace_gb, distributed, keep_intermediates (see unet.Predict).
    """

    def __init__(self, tif_file, result_name, model_params, resize_dim=(512, 512), invert=False,
                 normalization_mode='single', clip_threshold=(0.0, 99.98), add_tile=0, normalize_result=False,
                 show_progress=True, device: Union[torch.device, str] = 'auto',
                 progress_notifier: ProgressNotifier = ProgressNotifier.progress_notifier_tqdm(), *,
                 precision='tf32', workspace_gb=24.0, distributed=False, keep_intermediates=False):
        self.dist = DistContext(distributed)
        if device == 'auto':
            self.device = self.dist.device() if self.dist.active else get_device()
        else:
            self.device = torch.device(device)
        self.tif_file = tif_file
        self.add_tile = add_tile
        self.invert = invert
        self.normalization_mode = normalization_mode
        self.clip_threshold = clip_threshold
        self.result_name = result_name
        self.normalize_result = normalize_result
        self.show_progress = show_progress
        self.progress_notifier = progress_notifier
        if normalization_mode not in ('single', 'first', 'all'):
            raise ValueError(f'normalization_mode {normalization_mode} not valid!')

        # load model
        self.model_params = torch.load(model_params, map_location='cpu')
        self.engine = Engine('siam2d', self.model_params['state_dict'], self.model_params['n_filter'], 1,
                             [('', 1, 'sigmoid')], siam_mode=self.model_params['mode'], precision=precision,
                             device=self.device)

        if isinstance(tif_file, str):
            self.tif_len, page_shape = tiff.page_count_and_shape(tif_file)
            movie = tiff.imread(tif_file)
            movie = movie[None] if movie.ndim == 2 else movie
            os.makedirs(f'temp_{tif_file.split("/")[-1]}', exist_ok=True)    # siam_unet/predict.py:88,100
        else:
            movie = np.asarray(tif_file)
            movie = movie[None] if movie.ndim == 2 else movie
            self.tif_len, page_shape = movie.shape[0], movie.shape[1:]
        self.imgs_shape = [self.tif_len, page_shape[0], page_shape[1]]
        self.resize_dim = resize_dim if resize_dim is not None else (self.imgs_shape[1], self.imgs_shape[2])

        self.N_x, self.N_y, self.X_start, self.Y_start = P.tiling.grid_2d(self.imgs_shape[1], self.imgs_shape[2],
                                                                          self.resize_dim, add_tile)
        self.N_per_img = self.N_x * self.N_y
        self.N = self.N_x * self.N_y
        self._keep = {'patches': [], 'result_patches': []} if keep_intermediates else None

        print('Predicting data ...') if self.show_progress and self.dist.rank == 0 else None
        lo, hi = self.dist.shard(self.tif_len)
        local = self.__run(movie, lo, hi, workspace_gb)
        self.engine.close()
        del self.engine
        full = self.dist.gather_frames(local, self.tif_len, self.device)
        if full is not None:
            with tiff.TiffWriter(self.result_name, bigtiff=False) as tif:
                for frame in full:
                    tif.write(frame, contiguous=True)
        if self._keep is not None:
            self.patches = np.concatenate(self._keep['patches'])
            self.result_patches = np.concatenate(self._keep['result_patches'])
        del self.model_params
        torch.cuda.empty_cache()

    def __pair_indices(self, lo, hi):
        """(previous, current) frame index of every pair in [lo, hi) (siam_unet/predict.py:107-117)."""
        prev = [(1 if self.tif_len > 1 else 0) if i == 0 else i - 1 for i in range(lo, hi)]
        return prev, list(range(lo, hi))

    def __run(self, movie, lo, hi, workspace_gb):
        th, tw = self.resize_dim
        h, w = self.imgs_shape[1:]
        n_local = hi - lo
        out = np.zeros((n_local, h, w), dtype='uint8')
        if n_local == 0:
            return out
        tile_batch = P.pick_tile_batch(self.engine, (th, tw), n_local * self.N_per_img, int(workspace_gb * 2 ** 30))
        chunk = max(1, min(n_local, max(1, (4 * tile_batch) // self.N_per_img)))
        starts = range(lo, hi, chunk)
        it = self.progress_notifier.iterator(starts) if (self.show_progress and self.dist.rank == 0) else starts
        q_lo, q_hi = self.clip_threshold
        for s in it:
            e = min(s + chunk, hi)
            prev_idx, cur_idx = self.__pair_indices(s, e)
            needed = sorted(set(prev_idx + cur_idx))
            pos = {f: i for i, f in enumerate(needed)}
            frames = P.to_device_stack(np.stack([movie[f] for f in needed]), self.device)
            p_sel = torch.tensor([pos[f] for f in prev_idx], device=self.device)
            c_sel = torch.tensor([pos[f] for f in cur_idx], device=self.device)
            hist = P.E.histogram(frames)
            if self.normalization_mode == 'single':       # each frame on its own statistics (:128-136)
                lut, _ = P.E.norm_lut(hist, hist, len(needed), q_lo, q_hi, self.invert)
                norm = P.E.apply_lut(frames, lut)
                norm_prev, norm_cur = norm[p_sel].contiguous(), norm[c_sel].contiguous()
            else:                                          # statistics of the pair (:137-152): stack = [prev, cur]
                pair_range = (hist[p_sel] + hist[c_sel]).contiguous()
                bounds = hist[p_sel].contiguous() if self.normalization_mode == 'first' else pair_range
                lut, _ = P.E.norm_lut(bounds, pair_range, len(cur_idx), q_lo, q_hi, self.invert)
                norm_prev = P.E.apply_lut(_take(frames, p_sel), lut)
                norm_cur = P.E.apply_lut(_take(frames, c_sel), lut)
            n_pairs = e - s
            P.check_starts(self.X_start, th, h)
            P.check_starts(self.Y_start, tw, w)
            tiles_cur = P.E.gather_tiles(norm_cur.view(n_pairs, 1, h, w), [0], self.X_start, self.Y_start, (1, th, tw), 1)
            tiles_prev = P.E.gather_tiles(norm_prev.view(n_pairs, 1, h, w), [0], self.X_start, self.Y_start, (1, th, tw), 1)
            res_u8, _ = P.run_tiles(self.engine, tiles_cur, tile_batch, prev_tiles=tiles_prev)
            st = P.E.stitch_mean_u8(res_u8, n_pairs, 1, (h, w), self.X_start, self.Y_start, (th, tw))
            out[s - lo:e - lo] = st[:, 0].cpu().numpy()
            if self._keep is not None:
                both = torch.stack((tiles_cur[:, 0], tiles_prev[:, 0]), dim=1)       # ch0 = current, ch1 = previous
                self._keep['patches'].append(both.cpu().numpy().reshape(n_pairs, self.N, 2, th, tw))
                self._keep['result_patches'].append(res_u8.cpu().numpy().reshape(n_pairs, self.N, 1, th, tw))
        return out
