"""Patch-wise 3D U-Net prediction on B200 (reference: unet3d/predict.py:12-195)."""
from typing import Union

import numpy as np
import torch

from .. import pipeline2d as P
from .. import tiff
from .. import tiling
from ..dist import DistContext
from ..engine import Engine
from ..progress import ProgressNotifier
from ..utils import get_device, save_as_tif
from .unet3d import UNet3D


class Predict:
    """Prediction of movies or 3D stacks with 3D U-Net (constructor surface of unet3d/predict.py:52-55).

    `normalization_mode` is accepted and unused, as in the reference: percentiles are always taken over the whole
    volume (unet3d/predict.py:109-117). Engine-only keyword arguments as in unet.Predict; with `distributed=True`
    the patch list is sharded over ranks (contiguous ranges of the z -> x -> y ordered list), the intensity
    histogram is all-reduced, and rank 0 gathers the uint8 patches and stitches.
    """

    def __init__(self, vol, result_name, model_params, network=UNet3D, resize_dim=(512, 512),
                 invert=False, normalization_mode='single', clip_threshold=(0., 99.8), add_patch=0,
                 normalize_result=False, progress_bar=True, device: Union[torch.device, str] = 'auto',
                 progress_notifier: ProgressNotifier = ProgressNotifier.progress_notifier_tqdm(), *,
                 precision='tf32', workspace_gb=24.0, distributed=False, keep_intermediates=False):
        if isinstance(vol, str):
            vol = tiff.imread(vol)
        self.dist = DistContext(distributed)
        if device == 'auto':
            self.device = self.dist.device() if self.dist.active else get_device()
        else:
            self.device = torch.device(device)
        self.resize_dim = resize_dim
        self.add_patch = add_patch
        self.normalize_result = normalize_result
        self.invert = invert
        self.normalization_mode = normalization_mode
        self.clip_threshold = clip_threshold
        self.result_name = result_name
        self.progress_bar = progress_bar

        self.vol_shape = vol.shape
        if len(self.vol_shape) == 2:
            vol = np.expand_dims(vol, axis=0)
            self.vol_shape = vol.shape
        if len(resize_dim) != 3:
            raise IndexError('tuple index out of range')      # the reference indexes resize_dim[2] (:123)

        self.model_params = torch.load(model_params, map_location='cpu')
        use_interp = self.model_params.get('use_interpolation', False)
        self.engine = Engine('unet3d', self.model_params['state_dict'], self.model_params['n_filter'],
                             self.model_params['in_channels'], [('', self.model_params['out_channels'], 'sigmoid')],
                             use_interpolation=use_interp, precision=precision, device=self.device)

        (self.N_z, self.N_x, self.N_y, self.Z_start, self.X_start, self.Y_start) = tiling.grid_3d(
            self.vol_shape, resize_dim, add_patch)
        self.N = self.N_x * self.N_y * self.N_z
        print('Predicting data ...') if self.progress_bar and self.dist.rank == 0 else None

        vol_result = self.__run(vol, workspace_gb, keep_intermediates, progress_notifier)
        self.engine.close()
        del self.engine
        if vol_result is not None:
            save_as_tif(np.squeeze(vol_result), self.result_name, normalize=normalize_result)
        del self.model_params
        torch.cuda.empty_cache()

    def __run(self, vol, workspace_gb, keep, progress_notifier):
        d, h, w = (int(v) for v in self.resize_dim)
        z, x, y = self.vol_shape
        for starts, t, ext in ((self.Z_start, d, z), (self.X_start, h, x), (self.Y_start, w, y)):
            P.check_starts(starts, t, ext)
        dev = self.device
        # global percentile normalisation: every rank histograms its z-slab, one all-reduce, identical LUT everywhere
        vol_dev = P.to_device_stack(vol, dev)
        zlo, zhi = self.dist.shard(z)
        part = P.E.hist_sum(P.E.histogram(vol_dev[zlo:zhi])) if zhi > zlo else \
            torch.zeros((1, P.E.HIST_BINS), dtype=torch.int32, device=dev)
        total = self.dist.all_reduce_sum(part)
        lut, _ = P.E.norm_lut(total, total, 1, self.clip_threshold[0], self.clip_threshold[1], self.invert)
        norm = P.E.apply_lut(vol_dev, lut)                                       # (Z, X, Y) uint8
        del vol_dev
        patches = P.E.gather_tiles(norm.view(1, z, x, y), self.Z_start, self.X_start, self.Y_start, (d, h, w), 0)
        # this rank's contiguous share of the patch list
        lo, hi = self.dist.shard(self.N)
        mine = patches[lo:hi].reshape(hi - lo, 1, d, h, w)
        if hi > lo:
            tile_batch = P.pick_tile_batch(self.engine, (d, h, w), hi - lo, int(workspace_gb * 2 ** 30))
            res_local, _ = P.run_tiles(self.engine, mine, tile_batch)
            res_local = res_local.reshape(hi - lo, d, h, w)
        else:
            res_local = torch.zeros((0, d, h, w), dtype=torch.uint8, device=dev)
        res_all = self.dist.gather_frames(res_local.cpu().numpy(), self.N, dev)
        if keep:
            self.patches = patches.cpu().numpy()
            self.result_patches = res_all
        if res_all is None:
            return None
        st = P.E.stitch_mod3_u8(torch.from_numpy(res_all).to(dev), (z, x, y), self.Z_start, self.X_start, self.Y_start,
                                (d, h, w))
        return st.cpu().numpy()
