"""Tiled 2D U-Net prediction on B200 (reference: unet/predict.py:14-229; same constructor surface)."""
from typing import Union

import numpy as np
import torch

from .. import pipeline2d as P
from .. import tiff
from ..dist import DistContext
from ..engine import Engine
from ..progress import ProgressNotifier
from ..utils import get_device, save_as_tif
from .unet import Unet


class Session:
    """Reusable 2D predictor: checkpoint folded/packed once, weights and workspace resident on the device.

    ``predict(frames)`` takes a host (F, H, W) uint8/uint16 stack (ideally pinned) and returns the stitched
    (F, C, H, W) uint8 result on the host; ``predict_device`` does the same for a stack already in HBM.
    ``Predict`` (the reference's constructor-runs-everything class) is a thin wrapper around this.
    """

    def __init__(self, model_params, resize_dim=(512, 512), invert=False, normalization_mode='single',
                 clip_threshold=(0., 99.8), add_tile=0, device='cuda:0', precision='tf32', workspace_gb=24.0):
        if normalization_mode not in ('single', 'first', 'all'):
            raise ValueError(f'normalization_mode {normalization_mode} not valid!')
        params = torch.load(model_params, map_location='cpu') if isinstance(model_params, str) else model_params
        self.device = torch.device(device)
        self.resize_dim, self.add_tile, self.invert = tuple(resize_dim), add_tile, invert
        self.normalization_mode, self.clip_threshold = normalization_mode, clip_threshold
        self.out_channels = params['out_channels']
        self.workspace_bytes = int(workspace_gb * 2 ** 30)
        self.engine = Engine('unet2d', params['state_dict'], params['n_filter'], params['in_channels'],
                             [('', self.out_channels, 'sigmoid')], precision=precision, device=self.device)
        self.tile_batch = None
        self.fixed_lut = None          # set for 'first' / 'all' (stack-wide statistics)
        self.last = {}

    def _ensure_plan(self, total_tiles):
        if self.tile_batch is None or (total_tiles < self.tile_batch):
            self.tile_batch = P.pick_tile_batch(self.engine, self.resize_dim, max(1, total_tiles), self.workspace_bytes)

    def normalise_device(self, frames_dev):
        if self.fixed_lut is not None:
            return P.E.apply_lut(frames_dev, self.fixed_lut)
        if self.normalization_mode != 'single':
            hist = P.E.histogram(frames_dev)
            total = P.E.hist_sum(hist)
            bounds = total if self.normalization_mode == 'all' else hist[0:1].contiguous()
            lut, _ = P.E.norm_lut(bounds, total, 1, self.clip_threshold[0], self.clip_threshold[1], self.invert)
            return P.E.apply_lut(frames_dev, lut)
        return P.Normalizer2D('single', self.clip_threshold, self.invert)(frames_dev)

    def predict_device(self, frames_dev, keep=False):
        """(F, H, W) uint8/uint16 device tensor -> (F, C, H, W) uint8 device tensor."""
        f, h, w = frames_dev.shape
        n_x, n_y, _, _ = P.tiling.grid_2d(h, w, self.resize_dim, self.add_tile)
        self._ensure_plan(f * n_x * n_y)
        norm = self.normalise_device(frames_dev)
        out, grid, tiles, res_tiles = P.predict_frames_2d(self.engine, norm, self.resize_dim, self.add_tile,
                                                          self.out_channels, self.tile_batch)
        self.last = dict(grid=grid, norm=norm, tiles=tiles if keep else None, result_tiles=res_tiles if keep else None)
        return out

    def predict(self, frames):
        """Host stack in, host result out (H2D and D2H inside)."""
        if isinstance(frames, np.ndarray):
            frames_dev = P.to_device_stack(frames, self.device)
        else:
            frames_dev = frames.to(self.device, non_blocking=True)
        out = self.predict_device(frames_dev)
        return out.cpu().numpy()

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None


class Predict:
    """Prediction of movies and images with U-Net.

    1) load + intensity normalisation, 2) split into tiles of `resize_dim`, 3) U-Net forward, 4) stitch (mean of
    overlapping regions), 5) write a float16 TIFF — exactly the reference's steps, run as CUDA kernels.

    Parameters (identical to the reference, unet/predict.py:54-57)
    ----------
    imgs : ndarray or str      images to predict; a string is read as a TIFF file
    result_name : str          path of the result TIFF
    model_params : str         path of the checkpoint (.pt) written by the reference's Trainer
    network                    'Unet' (string or class); None reads model_params['network']
    resize_dim                 tile size (multiples of 16)
    invert, normalization_mode ('single' | 'first' | 'all'), clip_threshold, add_tile, normalize_result,
    show_progress, device, progress_notifier : as in the reference

    Engine-only keyword arguments (defaults reproduce the reference's behaviour)
    ----------
    precision : 'tf32' (default, sigmoid within 1e-3 of the fp32 reference) | 'bf16' (within 1e-2) | 'fp32'
    workspace_gb : activation workspace budget used to pick the tile batch size
    keep_intermediates : keep the uint8 tiles / result tiles as attributes (test hook)
    mutate_input : the reference overwrites the caller's array with the normalised frames in 'single' mode
        (unet/predict.py:131); kept by default
    distributed : shard frames over the ranks of an initialised torch.distributed process group (one GPU per rank)
    """

    def __init__(self, imgs, result_name, model_params, network='Unet', resize_dim=(512, 512),
                 invert=False, normalization_mode='single', clip_threshold=(0., 99.8), add_tile=0,
                 normalize_result=False, show_progress=True, device: Union[torch.device, str] = 'auto',
                 progress_notifier: ProgressNotifier = ProgressNotifier.progress_notifier_tqdm(), *,
                 precision='tf32', workspace_gb=24.0, mutate_input=True, distributed=False,
                 keep_intermediates=False):
        self.dist = DistContext(distributed)
        self._keep = {'patches': [], 'result_patches': []} if keep_intermediates else None
        if device == 'auto':
            self.device = self.dist.device() if self.dist.active else get_device()
        else:
            self.device = torch.device(device)

        if isinstance(imgs, str):
            imgs = tiff.imread(imgs)

        self.resize_dim = resize_dim
        self.add_tile = add_tile
        self.normalize_result = normalize_result
        self.invert = invert
        self.normalization_mode = normalization_mode
        self.clip_threshold = clip_threshold
        self.result_name = result_name
        self.show_progress = show_progress
        if normalization_mode not in ('single', 'first', 'all'):
            raise ValueError(f'normalization_mode {normalization_mode} not valid!')

        imgs = self.__reshape_data(imgs)

        # checkpoint -> engine
        self.model_params = torch.load(model_params, map_location='cpu')
        if network is None:
            if 'network' in self.model_params.keys():
                network = self.model_params['network']
            else:
                raise ValueError('network is not defined')
        name = network if isinstance(network, str) else getattr(network, '__name__', str(network))
        if name in ('AttentionUnet', 'Unet_v0'):
            raise NotImplementedError(f"network '{name}' is not implemented by the B200 engine yet (Unet is)")
        if name != 'Unet':
            raise ValueError(f"unknown network '{name}'")
        out_channels = self.model_params['out_channels']
        if self.model_params['in_channels'] != 1:
            # the reference's tile array has a single channel (unet/predict.py:158) and its .view() fails otherwise
            raise RuntimeError("shape '[1, %d, %d, %d]' is invalid for input of size %d" % (
                self.model_params['in_channels'], resize_dim[0], resize_dim[1], resize_dim[0] * resize_dim[1]))
        self.session = Session(self.model_params, resize_dim, invert, normalization_mode, clip_threshold, add_tile,
                               self.device, precision, workspace_gb)

        # frames of this rank
        t_total = self.imgs_shape[0]
        lo, hi = self.dist.shard(t_total)
        self.N_x, self.N_y, self.X_start, self.Y_start = P.tiling.grid_2d(self.imgs_shape[1], self.imgs_shape[2],
                                                                          resize_dim, add_tile)
        self.N_per_img = self.N_x * self.N_y
        self.N = self.N_per_img * t_total
        print('Predicting data ...') if self.show_progress and self.dist.rank == 0 else None

        result_local = self.__run(imgs, lo, hi, out_channels, workspace_gb, mutate_input, progress_notifier)
        self.session.close()
        del self.session

        imgs_result = self.dist.gather_frames(result_local, t_total, self.device)
        if imgs_result is not None:
            imgs_result = np.squeeze(imgs_result)
            save_as_tif(imgs_result, self.result_name, normalize=normalize_result)
        del self.model_params
        torch.cuda.empty_cache()

    def __reshape_data(self, imgs):
        self.imgs_shape = imgs.shape
        if len(self.imgs_shape) == 2:  # single image
            imgs = np.expand_dims(imgs, axis=0)
            self.imgs_shape = imgs.shape
        return imgs

    def __run(self, imgs, lo, hi, out_channels, workspace_gb, mutate_input, progress_notifier):
        th, tw = self.resize_dim
        h, w = self.imgs_shape[1:]
        n_local = hi - lo
        ses = self.session
        ses._ensure_plan(max(1, n_local * self.N_per_img))
        # frames per chunk: enough tiles to fill a few batches, bounded so the uint8 tile arrays stay small
        chunk = max(1, min(max(n_local, 1), max(1, (4 * ses.tile_batch) // self.N_per_img)))
        if self.normalization_mode in ('first', 'all'):
            ses.fixed_lut = self.__global_lut(imgs, lo, hi, chunk)
        out = np.zeros((n_local, out_channels, h, w), dtype='uint8')
        starts = range(lo, hi, chunk)
        it = progress_notifier.iterator(starts) if (self.show_progress and self.dist.rank == 0) else starts
        for s in it:
            e = min(s + chunk, hi)
            frames = P.to_device_stack(imgs[s:e], self.device)
            res = ses.predict_device(frames, keep=self._keep is not None)
            if ses.fixed_lut is None and mutate_input:
                imgs[s:e] = ses.last['norm'].cpu().numpy()      # unet/predict.py:131 (cast back to the input dtype)
            out[s - lo:e - lo] = res.cpu().numpy()
            if self._keep is not None:
                self._keep['patches'].append(ses.last['tiles'].cpu().numpy())
                self._keep['result_patches'].append(ses.last['result_tiles'].cpu().numpy())
        if self._keep is not None:   # test hook: what the reference's __split / __predict return
            self.patches = np.concatenate(self._keep['patches'])
            self.result_patches = np.concatenate(self._keep['result_patches'])
        return out

    def __global_lut(self, imgs, lo, hi, chunk):
        """'first' / 'all': bounds from frame 0 / the whole stack, min/max from the whole stack
        (unet/predict.py:132-147). Histograms are summed over chunks and over ranks."""
        total = None
        for s in range(lo, hi, chunk):
            frames = P.to_device_stack(imgs[s:min(s + chunk, hi)], self.device)
            part = P.E.hist_sum(P.E.histogram(frames))
            total = part if total is None else total + part
        if total is None:
            total = torch.zeros((1, P.E.HIST_BINS), dtype=torch.int32, device=self.device)
        total = self.dist.all_reduce_sum(total)
        if self.normalization_mode == 'all':
            bounds = total
        else:
            bounds = P.E.histogram(P.to_device_stack(imgs[0:1], self.device))
        lut, _ = P.E.norm_lut(bounds, total, 1, self.clip_threshold[0], self.clip_threshold[1], self.invert)
        return lut
