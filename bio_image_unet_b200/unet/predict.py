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

    KINDS = {'Unet': 'unet2d', 'AttentionUnet': 'attunet2d', 'Unet_v0': 'unet2d_v0'}

    def __init__(self, model_params, resize_dim=(512, 512), invert=False, normalization_mode='single',
                 clip_threshold=(0., 99.8), add_tile=0, device='cuda:0', precision='tf32', workspace_gb=24.0,
                 network='Unet'):
        if normalization_mode not in ('single', 'first', 'all'):
            raise ValueError(f'normalization_mode {normalization_mode} not valid!')
        params = torch.load(model_params, map_location='cpu') if isinstance(model_params, str) else model_params
        self.device = torch.device(device)
        self.resize_dim, self.add_tile, self.invert = tuple(resize_dim), add_tile, invert
        self.normalization_mode, self.clip_threshold = normalization_mode, clip_threshold
        self.out_channels = params['out_channels']
        self.workspace_bytes = int(workspace_gb * 2 ** 30)
        if network not in self.KINDS:
            raise ValueError(f"unknown network '{network}'")
        self.engine = Engine(self.KINDS[network], params['state_dict'], params['n_filter'], params['in_channels'],
                             [('', self.out_channels, 'sigmoid')], precision=precision, device=self.device)
        self.tile_batch = None
        self.fixed_lut = None          # set for 'first' / 'all' (stack-wide statistics)
        self.last = {}
        self._pin, self._streams, self._dev_in = {}, None, None

    def _ensure_plan(self, total_tiles):
        if self.tile_batch is None or (total_tiles < self.tile_batch):
            self.tile_batch = P.pick_tile_batch(self.engine, self.resize_dim, max(1, total_tiles), self.workspace_bytes)

    def normalise_device(self, frames_dev):
        if frames_dev.dtype == torch.float32:
            # float stacks: exact float32 percentiles by radix select; 'first' / 'all' need the whole stack in one call
            u8, f32, _ = P.E.normalize_f32(frames_dev.contiguous(), self.normalization_mode, self.clip_threshold[0],
                                           self.clip_threshold[1], self.invert, want_f32=True)
            self.last_norm_f32 = f32
            return u8
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
        if frames_dev.dtype == torch.float32:      # what the reference stores back into a float stack (:131)
            norm = self.last_norm_f32
        self.last = dict(grid=grid, norm=norm, tiles=tiles if keep else None, result_tiles=res_tiles if keep else None)
        return out

    def predict(self, frames):
        """Host stack in, host result out (H2D and D2H inside, pipelined against the compute)."""
        return self.predict_movie(frames)[0]

    def _pinned(self, key, shape, dtype):
        """Reusable pinned host buffer (grown on demand)."""
        n = int(np.prod(shape))
        buf = self._pin.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(n, dtype=dtype, pin_memory=True)
            self._pin[key] = buf
        return buf[:n].view(*shape)

    def predict_movie(self, frames, chunk_frames=None, want_norm=False):
        """Pipelined prediction of a host (F, H, W) uint8/uint16 stack (numpy array or torch tensor, ideally
        pinned): the frames go through the device in chunks, and the H2D copy of chunk i+1 and the D2H copy of
        chunk i-1 run on their own streams while chunk i computes. Returns (result (F, C, H, W) uint8 numpy
        array backed by a pinned buffer that the next call reuses, normalised frames (F, H, W) uint8 or None)."""
        if isinstance(frames, np.ndarray):
            if frames.dtype not in (np.uint8, np.uint16, np.float32):
                raise TypeError(f'bio_image_unet_b200 normalises uint8 / uint16 / float32 stacks on the device; got '
                                f'{frames.dtype}. Convert the stack (e.g. to uint16 or float32) before calling Predict.')
            host = torch.from_numpy(np.ascontiguousarray(frames))
        else:
            if frames.dtype not in (torch.uint8, torch.uint16, torch.float32):
                raise TypeError(f'bio_image_unet_b200 normalises uint8 / uint16 / float32 stacks on the device; got {frames.dtype}')
            host = frames.contiguous()
        f, h, w = host.shape
        n_x, n_y, _, _ = P.tiling.grid_2d(h, w, self.resize_dim, self.add_tile)
        self._ensure_plan(f * n_x * n_y)
        is_float = host.dtype == torch.float32
        if chunk_frames is None:
            chunk_frames = max(1, min(f, self.tile_batch // (n_x * n_y)))
        if is_float and self.normalization_mode != 'single':
            chunk_frames = f               # stack-wide float statistics are taken in one pass over the whole stack
        dev = self.device
        with torch.cuda.device(dev):
            if self._streams is None:
                self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_comp, s_out = self._streams
            cur = torch.cuda.current_stream(dev)
            for st in self._streams:
                st.wait_stream(cur)
            out_host = self._pinned('out', (f, self.out_channels, h, w), torch.uint8)
            norm_host = self._pinned('norm', (f, h, w), torch.float32 if is_float else torch.uint8) if want_norm else None
            pinned_in = host.is_pinned()
            key = (chunk_frames, h, w, host.dtype)
            if self._dev_in is None or self._dev_in[0] != key:
                self._dev_in = (key, [torch.empty((chunk_frames, h, w), dtype=host.dtype, device=dev) for _ in range(2)])
            dev_in = self._dev_in[1]
            stage = None if pinned_in else [self._pinned(f'stage{b}', (chunk_frames, h, w), host.dtype) for b in range(2)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]
            for i, s0 in enumerate(range(0, f, chunk_frames)):
                b = i & 1
                n = min(chunk_frames, f - s0)
                src = host[s0:s0 + n]
                if not pinned_in:
                    ev_in[b].synchronize()                      # the copy that last read this staging buffer is done
                    stage[b][:n].copy_(src)
                    src = stage[b][:n]
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_done[b])                 # the compute that last read dev_in[b] is done
                    dev_in[b][:n].copy_(src, non_blocking=True)
                    ev_in[b].record(s_in)
                with torch.cuda.stream(s_comp):
                    s_comp.wait_event(ev_in[b])
                    res = self.predict_device(dev_in[b][:n])
                    norm = self.last['norm'] if want_norm else None
                    ev_done[b].record(s_comp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    out_host[s0:s0 + n].copy_(res, non_blocking=True)
                    res.record_stream(s_out)
                    if want_norm:
                        norm_host[s0:s0 + n].copy_(norm, non_blocking=True)
                        norm.record_stream(s_out)
            s_out.synchronize()
            cur.wait_stream(s_comp)
        return out_host.numpy(), (norm_host.numpy() if want_norm else None)

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
        self._pin, self._dev_in = {}, None


class Predict:
    """Prediction of movies and images with U-Net.

    1) load + intensity normalisation, 2) split into tiles of `resize_dim`, 3) U-Net forward, 4) stitch (mean of
    overlapping regions), 5) write a float16 TIFF — exactly the reference's steps, run as CUDA kernels.

    Parameters (identical to the reference, unet/predict.py:54-57)
    ----------
    imgs : ndarray or str      images to predict; a string is read as a TIFF file
    result_name : str          path of the result TIFF
    model_params : str         path of the checkpoint (.pt) written by the reference's Trainer
    network                    'Unet' | 'AttentionUnet' | 'Unet_v0' (string or class); None reads model_params['network']
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
        # strings as in unet/predict.py:89-97, or a class (the reference's or this package's) identified by name
        name = network if isinstance(network, str) else getattr(network, '__name__', str(network))
        if name not in Session.KINDS:
            raise ValueError(f"unknown network '{name}'")
        if name == 'Unet_v0' and 'in_channels' not in self.model_params.keys():
            self.model_params['in_channels'] = 1          # old checkpoints, unet/predict.py:95-97
            self.model_params['out_channels'] = 1
        out_channels = self.model_params['out_channels']
        if self.model_params['in_channels'] != 1:
            # the reference's tile array has a single channel (unet/predict.py:158) and its .view() fails otherwise
            raise RuntimeError("shape '[1, %d, %d, %d]' is invalid for input of size %d" % (
                self.model_params['in_channels'], resize_dim[0], resize_dim[1], resize_dim[0] * resize_dim[1]))
        self.session = Session(self.model_params, resize_dim, invert, normalization_mode, clip_threshold, add_tile,
                               self.device, precision, workspace_gb, network=name)

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
        is_float = imgs.dtype == np.float32
        if is_float and self.normalization_mode in ('first', 'all'):
            if self.dist.active and self.dist.world > 1:
                raise NotImplementedError("float stacks with normalization_mode 'first' / 'all' are not sharded over ranks")
            chunk = max(n_local, 1)        # stack-wide float statistics: one pass over the whole stack on the device
        elif self.normalization_mode in ('first', 'all'):
            ses.fixed_lut = self.__global_lut(imgs, lo, hi, chunk)
        out = np.zeros((n_local, out_channels, h, w), dtype='uint8')
        # super-chunks bound the pinned host buffers; inside one, copies and compute are pipelined
        frames_per_call = max(chunk, min(max(n_local, 1), (1 << 30) // max(h * w * max(out_channels, 2), 1)))
        starts = range(lo, hi, frames_per_call)
        it = progress_notifier.iterator(starts) if (self.show_progress and self.dist.rank == 0) else starts
        want_norm = self.normalization_mode == 'single' and mutate_input
        for s in it:
            e = min(s + frames_per_call, hi)
            if self._keep is not None:     # test hook: one chunk at a time, tiles copied out
                for c in range(s, e, chunk):
                    ce = min(c + chunk, e)
                    res = ses.predict_device(P.to_device_stack(imgs[c:ce], self.device), keep=True)
                    if want_norm:
                        imgs[c:ce] = ses.last['norm'].cpu().numpy()
                    out[c - lo:ce - lo] = res.cpu().numpy()
                    self._keep['patches'].append(ses.last['tiles'].cpu().numpy())
                    self._keep['result_patches'].append(ses.last['result_tiles'].cpu().numpy())
                continue
            res, norm = ses.predict_movie(imgs[s:e], chunk_frames=chunk, want_norm=want_norm)
            if want_norm:
                imgs[s:e] = norm                                # unet/predict.py:131 (cast back to the input dtype)
            out[s - lo:e - lo] = res
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
