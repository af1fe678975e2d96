#!/usr/bin/env python
"""Benchmark of the tiled U-Net prediction path (BASELINE.json: output megapixels / second, voxels / second for 3D).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4,5}]     # this engine (default: cfg 2)
    python bench.py --impl reference ...                                            # the reference's CPU path

The headline workload is BASELINE configs[1] (cfg 2): Unet(n_filter=32) Predict on a 2048x2048 uint16 time-lapse,
512x512 tiles, add_tile=1 (25 overlapping tiles per frame), per-frame percentile normalisation, bf16 tensor-core
mode. A step = one pass over a FIXED movie of `--frames` frames (default 256); with N GPUs its frames are sharded
contiguously over the ranks (strong scaling) through Session.predict_movie_sharded, and the stitched uint8 slabs are
gathered on rank 0 over NCCL INSIDE the timed region. `value` = inputs resident in HBM, gathered result left in rank
0's HBM; `e2e` = pinned host stack in -> pinned host result out on rank 0 (H2D, gather and D2H timed).

The one JSON line also carries `extra_configs`: cfg 1 / 3 / 4 / 5 measured the same way (short runs, N = 1 only), each
through its own Session / Predict class with host buffers, with `e2e`, `roofline`, a truncated-slice `cpu_baseline`
and `parity` (its own output against the CPU reference outside the timed region).

CPU arm (`--impl reference`, `cpu_baseline`): the UNMODIFIED reference package (baseline/_ref, installed by
__graft_entry__.build(); kind "reference") with torch on all host cores; the oracle port (kind "port") if absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# 2 * sum(M*N*K) of the reference layer shapes
FLOP_UNET32_PER_TILE_PX = 367232
FLOP_UNET32_PER_TILE_PX_TC = 367232 - 2 * (9 * 32 + 32)      # minus encode1 and the 1x1 head (not tcgen05 launches)
FLOP_SIAM32_PER_TILE_PX = 478400
FLOP_UNET3D16_PER_VOXEL = 109872
FLOP_MO3D16_PER_VOXEL = 203040 + 16 * 3


# ----------------------------------------------------------------------------------------------------------------
# synthetic data and weights
# ----------------------------------------------------------------------------------------------------------------
def synth_frames(n, shape, seed0=0, distinct=32):
    """Uniform 12-bit noise, default_rng(seed=frame) per frame; `distinct` frames are generated and
    repeated (every frame is still far larger than would stay in L2 between its uses)."""
    base = [np.random.default_rng(seed0 + i).integers(0, 4096, shape).astype('uint16') for i in range(min(n, distinct))]
    out = np.empty((n, *shape), dtype='uint16')
    for i in range(n):
        out[i] = base[i % len(base)]
    return out


def kaiming_state_dict(module, seed):
    """Random-init weights of the architecture as the reference's Trainer initialises them (Kaiming-normal conv
    weights, utils/utils.py:76-78) plus randomised BatchNorm statistics, so that the folded scale / shift path is
    exercised and the output spreads over the whole range (PyTorch's default init gives 134..136 everywhere)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k, v in sd.items():
        if k.endswith('num_batches_tracked') or not v.is_floating_point():
            continue
        if v.dim() >= 4:
            is_up = k.startswith('up') and k.count('.') == 1
            fan_in = v.shape[0] if is_up else v[0].numel()
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / 1.01 / fan_in) ** 0.5
        elif k.endswith('running_var') or k.endswith('.1.weight'):
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('.bias') and not k.endswith('.1.bias'):
            sd[k] = torch.randn(v.shape, generator=g) * 0.05
        else:
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
    return sd


def unit_logit_head(sd, logits, w_key, b_key):
    sd[w_key] = sd[w_key] / logits.std()
    sd[b_key] = (sd[b_key] - logits.mean()) / logits.std()
    return sd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r[1:] for r in self.rows if len(r) >= 9]
        if self.t_begin is not None and self.t_end is not None:
            inside = [r[1:] for r in self.rows if len(r) >= 9 and self.t_begin <= r[0] <= self.t_end + 0.06]
            if inside:
                rows = inside
            elif rows:
                near = sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (self.t_begin + self.t_end)))[:2]
                rows = [r[1:] for r in near if len(r) >= 9]
        sm = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows for i in range(4) if r[4 + i].lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return p, 'MEASURED_PEAKS.json'
    except Exception:  # noqa: BLE001
        return {'bf16_tflops_sustained': 1400.0, 'bf16_tflops': 1670.0, 'hbm_gbs': 6500.0}, 'fallback (B200_PROFILING.md)'


def load_traffic():
    """Per-forward DRAM traffic of the tcgen05 conv launches from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:  # noqa: BLE001
        return {}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref) or the oracle port
# ----------------------------------------------------------------------------------------------------------------
class CpuReference:
    """Runs the reference's own Predict classes on the host cores (all of them, torch intra-op threads)."""

    def __init__(self):
        import torch
        self.cores = os.cpu_count()
        torch.set_num_threads(self.cores)
        self.ref, self.kind = None, 'port'
        try:
            from oracle import ref_import
            if ref_import.available():
                self.ref = ref_import.import_reference()
                self.store = ref_import.TIFF_STORE
                self.kind = 'reference'
        except Exception as e:  # noqa: BLE001
            print(f'[bench] reference package not importable ({e}); using the oracle port', file=sys.stderr)
        self.tmp = tempfile.mkdtemp(prefix='biu_cpu_')

    def _ckpt(self, params, name):
        import torch
        path = os.path.join(self.tmp, name)
        torch.save(params, path)
        return path

    def unet(self, frames, params, tile, add_tile):
        if self.ref is not None:
            from bio_image_unet.unet import Predict
            Predict(frames.copy(), 'cpu_unet.tif', self._ckpt(params, 'unet.pt'), network='Unet', resize_dim=tile,
                    add_tile=add_tile, show_progress=False, device='cpu')
            return np.asarray(self.store['cpu_unet.tif']).astype(np.float32)
        from oracle import pipeline as opipe
        return opipe.unet_predict(frames.copy(), params['state_dict'], tile, False, 'single', (0., 99.8), add_tile).astype(np.float32)

    def siam(self, movie, params, tile, add_tile):
        if self.ref is not None:
            from bio_image_unet.siam_unet import Predict
            self.store['cpu_siam_in.tif'] = movie.copy()
            cwd = os.getcwd()
            os.chdir(self.tmp)                      # the reference creates a temp_<file> directory in the cwd
            try:
                Predict('cpu_siam_in.tif', 'cpu_siam.tif', self._ckpt(params, 'siam.pt'), resize_dim=tile,
                        add_tile=add_tile, show_progress=False, device='cpu')
            finally:
                os.chdir(cwd)
            return np.asarray(self.store['cpu_siam.tif']).astype(np.float32)
        from oracle import pipeline as opipe
        return opipe.siam_predict(movie.copy(), params['state_dict'], params['mode'], tile, False, 'single', (0., 99.98),
                                  add_tile).astype(np.float32)

    def unet3d(self, vol, params, patch, add_patch):
        if self.ref is not None:
            from bio_image_unet.unet3d import Predict
            Predict(vol.copy(), 'cpu_u3d.tif', self._ckpt(params, 'u3d.pt'), resize_dim=patch, add_patch=add_patch,
                    progress_bar=False, device='cpu')
            return np.asarray(self.store['cpu_u3d.tif']).astype(np.float32)
        from oracle import pipeline as opipe
        return opipe.unet3d_predict(vol.copy(), params['state_dict'], patch, False, (0., 99.8), add_patch).astype(np.float32)

    def mo3d(self, imgs, params, max_patch, overlap):
        heads = params['output_heads']
        if self.ref is not None:
            from bio_image_unet.multi_output_unet3d import Predict
            p = Predict(imgs.astype('float32'), self._ckpt(params, 'mo3d.pt'), result_path=None, max_patch_size=max_patch,
                        overlap_factor=overlap, normalization_mode='all', show_progress=False, device='cpu')
            return {k: np.asarray(v, dtype=np.float32) for k, v in p.result.items()}
        from oracle import pipeline as opipe
        return opipe.mo3d_predict(imgs.astype('float32'), params['state_dict'], heads, True, max_patch, overlap, 'all',
                                  (0., 99.98))

    def timed(self, fn, reps=1, warm=None):
        if warm is not None:
            warm()
        best, out = None, None
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, out


# ----------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE config. Sub-classes provide setup(), step_device(), step_e2e(), cpu_sample(), parity()."""
    name = ''
    metric, unit = 'megapixels/sec', 'MP/s'
    dtype = 'bf16'
    scaling = 'strong'

    def __init__(self, args, device, ctx):
        self.args, self.device, self.ctx = args, device, ctx
        self.precision = args.precision or self.dtype

    # units processed by ALL ranks in one step, and the algorithmic FLOPs of one step
    units_per_step = 0.0
    flop_per_step = 0.0
    h2d_bytes = d2h_bytes = 0

    def engine(self):
        raise NotImplementedError

    def finish_e2e(self):
        pass


class Cfg2(Workload):
    name = 'unet2d_nf32_2048x2048_tiles512_addtile1 (BASELINE configs[1])'
    FRAME, TILE, ADD = (2048, 2048), (512, 512), 1

    def setup(self):
        import torch
        from bio_image_unet_b200.unet import Session, Unet
        from oracle import models as omodels
        a = self.args
        self.n_total = a.frames or 256
        self.lo, self.hi = self.ctx.shard(self.n_total)
        torch.manual_seed(0)
        sd = kaiming_state_dict(Unet(n_filter=32), 0)
        probe = torch.from_numpy(synth_frames(1, (256, 256), 999)[0].astype('float32') / 4096)[None, None]
        with torch.no_grad():
            sd = unit_logit_head(sd, omodels.unet_forward(sd, probe)[1], 'final.0.weight', 'final.0.bias')
        self.params = {'state_dict': sd, 'n_filter': 32, 'in_channels': 1, 'out_channels': 1}
        self.ses = Session(self.params, resize_dim=self.TILE, add_tile=self.ADD, normalization_mode='single',
                           clip_threshold=(0., 99.8), device=self.device, precision=self.precision, workspace_gb=40.0)
        f = self.hi - self.lo
        movie = synth_frames(self.n_total, self.FRAME)[self.lo:self.hi]
        self.host = torch.from_numpy(movie).pin_memory()
        self.dev = self.host.to(self.device)
        self.tiles_per_frame = 25
        self.units_per_step = self.n_total * self.FRAME[0] * self.FRAME[1] / 1e6
        self.flop_per_step = self.n_total * self.tiles_per_frame * self.TILE[0] * self.TILE[1] * FLOP_UNET32_PER_TILE_PX
        self.h2d_bytes = self.n_total * self.FRAME[0] * self.FRAME[1] * 2
        self.d2h_bytes = self.n_total * self.FRAME[0] * self.FRAME[1]
        self.chunk = a.chunk_frames
        self.config = {'workload': self.name, 'frames_per_step': self.n_total, 'frames_per_step_per_gpu': f,
                       'tiles_per_step': self.n_total * self.tiles_per_frame, 'chunk_frames': self.chunk,
                       'normalization_mode': 'single', 'weights': 'Kaiming-normal + randomised BatchNorm statistics, seed 0',
                       'l2_policy': f'every step streams the whole resident movie ({f * 8} MiB per GPU) and >1 GiB of '
                                    f'activations per forward, both larger than the 126 MB L2'}
        return self

    def engine(self):
        return self.ses.engine

    def step_device(self):
        return self.ses.predict_movie_sharded(self.dev, self.ctx, self.n_total, chunk_frames=self.chunk, to_host=False)

    def step_e2e(self):
        return self.ses.predict_movie_sharded(self.host, self.ctx, self.n_total, chunk_frames=self.chunk, to_host=True)

    def comm(self):
        return dict(self.ses.comm)

    def cpu_sample(self, cpu):
        frame = self.host[0:1].numpy().copy() if self.lo == 0 else synth_frames(1, self.FRAME)
        dt, out = cpu.timed(lambda: cpu.unet(frame, self.params, self.TILE, self.ADD), reps=2)     # best of 2: the first call pays oneDNN's set-up
        self._cpu_out = out
        return dt, frame.shape[1] * frame.shape[2] / 1e6, \
            'one full 2048x2048 frame of the movie (25 tiles of 512x512, add_tile=1) through unet.Predict(device="cpu"), best of 2 calls'

    def parity(self, result):
        """Frame 0 of the engine's end-to-end result against the CPU reference's output for the same frame."""
        if result is None or getattr(self, '_cpu_out', None) is None:
            return None
        d = np.abs(np.squeeze(result[0]).astype(np.float32) - np.squeeze(self._cpu_out))
        return {'against': 'CPU reference, frame 0 (stitched uint8 result)', 'lsb_max': float(d.max()),
                'lsb_mean': float(d.mean()), 'pixels': int(d.size), 'result_std_lsb': float(np.std(self._cpu_out))}


class Cfg1(Workload):
    name = 'unet2d_nf32_1024x1024_tiles256_addtile1 (BASELINE configs[0])'
    dtype = 'tf32'

    def setup(self):
        import torch
        from bio_image_unet_b200.unet import Session, Unet
        from oracle import models as omodels
        torch.manual_seed(0)
        sd = kaiming_state_dict(Unet(n_filter=32), 0)
        probe = torch.from_numpy(synth_frames(1, (256, 256), 999)[0].astype('float32') / 4096)[None, None]
        with torch.no_grad():
            sd = unit_logit_head(sd, omodels.unet_forward(sd, probe)[1], 'final.0.weight', 'final.0.bias')
        self.params = {'state_dict': sd, 'n_filter': 32, 'in_channels': 1, 'out_channels': 1}
        self.ses = Session(self.params, resize_dim=(256, 256), add_tile=1, device=self.device, precision=self.precision,
                           workspace_gb=8.0)
        img = np.random.default_rng(0).integers(0, 4096, (1024, 1024)).astype('uint16')
        self.host = torch.from_numpy(img[None]).pin_memory()
        self.dev = self.host.to(self.device)
        self.units_per_step = 1024 * 1024 / 1e6
        self.flop_per_step = 25 * 256 * 256 * FLOP_UNET32_PER_TILE_PX
        self.h2d_bytes, self.d2h_bytes = 1024 * 1024 * 2, 1024 * 1024
        self.config = {'workload': self.name, 'tiles_per_step': 25, 'precision': self.precision,
                       'note': 'BASELINE says fp32: run in the fp32-storage TF32 tensor-core mode (the Predict default, '
                               'north_star "fp32/TF32 mode"); one image per step, so launch latency dominates'}
        return self

    def engine(self):
        return self.ses.engine

    def step_device(self):
        return self.ses.predict_device(self.dev)

    def step_e2e(self):
        return self.ses.predict_movie(self.host)[0]

    def cpu_sample(self, cpu):
        img = self.host.numpy()[0].copy()
        dt, out = cpu.timed(lambda: cpu.unet(img, self.params, (256, 256), 1))
        self._cpu_out = out
        return dt, 1024 * 1024 / 1e6, 'the whole cfg-1 image (25 tiles of 256x256) through unet.Predict(device="cpu")'

    def parity(self, result):
        d = np.abs(np.squeeze(result).astype(np.float32) - np.squeeze(self._cpu_out))
        return {'against': 'CPU reference, whole image', 'lsb_max': float(d.max()), 'lsb_mean': float(d.mean()),
                'pixels': int(d.size), 'result_std_lsb': float(np.std(self._cpu_out))}


class Cfg3(Workload):
    name = 'siam_unet_nf32_concat_1024x1024_tiles512_addtile1 (BASELINE configs[2])'

    def setup(self):
        import torch
        from bio_image_unet_b200.siam_unet import Session, Siam_UNet
        from bio_image_unet_b200.siam_unet.predict import _ArraySource
        from oracle import models as omodels
        torch.manual_seed(0)
        sd = kaiming_state_dict(Siam_UNet(n_filter=32, mode='concat'), 0)
        pr = torch.from_numpy(synth_frames(2, (128, 128), 999).astype('float32') / 4096)[:, None]
        with torch.no_grad():
            sd = unit_logit_head(sd, omodels.siam_forward(sd, pr[:1], pr[1:], 'concat')[1], 'final.0.weight', 'final.0.bias')
        self.params = {'state_dict': sd, 'n_filter': 32, 'mode': 'concat'}
        self.ses = Session(self.params, resize_dim=(512, 512), add_tile=1, device=self.device, precision=self.precision,
                           workspace_gb=40.0)
        self.n = self.args.frames or 64
        movie = synth_frames(self.n, (1024, 1024))
        self.host = torch.from_numpy(movie).pin_memory()
        self.source = _ArraySource(self.host.numpy())
        self.dev = self.host.to(self.device)
        self.units_per_step = self.n * 1024 * 1024 / 1e6
        self.flop_per_step = self.n * 9 * 512 * 512 * FLOP_SIAM32_PER_TILE_PX
        self.h2d_bytes, self.d2h_bytes = self.n * 1024 * 1024 * 2, self.n * 1024 * 1024
        self.out = torch.empty((self.n, 1024, 1024), dtype=torch.uint8).pin_memory()
        # 'single' normalisation: frame t's encoder pass serves pair t and pair t + 1, so it is executed once
        # (SURVEY.md §8d cfg 3): 3341 of the 15676 MMAC per tile pair are not executed
        enc_px = 2 * 3341e6 / (256 * 256)
        pairs_per_fwd = 144
        self.executed_flop_per_step = self.flop_per_step - self.n * 9 * 512 * 512 * enc_px * (pairs_per_fwd - 9) / pairs_per_fwd
        self.config = {'workload': self.name, 'frames_per_step': self.n, 'tile_pairs_per_step': self.n * 9,
                       'flops': 'roofline / TFLOP figures use the REFERENCE count (both encoder passes of every pair, '
                                '478 400 FLOP per tile px); the engine executes the shared-weight encoder once per frame '
                                '(normalization_mode "single"), see executed_flop_fraction',
                       'executed_flop_fraction': self.executed_flop_per_step / self.flop_per_step}
        # resident leg: all pairs of the step on the device at once, frames in upload order
        needed, p_pos, c_pos = self.ses.chunk_frames(self.n, 0, self.n)
        self.dev = self.dev.view(torch.int16)[torch.tensor(needed, device=self.device)].contiguous().view(torch.uint16)
        self.p_sel = torch.tensor(p_pos, device=self.device)
        self.c_sel = torch.tensor(c_pos, device=self.device)
        rd, n_x, n_y, _, _ = self.ses.grid(1024, 1024)
        self.ses._ensure_plan(rd, 16 * n_x * n_y, n_x * n_y)
        return self

    def engine(self):
        return self.ses.engine

    def step_device(self):
        return self.ses.predict_pairs_device(self.dev, self.p_sel, self.c_sel)

    def step_e2e(self):
        out = self.out.numpy()

        def sink(first, pages):
            out[first:first + len(pages)] = pages
        self.ses.predict_stream(self.source, 0, self.n, sink, chunk_pairs=16)
        return out

    def cpu_sample(self, cpu):
        movie = self.host.numpy()[:2].copy()
        dt, out = cpu.timed(lambda: cpu.siam(movie, self.params, (512, 512), 1))
        self._cpu_out = out
        return dt, 2 * 1024 * 1024 / 1e6, 'the first 2 frames of the movie (2 pairs x 9 tile pairs of 512x512) through siam_unet.Predict(device="cpu")'

    def parity(self, result):
        # pair 0 = (frame 1, frame 0), pair 1 = (frame 0, frame 1): identical in the 2-frame sample and the full movie
        d = np.abs(result[:2].astype(np.float32) - self._cpu_out.reshape(2, 1024, 1024))
        return {'against': 'CPU reference, frames 0-1', 'lsb_max': float(d.max()), 'lsb_mean': float(d.mean()),
                'pixels': int(d.size), 'result_std_lsb': float(np.std(self._cpu_out))}


class Cfg4(Workload):
    name = 'unet3d_nf16_256x1024x1024_patches64x128x128_addpatch0 (BASELINE configs[3])'
    metric, unit = 'voxels/sec', 'Mvoxel/s'
    VOL, PATCH = (256, 1024, 1024), (64, 128, 128)

    def setup(self):
        import torch
        from bio_image_unet_b200.unet3d import Session, UNet3D
        from oracle import models as omodels
        torch.manual_seed(0)
        sd = kaiming_state_dict(UNet3D(n_filter=16), 0)
        pr = torch.rand((1, 1, 16, 32, 32), generator=torch.Generator().manual_seed(9))
        with torch.no_grad():
            sd = unit_logit_head(sd, omodels.unet3d_forward(sd, pr)[1], 'final.weight', 'final.bias')
        self.params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'out_channels': 1}
        self.ses = Session(self.params, self.PATCH, add_patch=self.args.add_patch, device=self.device,
                           precision=self.precision, workspace_gb=60.0, dist=self.ctx)
        if getattr(self.args, 'vol', None):          # profiling runs: a smaller volume (tools/profile_step.py --vol)
            self.VOL = tuple(int(v) for v in self.args.vol.split(','))
        vol = np.random.default_rng(0).integers(0, 4096, self.VOL).astype('uint16')
        self.host = torch.from_numpy(vol).pin_memory()
        self.dev = self.host.to(self.device)
        nvox = float(np.prod(self.VOL))
        self.units_per_step = nvox / 1e6
        n_patches = int(np.prod([-(-v // p) for v, p in zip(self.VOL, self.PATCH)])) if self.args.add_patch == 0 else 550
        self.flop_per_step = n_patches * float(np.prod(self.PATCH)) * FLOP_UNET3D16_PER_VOXEL
        self.h2d_bytes, self.d2h_bytes = int(nvox) * 2, int(nvox)
        self.config = {'workload': self.name, 'patches_per_step': n_patches, 'add_patch': self.args.add_patch,
                       'sharding': 'z-rows of the patch grid per rank, histogram all-reduce, boundary patches by send/recv, '
                                   'slab gather on rank 0'}
        return self

    def engine(self):
        return self.ses.engine

    def step_device(self):
        return self.ses.predict(self.dev, to_host=False)

    def step_e2e(self):
        return self.ses.predict(self.host, to_host=True)

    def cpu_sample(self, cpu):
        self._sub = self.host.numpy()[:64, :256, :512].copy()
        dt, out = cpu.timed(lambda: cpu.unet3d(self._sub, self.params, self.PATCH, 0))
        self._cpu_out = out
        return dt, self._sub.size / 1e6, 'a 64x256x512 corner of the volume (8 patches of 64x128x128) through unet3d.Predict(device="cpu")'

    def parity(self, result):
        import torch
        got = self.ses.predict(torch.from_numpy(self._sub), to_host=True)      # same sub-volume (global percentiles!)
        d = np.abs(np.squeeze(got).astype(np.float32) - np.squeeze(self._cpu_out))
        return {'against': 'CPU reference, the 64x256x512 sample volume', 'lsb_max': float(d.max()), 'lsb_mean': float(d.mean()),
                'pixels': int(d.size), 'result_std_lsb': float(np.std(self._cpu_out))}


class Cfg5(Workload):
    name = 'mo_unet3d_nf16_3heads_64x512x512_patches64x256x256 (BASELINE configs[4])'
    metric, unit = 'voxels/sec', 'Mvoxel/s'

    def setup(self):
        import torch
        from bio_image_unet_b200.multi_output_unet3d import MultiOutputUnet3D, Session
        self.heads = {f'h{i}': {'channels': 1, 'activation': 'sigmoid'} for i in range(3)}
        torch.manual_seed(0)
        sd = kaiming_state_dict(MultiOutputUnet3D(1, self.heads, 16, True), 0)
        self.params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'output_heads': self.heads,
                       'use_interpolation': True}
        self.ses = Session(self.params, (64, 256, 256), 0.1, 'all', (0., 99.98), self.device, self.precision, 60.0,
                           dist=self.ctx)
        self.n = self.args.frames or 4
        vols = np.stack([np.random.default_rng(i).integers(0, 4096, (64, 512, 512)).astype('uint16') for i in range(self.n)])
        self.host = torch.from_numpy(vols).pin_memory()
        self.dev = self.host.to(self.device)
        nvox = float(vols.size)
        self.units_per_step = nvox / 1e6
        self.flop_per_step = self.n * 9 * 64 * 256 * 256 * FLOP_MO3D16_PER_VOXEL
        self.h2d_bytes, self.d2h_bytes = int(nvox) * 2, int(nvox) * 3 * 4
        self.config = {'workload': self.name, 'volumes_per_step': self.n, 'patches_per_step': self.n * 9,
                       'normalization_mode': 'all (the reference\'s "single" fails on numpy >= 2)'}
        return self

    def engine(self):
        return self.ses.engine

    def step_device(self):
        return self.ses.predict(self.dev, to_host=False)

    def step_e2e(self):
        return self.ses.predict(self.host, to_host=True)

    def cpu_sample(self, cpu):
        self._sub = self.host.numpy()[:1, :, :256, :256].copy()
        dt, out = cpu.timed(lambda: cpu.mo3d(self._sub, self.params, (64, 256, 256), 0.1))
        self._cpu_out = out
        return dt, self._sub.size / 1e6, 'one 64x256x256 volume (1 patch) through multi_output_unet3d.Predict(device="cpu", normalization_mode="all")'

    def parity(self, result):
        import torch
        got = self.ses.predict(torch.from_numpy(self._sub), to_host=True)
        d = np.concatenate([np.abs(np.squeeze(got[k]) - np.squeeze(self._cpu_out[k])).ravel() for k in self.heads])
        return {'against': 'CPU reference, the 64x256x256 sample volume, 3 sigmoid heads (float32)', 'abs_max': float(d.max()),
                'abs_mean': float(d.mean()), 'voxels': int(d.size)}


WORKLOADS = {1: Cfg1, 2: Cfg2, 3: Cfg3, 4: Cfg4, 5: Cfg5}


# ----------------------------------------------------------------------------------------------------------------
# measurement
# ----------------------------------------------------------------------------------------------------------------
def measure(wl, steps, warmup, ctx, lib, local_rank, with_cpu, cpu, sample_clocks=True):
    """Device-resident leg, per-kernel profile, end-to-end leg, CPU baseline, parity. Returns the JSON record."""
    import torch
    import torch.distributed as dist
    device = wl.device
    world, rank = ctx.world, ctx.rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    eng = wl.engine()
    # ---------------- device-resident leg -------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        wl.step_device()
    eng.set_profile(1)
    barrier()
    if sampler:
        sampler.mark_begin()
    launches0 = lib.biu_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        wl.step_device()
    ev1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = lib.biu_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    value = wl.units_per_step * steps / (dev_ms / 1e3)
    comm = wl.comm() if hasattr(wl, 'comm') else None

    # per-kernel timing of the dominant kernels (tcgen05 implicit-GEMM convs): CUDA events on the launching stream around
    # every op of every timed forward (set_profile above); the ones of the LAST forward of the timed loop are read here
    try:
        kinds, ms = eng.read_profile()
    except Exception:  # noqa: BLE001  (a rank that owns no work - more ranks than z-rows of patches - ran no forward)
        kinds, ms = [], []
    # ... averaged with the last forward of a few more steps (outside the timed region, same state): one forward's events
    # scatter by +-3 % from run to run
    # (every rank runs the extra steps - they contain the collectives of the sharded path - whether it owns work or not)
    acc, n_acc = np.array(ms, dtype=np.float64), 1
    for _ in range(min(4, max(1, steps))):
        wl.step_device()
        torch.cuda.synchronize()
        if not kinds:
            continue
        k2, m2 = eng.read_profile()
        if list(k2) == list(kinds):
            acc += np.array(m2, dtype=np.float64)
            n_acc += 1
    if kinds:
        ms = list(acc / n_acc)
    eng.set_profile(0)
    tc_ms = sum(m for k, m in zip(kinds, ms) if k in (1, 2, 3))
    tc_launches = sum(1 for k in kinds if k in (1, 2, 3))
    all_ms = sum(ms)
    fallback_ops = sum(1 for k in kinds if 16 <= k < 32)
    if rank == 0 and os.environ.get('BIU_BENCH_VERBOSE'):
        names = {0: 'first_conv', 1: 'conv_tc', 2: 'conv_tc+head', 3: 'up_tc', 4: 'pool', 5: 'up_nearest', 6: 'max_join'}
        for i, (k, m) in enumerate(zip(kinds, ms)):
            tag = ' (cuda-core)' if 16 <= k < 32 else (' (fused)' if k >= 32 else '')
            print(f'[{wl.name[:12]} op {i:2d}] {names.get(k % 16, "?"):14s}{tag:12s} {m:8.3f} ms', file=sys.stderr)
    peaks, peak_src = load_peaks()
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    step_tf = wl.flop_per_step / world * steps / (dev_ms / 1e3) / 1e12          # per GPU
    batch = eng.batch
    d, h, w = eng.tile if eng.tile else (0, 0, 0)
    per_px = {'unet2d': FLOP_UNET32_PER_TILE_PX_TC, 'siam2d': FLOP_SIAM32_PER_TILE_PX - 2 * 2 * (9 * 32) - 2 * 32,
              'unet3d': FLOP_UNET3D16_PER_VOXEL - 2 * (27 * 8) - 2 * 8, 'mo3d': FLOP_MO3D16_PER_VOXEL - 2 * (27 * 8) - 2 * 8 * 3}[eng.kind]
    flops_last_fwd = per_px * batch * d * h * w                                  # padded tail batches compute full batches
    achieved_tf = flops_last_fwd / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    traffic = load_traffic().get(wl.name.split(' ')[0])
    roofline = {'bound': 'tensor',
                'kernel': f'biu::conv_rows_kernel / biu::conv_halo_kernel: the {tc_launches} tcgen05 conv / transposed-conv launches of one forward',
                'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf,
                'peak_source': f'{peak_src} bf16_tflops_sustained (kernels timed inside a long step)',
                'traffic': (traffic or {}).get('bytes_per_forward'), 'traffic_source': (traffic or {}).get('source'),
                'tc_ms_per_forward': tc_ms, 'all_ops_ms_per_forward': all_ms, 'tc_launches_per_forward': tc_launches,
                'tiles_per_forward': batch, 'cuda_core_fallback_ops': fallback_ops,
                'whole_step_tflops_per_gpu': step_tf, 'whole_step_frac': step_tf / peak_tf}

    # ---------------- end-to-end leg: pinned host buffers through the public Session API ----------------------------
    result = wl.step_e2e()                                                        # warm-up: allocates pinned buffers
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = max(1, steps)
    for _ in range(e2e_steps):
        result = wl.step_e2e()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_val = wl.units_per_step * e2e_steps / (e2e_ms / 1e3)
    comm_e2e = wl.comm() if hasattr(wl, 'comm') else None

    # ---------------- CPU baseline + parity (rank 0, N = 1) -----------------------------------------------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and with_cpu:
        dt, units, sample = wl.cpu_sample(cpu)
        cpu_baseline = {'value': units / dt, 'unit': wl.unit, 'cores': cpu.cores, 'kind': cpu.kind,
                        'sample': f'{sample}, torch CPU fp32, {cpu.cores} threads, {dt:.1f} s per call'}
        try:
            parity = wl.parity(result)
            if parity is not None:
                parity['precision'] = wl.precision
        except Exception as e:  # noqa: BLE001
            parity = {'error': repr(e)}

    rec = {'metric': wl.metric, 'value': value, 'unit': wl.unit, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
           'ms_per_step': dev_ms / steps, 'higher_is_better': True, 'scaling': wl.scaling, 'vs_baseline': None,
           'dtype': wl.precision, 'data': 'synthetic', 'config': wl.config, 'clocks': clocks, 'gpu_launches': int(launches),
           'e2e': {'value': e2e_val, 'unit': wl.unit, 'h2d_bytes_per_step': wl.h2d_bytes, 'd2h_bytes_per_step': wl.d2h_bytes,
                   'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps},
           'roofline': roofline, 'cpu_baseline': cpu_baseline, 'parity': parity}
    if world > 1 and comm is not None:
        rec['multi_gpu'] = {'collective': comm['collective'], 'bytes_per_step': comm['bytes'],
                            'ms': comm['ms'], 'gbps_on_rank0': (comm['bytes'] / 1e9) / (comm['ms'] / 1e3) if comm['ms'] > 0 else None,
                            'e2e_ms': comm_e2e['ms'] if comm_e2e else None,
                            'note': 'ms = time of the gather on rank 0\'s communication stream, summed over the segments of the '
                                    'last step (includes waiting for the slowest sender); it overlaps the next segment\'s compute'}
    return rec


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores; one bounded sample of the
    workload per step (cfg 2: one full 2048x2048 frame = 25 tiles)."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cpu = CpuReference()
    # weights / inputs are built without touching CUDA
    params, sample_fn, units, what = reference_sample(args.config, cpu)
    warm = min(args.warmup, 2)
    for _ in range(warm):
        sample_fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sample_fn()
    dt = (time.perf_counter() - t0) / args.steps
    val = units / dt
    unit = WORKLOADS[args.config].unit
    print(json.dumps({
        'impl': 'reference', 'metric': WORKLOADS[args.config].metric, 'value': val, 'unit': unit, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': warm, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.config].name, 'sample_per_step': what, 'warmup_note': 'CPU warm-up capped at 2 steps'},
        'cpu_baseline': {'value': val, 'unit': unit, 'cores': cpu.cores, 'kind': cpu.kind,
                         'sample': f'{args.steps} x {what}, torch CPU fp32, {cpu.cores} threads'},
        'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def reference_sample(config, cpu):
    """(params, callable, units per call, description) of the CPU arm's bounded sample for `config` - the same
    weights and the same synthetic inputs as the engine's arm, built without CUDA."""
    import torch
    from oracle import models as omodels
    torch.manual_seed(0)
    if config in (1, 2):
        from bio_image_unet_b200.unet import Unet
        sd = kaiming_state_dict(Unet(n_filter=32), 0)
        probe = torch.from_numpy(synth_frames(1, (256, 256), 999)[0].astype('float32') / 4096)[None, None]
        with torch.no_grad():
            sd = unit_logit_head(sd, omodels.unet_forward(sd, probe)[1], 'final.0.weight', 'final.0.bias')
        params = {'state_dict': sd, 'n_filter': 32, 'in_channels': 1, 'out_channels': 1}
        if config == 2:
            frame = synth_frames(1, (2048, 2048))
            return params, (lambda: cpu.unet(frame, params, (512, 512), 1)), 2048 * 2048 / 1e6, \
                'one full 2048x2048 frame (25 tiles of 512x512, add_tile=1) through unet.Predict(device="cpu")'
        img = np.random.default_rng(0).integers(0, 4096, (1024, 1024)).astype('uint16')
        return params, (lambda: cpu.unet(img, params, (256, 256), 1)), 1024 * 1024 / 1e6, \
            'the whole 1024x1024 image (25 tiles of 256x256) through unet.Predict(device="cpu")'
    if config == 3:
        from bio_image_unet_b200.siam_unet import Siam_UNet
        sd = kaiming_state_dict(Siam_UNet(n_filter=32, mode='concat'), 0)
        params = {'state_dict': sd, 'n_filter': 32, 'mode': 'concat'}
        movie = synth_frames(2, (1024, 1024))
        return params, (lambda: cpu.siam(movie, params, (512, 512), 1)), 2 * 1024 * 1024 / 1e6, \
            '2 frames of 1024x1024 (2 pairs x 9 tile pairs) through siam_unet.Predict(device="cpu")'
    if config == 4:
        from bio_image_unet_b200.unet3d import UNet3D
        sd = kaiming_state_dict(UNet3D(n_filter=16), 0)
        params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'out_channels': 1}
        sub = np.random.default_rng(0).integers(0, 4096, (64, 256, 512)).astype('uint16')
        return params, (lambda: cpu.unet3d(sub, params, (64, 128, 128), 0)), sub.size / 1e6, \
            'a 64x256x512 volume (8 patches of 64x128x128) through unet3d.Predict(device="cpu")'
    from bio_image_unet_b200.multi_output_unet3d import MultiOutputUnet3D
    heads = {f'h{i}': {'channels': 1, 'activation': 'sigmoid'} for i in range(3)}
    sd = kaiming_state_dict(MultiOutputUnet3D(1, heads, 16, True), 0)
    params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'output_heads': heads, 'use_interpolation': True}
    sub = np.random.default_rng(0).integers(0, 4096, (1, 64, 256, 256)).astype('uint16')
    return params, (lambda: cpu.mo3d(sub, params, (64, 256, 256), 0.1)), sub.size / 1e6, \
        'one 64x256x256 volume (1 patch) through multi_output_unet3d.Predict(device="cpu")'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument('--frames', type=int, default=None, help='frames (cfg 2/3) or volumes (cfg 5) per step')
    ap.add_argument('--chunk-frames', type=int, default=8)
    ap.add_argument('--add-patch', type=int, default=0)
    ap.add_argument('--vol', default=None, help=argparse.SUPPRESS)
    ap.add_argument('--precision', default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--extra-steps', type=int, default=5)
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from bio_image_unet_b200 import _lib
    from bio_image_unet_b200.dist import DistContext

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    lib = _lib.load()
    ctx = DistContext(world > 1)
    with_cpu = not args.no_cpu_baseline
    cpu = CpuReference() if (rank == 0 and world == 1 and with_cpu) else None

    if world > 1 and args.config in (1, 3):
        raise SystemExit(f'bench.py --config {args.config} is a single-GPU line (cfg 2, 4 and 5 shard over ranks)')
    wl = WORKLOADS[args.config](args, device, ctx).setup()
    rec = measure(wl, args.steps, max(args.warmup, 3), ctx, lib, local_rank, with_cpu, cpu)
    rec['warmup'] = max(args.warmup, 3)
    del wl
    torch.cuda.empty_cache()

    extras = []
    if world == 1 and not args.no_extras:
        for k in (1, 3, 4, 5):
            if k == args.config:
                continue
            try:
                sub = argparse.Namespace(**vars(args))
                sub.frames, sub.precision = None, None
                w2 = WORKLOADS[k](sub, device, ctx).setup()
                r2 = measure(w2, args.extra_steps, 3, ctx, lib, local_rank, with_cpu, cpu, sample_clocks=False)
                r2.pop('clocks', None)
                extras.append(r2)
                del w2
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                extras.append({'config': {'workload': WORKLOADS[k].name}, 'error': repr(e)})
        rec['extra_configs'] = extras
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
