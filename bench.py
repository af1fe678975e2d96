#!/usr/bin/env python
"""Headline benchmark: tiled 2D U-Net prediction throughput (output megapixels / second).

Workload (BASELINE.json configs[1]): Unet(n_filter=32) Predict on 2048x2048 uint16 frames of a synthetic time-lapse,
512x512 tiles, add_tile=1 (5x5 = 25 overlapping tiles per frame), per-frame ('single') percentile normalisation,
bf16 tensor-core mode. A step = `--frames-per-step` frames per GPU (frames are independent, so ranks shard the
movie with no data-path collective: weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this engine
    python bench.py --impl reference ...                            # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).
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

FRAME = (2048, 2048)
TILE = (512, 512)
ADD_TILE = 1
N_FILTER = 32
FLOP_PER_TILE_PX = 367232          # 2 * sum(M*N*K) per tile pixel, Unet(n_filter=32)
FLOP_PER_TILE_PX_TC = 367232 - 2 * (9 * 32 + 32)   # minus encode1 and the 1x1 head, which are not tcgen05 launches
NCU_CONV_DRAM_BYTES_PER_FORWARD = 58.50e9          # measured DRAM traffic of those launches, 200 tiles (profiles/r01g_*)


def synth_frames(n, seed0=0):
    """Uniform 12-bit noise, one default_rng(seed=frame) per frame (cfg 2)."""
    out = np.empty((n, *FRAME), dtype='uint16')
    for i in range(n):
        out[i] = np.random.default_rng(seed0 + i).integers(0, 4096, FRAME).astype('uint16')
    return out


def random_checkpoint():
    import torch
    from bio_image_unet_b200.unet import Unet
    torch.manual_seed(0)
    m = Unet(n_filter=N_FILTER)
    return {'state_dict': m.state_dict(), 'n_filter': N_FILTER, 'in_channels': 1, 'out_channels': 1}


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
        # samples taken inside the timed region (the sampler itself starts before the warm-up so that nvidia-smi is
        # already streaming when the region begins); a region shorter than the 50 ms period keeps the nearest ones
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


def cpu_reference_step(frames, sd):
    """One pass of the reference's CPU algorithm (oracle port of unet.Predict) over `frames`."""
    from oracle import pipeline as opipe
    return opipe.unet_predict(frames, sd, TILE, False, 'single', (0., 99.8), ADD_TILE)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference package itself
    cannot travel to the GPU box) on all host cores, same workload/metric, bounded sample per step."""
    import torch
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    ckpt = random_checkpoint()
    rows = 512 if args.ref_rows is None else args.ref_rows     # bounded sample: a 512 x 2048 strip of one frame
    frame = synth_frames(1)[:, :rows]
    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):
        cpu_reference_step(frame.copy(), ckpt['state_dict'])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(frame.copy(), ckpt['state_dict'])
    dt = (time.perf_counter() - t0) / args.steps
    mp = frame.shape[1] * frame.shape[2] / 1e6
    val = mp / dt
    sample = (f'{args.steps} x one {rows}x2048 strip of a cfg-2 frame (512x512 tiles, add_tile=1 -> '
              f'{(int(np.ceil(rows / 512)) + 1) * 5} tiles), oracle port of unet.Predict, torch CPU fp32, {cores} threads')
    print(json.dumps({
        'impl': 'reference', 'metric': 'megapixels/sec', 'value': val, 'unit': 'MP/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'unet2d_nf32_2048x2048_tiles512_addtile1 (BASELINE configs[1])', 'sample_rows': rows},
        'cpu_baseline': {'value': val, 'unit': 'MP/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--frames-per-step', type=int, default=8)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--ref-rows', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from bio_image_unet_b200 import _lib
    from bio_image_unet_b200.unet import Session

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    lib = _lib.load()

    f = args.frames_per_step
    ckpt = random_checkpoint()
    ses = Session(ckpt, resize_dim=TILE, add_tile=ADD_TILE, normalization_mode='single', clip_threshold=(0., 99.8),
                  device=device, precision=args.precision, workspace_gb=40.0)
    # a pool of distinct synthetic frames (3 steps' worth, > L2) so consecutive steps do not re-read cached input
    pool_steps = 3
    host_pool = torch.from_numpy(synth_frames(f * pool_steps, seed0=rank * 100000)).pin_memory()
    dev_pool = host_pool.to(device)
    tiles_per_frame = 25
    mp_per_step = f * FRAME[0] * FRAME[1] / 1e6
    tile_px_per_step = f * tiles_per_frame * TILE[0] * TILE[1]

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

    # ---------------- device-resident leg: inputs already in HBM -------------------------------------------------
    import ctypes
    lib.biu_net_set_profile(ses.engine.handle, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        ses.predict_device(dev_pool[(i % pool_steps) * f:(i % pool_steps + 1) * f])
    barrier()
    sampler.mark_begin()
    launches0 = lib.biu_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_ms, conv_launches, other_ms = 0.0, 0, 0.0
    kinds = (ctypes.c_int * 64)()
    ms = (ctypes.c_float * 64)()
    n_ops = ctypes.c_int(0)
    ev0.record()
    for i in range(args.steps):
        ses.predict_device(dev_pool[(i % pool_steps) * f:(i % pool_steps + 1) * f])
    ev1.record()
    barrier()
    sampler.mark_end()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = lib.biu_launch_count() - launches0
    clocks = sampler.stop()
    value = mp_per_step * args.steps * world / (dev_ms / 1e3)

    # per-kernel timing of the dominant kernel (tcgen05 implicit-GEMM conv), CUDA events on the launching stream:
    # every op of every timed forward was bracketed by events (biu_net_set_profile above); the ones of the LAST forward
    # of the timed region are read back here, after the closing synchronize - i.e. at the sustained, power-capped
    # clocks of the timed loop, not of a cold extra step
    fwd_per_step = int(np.ceil(f * tiles_per_frame / ses.tile_batch))
    _lib.check(lib.biu_net_profile_read(ses.engine.handle, 64, kinds, ms, ctypes.byref(n_ops)))
    tc_ms = sum(ms[i] for i in range(n_ops.value) if kinds[i] in (1, 2, 3))
    tc_launches = sum(1 for i in range(n_ops.value) if kinds[i] in (1, 2, 3))
    all_ms = sum(ms[i] for i in range(n_ops.value))
    fallback_ops = sum(1 for i in range(n_ops.value) if 16 <= kinds[i] < 32)
    if rank == 0 and os.environ.get('BIU_BENCH_VERBOSE'):
        names = {0: 'first_conv', 1: 'conv_tc', 2: 'conv_tc+head', 3: 'up_tc', 4: 'pool', 5: 'up_nearest', 6: 'max_join'}
        for i in range(n_ops.value):
            print(f'[op {i:2d}] {names.get(kinds[i] % 16, "?"):14s}{" (cuda-core)" if 16 <= kinds[i] < 32 else (" (fused)" if kinds[i] >= 32 else ""):12s} {ms[i]:8.3f} ms',
                  file=sys.stderr)
    tiles_last_fwd = f * tiles_per_frame - (fwd_per_step - 1) * ses.tile_batch
    flops_last_fwd = FLOP_PER_TILE_PX_TC * ses.tile_batch * TILE[0] * TILE[1]   # the padded tail batch computes full batches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    achieved_tf = flops_last_fwd / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    roofline = {'bound': 'tensor', 'kernel': 'biu::conv_halo_kernel / biu::conv_rows_kernel (the 21 tcgen05 conv / transposed-conv launches of one forward)',
                'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s sustained',
                'traffic': NCU_CONV_DRAM_BYTES_PER_FORWARD * ses.tile_batch / 200 if args.precision == 'bf16' else None,
                'traffic_source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum summed over the 21 launches of a 200-tile forward (profiles/r01g_ncu_conv_kernels_full.csv)',
                'tc_ms_per_forward': tc_ms, 'all_ops_ms_per_forward': all_ms,
                'tc_launches_per_forward': tc_launches, 'tiles_per_forward': ses.tile_batch,
                'cuda_core_fallback_ops': fallback_ops,
                'whole_step_tflops': FLOP_PER_TILE_PX * tile_px_per_step * args.steps / (dev_ms / 1e3) / 1e12}
    lib.biu_net_set_profile(ses.engine.handle, 0)

    # ---------------- end-to-end leg: host (pinned) buffers through the public Session.predict_movie -------------
    # The user-facing call for a movie: all K steps' frames in ONE pinned host stack; the engine moves them through
    # the device chunk by chunk (one chunk = one step's frames), every chunk's H2D copy and the D2H read of its
    # stitched result are inside the timed region (overlapped with the neighbouring chunks' compute).
    reps = -(-args.steps // pool_steps)
    host_movie = host_pool.repeat(reps, 1, 1)[:args.steps * f].contiguous().pin_memory()
    out, _ = ses.predict_movie(host_movie, chunk_frames=f)           # warm-up: allocates the pinned result buffer
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, _ = ses.predict_movie(host_movie, chunk_frames=f)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_val = mp_per_step * args.steps * world / (e2e_ms / 1e3)
    h2d = f * FRAME[0] * FRAME[1] * 2
    d2h = int(out.nbytes) // args.steps

    # ---------------- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ---------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        strip = host_pool[0:1, :512].numpy().copy()
        cpu_reference_step(strip.copy(), ckpt['state_dict'])
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            cpu_reference_step(strip.copy(), ckpt['state_dict'])
        dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {'value': strip.shape[1] * strip.shape[2] / 1e6 / dt, 'unit': 'MP/s', 'cores': cores, 'kind': 'port',
                        'sample': f'{reps} x one 512x2048 strip of a cfg-2 frame (10 tiles of 512x512), oracle port of '
                                  f'unet.Predict, torch CPU fp32'}

    if rank == 0:
        print(json.dumps({
            'metric': 'megapixels/sec', 'value': value, 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': 'unet2d_nf32_2048x2048_tiles512_addtile1 (BASELINE configs[1])',
                       'frames_per_step_per_gpu': f, 'tiles_per_step_per_gpu': f * tiles_per_frame,
                       'tile_batch': ses.tile_batch, 'normalization_mode': 'single',
                       'l2_policy': f'inputs cycle through a pool of {pool_steps} steps ({pool_steps * h2d >> 20} MiB) and every '
                                    f'forward streams >1 GiB of activations, both larger than the 126 MB L2'},
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_val, 'unit': 'MP/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms / args.steps},
            'roofline': roofline, 'cpu_baseline': cpu_baseline,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
