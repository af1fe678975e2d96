"""Per-op device times of one forward of a BASELINE config (bench.py workload objects, CUDA events around every op of the
network handle): `python tools/ops_profile.py 4`. Op kinds: 0 first block, 1 conv block, 2 conv block + heads, 3 transposed
conv, 4 pool, 5 nearest upsample (+16: CUDA-core fallback, +32: fused into the previous op). Set BIU_PLAN_DEBUG=1 to print
the row / plane plans of the 3D blocks."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, argparse
dev = torch.device('cuda', 0)
from bio_image_unet_b200.dist import DistContext
args = argparse.Namespace(config=int(sys.argv[1]), frames=None, chunk_frames=8, add_patch=0, precision=None, vol=None)
wl = bench.WORKLOADS[args.config](args, dev, DistContext(False)).setup()
for _ in range(3): wl.step_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): wl.step_device()
e1.record(); torch.cuda.synchronize()
print('config', args.config, 'device step %.2f ms' % (e0.elapsed_time(e1) / 3))
eng = wl.engine()
eng.set_profile(True)
wl.step_device(); torch.cuda.synchronize()
kinds, ms = eng.read_profile()
eng.set_profile(False)
print('ops', len(kinds), 'sum %.3f ms' % sum(ms))
print(' '.join(f'{k}:{m:.3f}' for k, m in zip(kinds, ms)))
