"""One benchmark step of a BASELINE config for ncu (profiles/): the workload objects of bench.py, three untimed warm-up
steps, then ONE step between cudaProfilerStart / cudaProfilerStop (run ncu with --profile-from-start off).

    python tools/profile_step.py --config 2 [--frames 8]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from bio_image_unet_b200 import _lib  # noqa: E402
from bio_image_unet_b200.dist import DistContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2)
    ap.add_argument('--frames', type=int, default=None)
    ap.add_argument('--chunk-frames', type=int, default=8)
    ap.add_argument('--add-patch', type=int, default=0)
    ap.add_argument('--precision', default=None)
    ap.add_argument('--vol', default=None, help='cfg 4: volume extents z,x,y (default: the full 256,1024,1024)')
    args = ap.parse_args()
    torch.cuda.set_device(0)
    lib = _lib.load()
    wl = bench.WORKLOADS[args.config](args, torch.device('cuda', 0), DistContext(False)).setup()
    for _ in range(3):
        wl.step_device()
    torch.cuda.synchronize()
    n0 = lib.biu_launch_count()
    torch.cuda.profiler.start()
    wl.step_device()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f'[profile_step] config {args.config}: {lib.biu_launch_count() - n0} kernel launches in the profiled step')


if __name__ == '__main__':
    main()
