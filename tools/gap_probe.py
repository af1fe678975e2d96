#!/usr/bin/env python
"""Where does the forward's time go between the kernels? Times N forwards of the cfg-2 network with and without the
per-op CUDA events and prints the per-op sum beside the whole-forward time (development tool)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bio_image_unet_b200.engine import Engine  # noqa: E402
from bio_image_unet_b200.unet import Unet  # noqa: E402


def timed(eng, x, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng.forward(x, None, want_val=False, want_u8=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.manual_seed(0)
    eng = Engine('unet2d', Unet(n_filter=32).state_dict(), 32, 1, [('', 1, 'sigmoid')], precision='bf16', device='cuda:0')
    eng.plan(batch, (512, 512))
    x = torch.randint(0, 256, (batch, 1, 512, 512), dtype=torch.uint8, device='cuda')
    for _ in range(5):
        eng.forward(x, None, want_val=False, want_u8=True)
    for rnd in range(3):
        eng.set_profile(False)
        plain = timed(eng, x, 10)
        eng.set_profile(True)
        prof = timed(eng, x, 10)
        kinds, ms = eng.read_profile()
        print(f'round {rnd}: plain {plain:.3f} ms/forward, with per-op events {prof:.3f} ms/forward, '
              f'sum of the last forward\'s {len(ms)} op events {sum(ms):.3f} ms', flush=True)
    print('ops:', ' '.join(f'{k}:{m:.3f}' for k, m in zip(kinds, ms)))


if __name__ == '__main__':
    main()
