#!/usr/bin/env python
"""Forward-only throughput of the other BASELINE configs' networks (cfg 3 Siam, cfg 4 UNet3D, cfg 5 MO-3D) through
the C-ABI engine: tile pixels / voxels per second, TFLOP/s against the reference layer FLOP counts
and the per-op CUDA-event breakdown. Development / evidence tool; the headline contract lives in bench.py.

    python tools/bench_families.py [siam] [unet3d] [unet3d64] [mo3d] [attunet] [unet] [unet_tf32] [nested]
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bio_image_unet_b200 import _lib  # noqa: E402
from bio_image_unet_b200.engine import Engine  # noqa: E402

NAMES = {0: 'first_conv', 1: 'conv', 2: 'conv+head', 3: 'up', 4: 'pool', 5: 'up_nearest', 6: 'max_join', 7: 'gate',
         8: 'mul_psi', 9: 'xcorr', 10: 'up_trilinear', 11: 'up_bilinear'}


def run(name, kind, module, spec, tile, batch, flop_per_px, precision='bf16', in_float=False, reps=5, prev=False):
    torch.manual_seed(0)
    sd = module.state_dict()
    eng = Engine(kind, sd, precision=precision, device='cuda:0', **spec)
    eng.plan(batch, tile)
    shape = (batch, 1, *tile)
    if in_float:
        x = torch.rand(shape, device='cuda')
    else:
        x = torch.randint(0, 256, shape, dtype=torch.uint8, device='cuda')
    xp = torch.randint(0, 256, shape, dtype=torch.uint8, device='cuda') if prev else None
    lib = _lib.load()
    for _ in range(3):
        eng.forward(x, xp, want_val=in_float, want_u8=not in_float)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.forward(x, xp, want_val=in_float, want_u8=not in_float)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.biu_net_set_profile(eng.handle, 1)
    eng.forward(x, xp, want_val=in_float, want_u8=not in_float)
    torch.cuda.synchronize()
    kinds, opms, n_ops = (ctypes.c_int * 96)(), (ctypes.c_float * 96)(), ctypes.c_int(0)
    _lib.check(lib.biu_net_profile_read(eng.handle, 96, kinds, opms, ctypes.byref(n_ops)))
    px = batch
    for t in tile:
        px *= t
    ops = [f'{NAMES.get(kinds[i] % 16, "?")}{"(cc)" if 16 <= kinds[i] < 32 else ("(fused)" if kinds[i] >= 32 else "")}:{opms[i]:.3f}'
           for i in range(n_ops.value)]
    out = {'family': name, 'precision': precision, 'tile': list(tile), 'batch': batch, 'ms_per_forward': ms,
           'tile_mpx_per_s': px / ms / 1e3, 'tflops': flop_per_px * px / ms / 1e9,
           'cuda_core_fallback_ops': sum(1 for i in range(n_ops.value) if 16 <= kinds[i] < 32), 'ops_ms': ops}
    print(json.dumps(out), flush=True)
    eng.close()


def main():
    which = sys.argv[1:] or ['siam', 'unet3d', 'unet3d64', 'mo3d', 'attunet', 'unet_tf32']
    from bio_image_unet_b200.multi_output_unet3d import MultiOutputUnet3D
    from bio_image_unet_b200.siam_unet import Siam_UNet
    from bio_image_unet_b200.unet import AttentionUnet, Unet
    from bio_image_unet_b200.unet3d import UNet3D
    if 'siam' in which:      # cfg 3: Siam_UNet(32, concat), 512x512 tile pairs
        run('siam_concat_nf32', 'siam2d', Siam_UNet(32, 'concat'), dict(n_filter=32, in_channels=1, heads=[('', 1, 'sigmoid')],
            siam_mode='concat'), (512, 512), 72, 478400, prev=True)
    if 'unet3d' in which:    # cfg 4: UNet3D(n_filter=16), 64x128x128 patches
        run('unet3d_nf16', 'unet3d', UNet3D(n_filter=16), dict(n_filter=16, in_channels=1, heads=[('', 1, 'sigmoid')]),
            (64, 128, 128), 32, 109872)
    if 'unet3d64' in which:  # cfg 4, Trainer default width
        run('unet3d_nf64', 'unet3d', UNet3D(n_filter=64), dict(n_filter=64, in_channels=1, heads=[('', 1, 'sigmoid')]),
            (64, 128, 128), 8, 1752576)
    if 'mo3d' in which:      # cfg 5: MultiOutputUnet3D(nf=16, interp), three 1-channel sigmoid heads, 64x256x256 patches
        heads = {f'h{i}': {'channels': 1, 'activation': 'sigmoid'} for i in range(3)}
        run('mo3d_nf16_interp', 'mo3d', MultiOutputUnet3D(1, heads, 16, True),
            dict(n_filter=16, in_channels=1, heads=[(k, 1, 'sigmoid') for k in heads], use_interpolation=True),
            (64, 256, 256), 8, 203088, in_float=True)
    if 'attunet' in which:
        run('attention_unet_nf32', 'attunet2d', AttentionUnet(n_filter=32), dict(n_filter=32, in_channels=1,
            heads=[('', 1, 'sigmoid')]), (512, 512), 100, 367232 + 2 * (64 * 16 + 128 * 32 / 4 + 256 * 64 / 16 + 512 * 128 / 64))
    if 'unet' in which:      # cfg 2 network, forward only
        run('unet_nf32', 'unet2d', Unet(n_filter=32), dict(n_filter=32, in_channels=1, heads=[('', 1, 'sigmoid')]),
            (512, 512), 200, 367232)
    if 'nested' in which:    # multi_output_unet.Predict's default network: U-Net++ (4 pools), n_filter 32, float32 patches
        from bio_image_unet_b200.multi_output_unet import MultiOutputNestedUNet
        nf, depth, flop = 32, 4, 0.0
        for l in range(depth + 1):
            c = nf << l
            for j in range(depth - l + 1):
                cin = (1 if l == 0 else c // 2) if j == 0 else (j + 2) * c
                flop += 2 * 9 * (cin * c + c * c) / 4 ** l
        flop += 2 * nf * 1
        heads = {'seg': {'channels': 1, 'activation': 'sigmoid'}}
        run('nested_unet_nf32', 'nested2d', MultiOutputNestedUNet(1, heads, nf),
            dict(n_filter=nf, in_channels=1, heads=[('seg', 1, 'sigmoid')]), (512, 512), 48, flop, in_float=True)
    if 'unet_tf32' in which:
        run('unet_nf32', 'unet2d', Unet(n_filter=32), dict(n_filter=32, in_channels=1, heads=[('', 1, 'sigmoid')]),
            (512, 512), 100, 367232, precision='tf32')


if __name__ == '__main__':
    main()
