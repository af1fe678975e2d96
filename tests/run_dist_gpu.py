"""torchrun worker: every predictor with distributed=True must reproduce the single-process result exactly.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_dist_gpu.py OUTDIR
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bio_image_unet_b200 import tiff  # noqa: E402
from tests import _golden  # noqa: E402


def main(out_dir):
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    world = int(os.environ['WORLD_SIZE'])
    if torch.cuda.device_count() >= world:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:       # 1-GPU box: both ranks share cuda:0, collectives go through gloo (NCCL refuses two ranks on a device)
        local = local % torch.cuda.device_count()
        torch.cuda.set_device(local)
        dist.init_process_group('gloo')
    print(f'[dist] rank {rank}: cuda:{local}, backend {dist.get_backend()}', flush=True)
    from bio_image_unet_b200.multi_output_unet3d import Predict as PredictMO
    from bio_image_unet_b200.siam_unet import Predict as PredictSiam
    from bio_image_unet_b200.unet import Predict as PredictUnet
    from bio_image_unet_b200.unet3d import Predict as Predict3D
    ok = True
    # ---- unet, modes 'single' and 'all' (histogram all-reduce) ----
    for name in ('unet_single', 'unet_all_invert'):
        g = _golden.load(name)
        ckpt = os.path.join(out_dir, f'{name}_{rank}.pt')
        torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1, 'out_channels': 1}, ckpt)
        kw = dict(resize_dim=tuple(int(v) for v in g['resize_dim']), invert=bool(g['invert']), normalization_mode=str(g['mode']),
                  clip_threshold=tuple(g['clip']), add_tile=int(g['add_tile']), show_progress=False, precision='fp32')
        res_d = os.path.join(out_dir, f'{name}_dist.tif')
        PredictUnet(g['imgs'].copy(), res_d, ckpt, distributed=True, **kw)
        dist.barrier()
        if rank == 0:
            res_s = os.path.join(out_dir, f'{name}_single.tif')
            PredictUnet(g['imgs'].copy(), res_s, ckpt, device=f'cuda:{local}', **kw)
            same = np.array_equal(tiff.imread(res_d), tiff.imread(res_s))
            print(f'[dist] {name}: distributed == single: {same}')
            ok &= same
    # ---- siam ----
    g = _golden.load('siam_concat')
    ckpt = os.path.join(out_dir, f'siam_{rank}.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'mode': 'concat'}, ckpt)
    kw = dict(resize_dim=tuple(int(v) for v in g['resize_dim']), normalization_mode='single', clip_threshold=tuple(g['clip']),
              add_tile=int(g['add_tile']), show_progress=False, precision='fp32')
    PredictSiam(g['movie'].copy(), os.path.join(out_dir, 'siam_dist.tif'), ckpt, distributed=True, **kw)
    dist.barrier()
    if rank == 0:
        out = tiff.imread(os.path.join(out_dir, 'siam_dist.tif'))
        same = np.abs(out.astype(np.int16) - g['result'].astype(np.int16)).max() <= 1
        print(f'[dist] siam_concat: distributed within 1 LSB of the reference golden: {same}')
        ok &= bool(same)
    # ---- unet3d (patch list sharded, global histogram all-reduced, mod-3 stitch on rank 0) ----
    g = _golden.load('unet3d_overlap')
    ckpt = os.path.join(out_dir, f'u3d_{rank}.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1, 'out_channels': 1}, ckpt)
    Predict3D(g['vol'].copy(), os.path.join(out_dir, 'u3d_dist.tif'), ckpt, resize_dim=tuple(int(v) for v in g['resize_dim']),
              clip_threshold=tuple(g['clip']), add_patch=int(g['add_patch']), progress_bar=False, precision='fp32', distributed=True)
    dist.barrier()
    if rank == 0:
        out = tiff.imread(os.path.join(out_dir, 'u3d_dist.tif'))
        same = np.abs(out.astype(np.float32) - g['result_file'].astype(np.float32)).max() <= 1
        print(f'[dist] unet3d_overlap: distributed within 1 LSB of the reference golden: {same}')
        ok &= bool(same)
    # ---- unet3d, several z-rows per rank with overlap: boundary patches travel between ranks, every rank stitches
    #      its own planes; must equal the single-process run bit for bit ----
    from bio_image_unet_b200.unet3d import UNet3D
    torch.manual_seed(5)
    sd3 = UNet3D(n_filter=4).state_dict()
    ckpt = os.path.join(out_dir, f'u3d_syn_{rank}.pt')
    torch.save({'state_dict': sd3, 'n_filter': 4, 'in_channels': 1, 'out_channels': 1}, ckpt)
    vol = np.random.default_rng(11).integers(0, 3000, (44, 40, 48)).astype('uint16')
    for add_patch in (1, 0):
        kw = dict(resize_dim=(8, 16, 16), add_patch=add_patch, progress_bar=False, precision='fp32')
        res_d = os.path.join(out_dir, f'u3d_syn{add_patch}_dist.tif')
        p3 = Predict3D(vol.copy(), res_d, ckpt, distributed=True, **kw)
        dist.barrier()
        if rank == 0:
            res_s = os.path.join(out_dir, f'u3d_syn{add_patch}_single.tif')
            Predict3D(vol.copy(), res_s, ckpt, device=f'cuda:{local}', **kw)
            same = np.array_equal(tiff.imread(res_d), tiff.imread(res_s))
            print(f'[dist] unet3d synthetic add_patch={add_patch} (N_z={p3.N_z}): z-slab sharded == single: {same}')
            ok &= bool(same)
    # ---- multi-output 3D (volumes sharded) ----
    g = _golden.load('mo3d_interp')
    ckpt = os.path.join(out_dir, f'mo_{rank}.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1,
                'output_heads': _golden.MO3D_HEADS, 'use_interpolation': True}, ckpt)
    p = PredictMO(g['imgs'].copy(), ckpt, max_patch_size=tuple(int(v) for v in g['max_patch']), overlap_factor=float(g['overlap']),
                  normalization_mode=str(g['norm_mode']), clip_threshold=tuple(g['clip']), show_progress=False, precision='fp32',
                  distributed=True)
    if rank == 0:
        ref = _golden.sub(g, 'result')
        same = all(np.abs(p.result[k] - ref[k]).max() < 2e-4 for k in ref)
        print(f'[dist] mo3d_interp: distributed within 2e-4 of the reference golden: {same}')
        ok &= bool(same)
    else:
        assert p.result is None
    flag = torch.tensor([1 if ok else 0], device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        open(os.path.join(out_dir, 'DIST_OK' if ok else 'DIST_FAIL'), 'w').write('done')
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == '__main__':
    main(sys.argv[1])
