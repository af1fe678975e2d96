"""Generate the golden fixtures by executing the UNMODIFIED reference (authoring container only).

    python tests/golden/make_golden.py

Imports bio_image_unet from /root/reference (recipe: oracle/ref_import.py), builds small seeded models with a
"stress" initialisation (Kaiming conv weights, randomised BatchNorm statistics), runs the reference's Predict
classes on CPU on small synthetic stacks and stores inputs, weights and every intermediate the reference
produces (tile starts, uint8 tiles, uint8 result tiles, stitched result) as tests/golden/*.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

ref = ref_import.import_reference()
from bio_image_unet.unet.predict import Predict as UnetPredict  # noqa: E402
from bio_image_unet.unet.unet import Unet  # noqa: E402
from bio_image_unet.unet.attention_unet import AttentionUnet  # noqa: E402
from bio_image_unet.unet.unet_v0 import Unet_v0  # noqa: E402
from bio_image_unet.siam_unet.predict import Predict as SiamPredict  # noqa: E402
from bio_image_unet.siam_unet.siam_unet import Siam_UNet  # noqa: E402
from bio_image_unet.unet3d.predict import Predict as Unet3dPredict  # noqa: E402
from bio_image_unet.unet3d.unet3d import UNet3D  # noqa: E402
from bio_image_unet.multi_output_unet3d.predict import Predict as Mo3dPredict  # noqa: E402
from bio_image_unet.multi_output_unet3d.multi_output_unet3d import MultiOutputUnet3D  # noqa: E402
from bio_image_unet.multi_output_unet.predict import Predict as Mo2dPredict  # noqa: E402
from bio_image_unet.multi_output_unet.multi_output_unet import MultiOutputUnet  # noqa: E402
from bio_image_unet.multi_output_unet.multi_output_nested_unet import (MultiOutputNestedUNet,  # noqa: E402
                                                                        MultiOutputNestedUNet_3Levels)


def stress_init(model, seed, head_gain=4.0):
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.ConvTranspose2d, torch.nn.ConvTranspose3d)):
            fan_in = m.weight[0].numel() if not isinstance(m, (torch.nn.ConvTranspose2d, torch.nn.ConvTranspose3d)) \
                else m.weight.shape[0]
            std = (2.0 / (1 + 0.1 ** 2) / fan_in) ** 0.5
            m.weight.data = torch.randn(m.weight.shape, generator=g) * std
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
            if name.startswith('final') or name.startswith('output_layers'):
                m.weight.data *= head_gain
        elif isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.1
            m.running_var.data = torch.rand(m.running_var.shape, generator=g) + 0.5
    return model


def blobs(shape, seed, dtype='uint16', peak=3000):
    """Sum of Gaussian blobs + Poisson-like noise, so percentiles / clipping are non-degenerate."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(s, dtype='float64') for s in shape[-2:]], indexing='ij')
    out = np.zeros(shape, dtype='float64')
    flat = out.reshape(-1, *shape[-2:])
    for f in range(flat.shape[0]):
        for _ in range(6):
            cy, cx = rng.uniform(0, shape[-2]), rng.uniform(0, shape[-1])
            s = rng.uniform(3, 9)
            flat[f] += rng.uniform(0.3, 1.0) * np.exp(-((grids[0] - cy) ** 2 + (grids[1] - cx) ** 2) / (2 * s * s))
    out = out * peak + rng.uniform(80, 120) + rng.normal(0, 12, shape)
    if np.dtype(dtype).kind == 'f':            # float stacks: non-integer values, also negative ones
        return ((out - 150.0) / 37.0).astype(dtype)
    out = np.clip(out, 0, np.iinfo(dtype).max)
    return out.astype(dtype)


def sd_arrays(model):
    return {'sd/' + k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


class Capture:
    """Wrap the name-mangled private methods of a reference Predict class to record what they return."""

    def __init__(self, cls, names):
        self.cls, self.names, self.saved, self.out = cls, names, {}, {}

    def __enter__(self):
        for n in self.names:
            attr = f'_Predict__{n}'
            orig = getattr(self.cls, attr)
            self.saved[attr] = orig

            def make(orig=orig, n=n):
                def wrapper(inner_self, *a, **k):
                    r = orig(inner_self, *a, **k)
                    self.out.setdefault(n, []).append(np.array(r) if isinstance(r, np.ndarray) else r)
                    return r
                return wrapper
            setattr(self.cls, attr, make())
        return self

    def __exit__(self, *exc):
        for attr, orig in self.saved.items():
            setattr(self.cls, attr, orig)
        return False


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB')


def gen_unet(name, shape, resize_dim, add_tile, mode, invert, seed, dtype='uint16', nf=4, network='Unet'):
    torch.manual_seed(seed)
    model = stress_init({'Unet': Unet, 'AttentionUnet': AttentionUnet, 'Unet_v0': Unet_v0}[network](n_filter=nf), seed)
    ckpt = f'/tmp/golden_{name}.pt'
    params = {'state_dict': model.state_dict(), 'n_filter': nf, 'in_channels': 1, 'out_channels': 1}
    if network == 'Unet_v0':          # old checkpoints carry no channel counts (unet/predict.py:94-97)
        del params['in_channels'], params['out_channels']
    torch.save(params, ckpt)
    imgs = blobs(shape, seed, dtype)
    original = imgs.copy()
    with Capture(UnetPredict, ['split', 'predict', 'stitch']) as cap:
        p = UnetPredict(imgs, 'res_' + name, ckpt, network=network, resize_dim=resize_dim, invert=invert,
                        normalization_mode=mode, clip_threshold=(0., 99.8), add_tile=add_tile, show_progress=False,
                        device='cpu')
    result_file = ref_import.TIFF_STORE['res_' + name]
    save(name, imgs=original, imgs_after=imgs, resize_dim=np.array(resize_dim), add_tile=add_tile,
         invert=int(invert), mode=mode, clip=np.array([0., 99.8]), n_filter=nf, network=network,
         N_x=p.N_x, N_y=p.N_y, X_start=p.X_start, Y_start=p.Y_start, patches=cap.out['split'][0],
         result_patches=cap.out['predict'][0], result=cap.out['stitch'][0], result_file=result_file,
         **sd_arrays(model))


def gen_siam(name, shape, resize_dim, add_tile, siam_mode, norm_mode, seed, nf=4):
    torch.manual_seed(seed)
    model = stress_init(Siam_UNet(n_filter=nf, mode=siam_mode), seed)
    ckpt = f'/tmp/golden_{name}.pt'
    torch.save({'state_dict': model.state_dict(), 'n_filter': nf, 'mode': siam_mode}, ckpt)
    movie = blobs(shape, seed)
    ref_import.TIFF_STORE['movie_' + name] = movie.copy()
    with Capture(SiamPredict, ['split', 'predict']) as cap:
        p = SiamPredict('movie_' + name, 'res_' + name, ckpt, resize_dim=resize_dim, normalization_mode=norm_mode,
                        clip_threshold=(0.0, 99.98), add_tile=add_tile, show_progress=False, device='cpu')
    result = ref_import.TIFF_STORE['res_' + name]
    save(name, movie=movie, resize_dim=np.array(resize_dim), add_tile=add_tile, siam_mode=siam_mode,
         norm_mode=norm_mode, clip=np.array([0., 99.98]), n_filter=nf, N_x=p.N_x, N_y=p.N_y, X_start=p.X_start,
         Y_start=p.Y_start, patches=np.stack(cap.out['split']), result_patches=np.stack(cap.out['predict']),
         result=result, **sd_arrays(model))
    import shutil
    shutil.rmtree('temp_movie_' + name, ignore_errors=True)


def gen_unet3d(name, shape, resize_dim, add_patch, seed, nf=4, dtype='uint16', interp=False):
    torch.manual_seed(seed)
    model = stress_init(UNet3D(n_filter=nf, use_interpolation=interp), seed)
    ckpt = f'/tmp/golden_{name}.pt'
    params = {'state_dict': model.state_dict(), 'n_filter': nf, 'in_channels': 1, 'out_channels': 1}
    if interp:                # unet3d/predict.py:82-86 reads the flag from the checkpoint
        params['use_interpolation'] = True
    torch.save(params, ckpt)
    vol = blobs(shape, seed, dtype)
    with Capture(Unet3dPredict, ['split', 'predict', 'stitch']) as cap:
        p = Unet3dPredict(vol.copy(), 'res_' + name, ckpt, resize_dim=resize_dim, clip_threshold=(0., 99.8),
                          add_patch=add_patch, progress_bar=False, device='cpu')
    save(name, vol=vol, resize_dim=np.array(resize_dim), add_patch=add_patch, clip=np.array([0., 99.8]), n_filter=nf,
         interp=int(interp), N_z=p.N_z, N_x=p.N_x, N_y=p.N_y, Z_start=p.Z_start, X_start=p.X_start, Y_start=p.Y_start,
         patches=cap.out['split'][0], result_patches=cap.out['predict'][0], result=cap.out['stitch'][0],
         result_file=ref_import.TIFF_STORE['res_' + name], **sd_arrays(model))


def gen_mo3d(name, shape, max_patch, overlap, norm_mode, interp, seed, nf=4, batch_size=2):
    torch.manual_seed(seed)
    heads = {'seg': {'channels': 1, 'activation': 'sigmoid'}, 'flow': {'channels': 2, 'activation': None},
             'dist': {'channels': 1, 'activation': 'relu'}}
    model = stress_init(MultiOutputUnet3D(1, heads, nf, interp), seed, head_gain=2.0)
    ckpt = f'/tmp/golden_{name}.pt'
    torch.save({'state_dict': model.state_dict(), 'n_filter': nf, 'in_channels': 1, 'output_heads': heads,
                'use_interpolation': interp}, ckpt)
    imgs = blobs(shape, seed)
    with Capture(Mo3dPredict, ['split', 'predict']) as cap:
        p = Mo3dPredict(imgs.copy(), ckpt, result_path=None, max_patch_size=max_patch, overlap_factor=overlap,
                        batch_size=batch_size, normalization_mode=norm_mode, clip_threshold=(0., 99.98),
                        show_progress=False, device='cpu')
    arrays = {f'result/{k}': v for k, v in p.result.items()}
    arrays.update({f'result_patches/{k}': v for k, v in cap.out['predict'][0].items()})
    save(name, imgs=imgs, max_patch=np.array(max_patch), overlap=overlap, norm_mode=norm_mode, interp=int(interp),
         clip=np.array([0., 99.98]), n_filter=nf, patch_size=np.array(p.patch_size), Z_start=np.array(p.Z_start),
         Y_start=np.array(p.Y_start), X_start=np.array(p.X_start), patches=cap.out['split'][0], **arrays,
         **sd_arrays(model))


def gen_mo2d(name, shape, max_patch, add_tile, norm_mode, seed, nf=4, batch_size=2, dtype='uint16',
             network=MultiOutputUnet, deep_supervision=False):
    torch.manual_seed(seed)
    heads = {'seg': {'channels': 1, 'activation': 'sigmoid'}, 'vec': {'channels': 2, 'activation': None},
             'dist': {'channels': 1, 'activation': 'relu'}}
    if network is MultiOutputUnet:
        model = stress_init(MultiOutputUnet(1, heads, nf), seed, head_gain=2.0)
    else:
        model = stress_init(network(1, heads, nf, deep_supervision=deep_supervision), seed, head_gain=2.0)
    ckpt = f'/tmp/golden_{name}.pt'
    torch.save({'state_dict': model.state_dict(), 'n_filter': nf, 'in_channels': 1, 'output_heads': heads,
                'deep_supervision': deep_supervision}, ckpt)
    imgs = blobs(shape, seed, dtype)
    with Capture(Mo2dPredict, ['preprocess', 'split', 'predict']) as cap:
        p = Mo2dPredict(imgs.copy(), ckpt, result_path=None, network=network, max_patch_size=max_patch,
                        batch_size=batch_size, normalization_mode=norm_mode, clip_threshold=(0., 99.98),
                        add_tile=add_tile, show_progress=False, device='cpu')
    arrays = {f'result/{k}': v for k, v in p.result.items()}
    arrays.update({f'result_patches/{k}': v for k, v in cap.out['predict'][0].items()})
    save(name, imgs=imgs, max_patch=np.array(max_patch), add_tile=add_tile, norm_mode=norm_mode,
         clip=np.array([0., 99.98]), n_filter=nf, patch_size=np.array(p.patch_size), N_x=p.N_x, N_y=p.N_y,
         X_start=p.X_start, Y_start=p.Y_start, norm=cap.out['preprocess'][0], patches=cap.out['split'][0],
         network=network.__name__, deep_supervision=int(deep_supervision), **arrays, **sd_arrays(model))


if __name__ == '__main__':
    torch.set_num_threads(4)
    only = sys.argv[1:]            # optional: names of the fixtures to (re)generate
    if only:
        _save = save

        def save(name, **arrays):  # noqa: F811
            if name in only:
                _save(name, **arrays)
    gen_unet('attunet_single', (2, 70, 90), (32, 48), 1, 'single', False, seed=15, network='AttentionUnet', nf=8)
    gen_unet('unetv0_all', (2, 64, 80), (32, 32), 1, 'all', False, seed=16, network='Unet_v0')
    gen_unet('unet_f32_single', (2, 70, 90), (32, 48), 1, 'single', False, seed=17, dtype='float32')
    gen_unet('unet_f32_first_invert', (3, 64, 64), (32, 32), 1, 'first', True, seed=18, dtype='float32')
    gen_unet('unet_f32_all', (3, 48, 80), (32, 32), 0, 'all', False, seed=19, dtype='float32')
    gen_siam('siam_corr', (3, 48, 64), (32, 48), 1, 'corr', 'single', seed=24)
    gen_unet3d('unet3d_trilinear', (12, 40, 40), (8, 16, 24), 1, seed=33, interp=True)
    gen_mo2d('mo2d_single_overlap', (2, 100, 150), (64, 80), 1, 'single', seed=51)
    gen_mo2d('mo2d_all_pad', (3, 40, 70), (64, 64), 0, 'all', seed=52, dtype='uint8')
    gen_mo2d('mo2d_first_holes', (2, 96, 96), (48, 48), 2, 'first', seed=53)
    gen_mo2d('nested_single_overlap', (2, 100, 150), (64, 80), 1, 'single', seed=54, network=MultiOutputNestedUNet)
    gen_mo2d('nested3l_ds_all', (2, 72, 90), (48, 64), 1, 'all', seed=55, network=MultiOutputNestedUNet_3Levels,
             deep_supervision=True, dtype='uint8')
    if only:
        sys.exit(0)
    gen_unet('unet_single', (2, 70, 90), (32, 48), 1, 'single', False, seed=11)
    gen_unet('unet_all_invert', (3, 64, 64), (32, 32), 1, 'all', True, seed=12)
    gen_unet('unet_first_u8', (2, 48, 80), (32, 32), 0, 'first', False, seed=13, dtype='uint8')
    gen_unet('unet_small_reflect', (1, 20, 100), (32, 48), 0, 'single', False, seed=14)
    gen_siam('siam_concat', (3, 40, 56), (32, 32), 1, 'concat', 'single', seed=21)
    gen_siam('siam_max', (2, 48, 48), (32, 32), 0, 'max', 'all', seed=22)
    gen_siam('siam_control_small', (2, 24, 40), (32, 32), 0, 'control', 'first', seed=23)
    gen_unet3d('unet3d_overlap', (12, 40, 40), (8, 16, 16), 1, seed=31)
    gen_unet3d('unet3d_disjoint', (16, 32, 32), (8, 16, 16), 0, seed=32, dtype='uint8')
    gen_mo3d('mo3d_interp', (2, 12, 40, 40), (8, 32, 32), 0.25, 'all', True, seed=41)
    gen_mo3d('mo3d_convt', (1, 20, 24, 24), (8, 16, 16), 0.1, 'first', False, seed=42)
