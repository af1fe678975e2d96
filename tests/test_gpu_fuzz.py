"""Seeded shape fuzz over the kernel variants: for random tile shapes / batches / widths the default engine (row kernel
incl. plane mode and K split, CTA pairs) must agree with the plain halo-tile kernel on single CTAs - bit for bit where
only the CTA pairing differs, to rounding noise of the stored format where the tap folding changes the summation order
- and stay within the reference-arithmetic yardstick of the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import models as omodels
from tests import _parity

pytestmark = pytest.mark.gpu


def _engine_out(kind, sd, nf, x, precision, heads=(('', 1, 'sigmoid'),), **kw):
    from bio_image_unet_b200.engine import Engine
    eng = Engine(kind, sd, nf, 1, list(heads), precision=precision, device='cuda:0', **kw)
    eng.plan(x.shape[0], tuple(x.shape[2:]))
    val, _ = eng.forward(x.cuda(), want_val=True, want_u8=False)
    torch.cuda.synchronize()
    out = val.cpu()
    fb = eng.fallback_ops
    eng.close()
    return out, fb


def _randomised_bn(sd, seed):
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k.endswith('running_var') or k.endswith('.1.weight'):
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean') or k.endswith('.1.bias'):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
    return sd


@pytest.mark.parametrize('seed', range(8))
def test_fuzz_unet2d_variants(seed):
    from bio_image_unet_b200 import _lib
    from bio_image_unet_b200.unet import Unet
    lib = _lib.load()
    rng = np.random.default_rng(1000 + seed)
    nf = int(rng.choice([16, 32]))
    tile = (16 * int(rng.integers(1, 12)), 16 * int(rng.integers(1, 20)))
    batch = int(rng.integers(1, 6))
    precision = str(rng.choice(['bf16', 'tf32']))
    torch.manual_seed(seed)
    sd = _randomised_bn(Unet(n_filter=nf).state_dict(), seed)
    x = torch.randint(0, 256, (batch, 1, *tile), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
    try:
        base, fb = _engine_out('unet2d', sd, nf, x, precision)
        lib.biu_set_halo_cta2(0)
        single, _ = _engine_out('unet2d', sd, nf, x, precision)
        lib.biu_set_rows_kernel(0)
        halo, _ = _engine_out('unet2d', sd, nf, x, precision)
    finally:
        lib.biu_set_halo_cta2(1)
        lib.biu_set_rows_kernel(1)
    assert fb == 0
    assert torch.equal(base, single), (nf, tile, batch, precision)                 # CTA pairs: same MMAs, same order
    xf = x.float() / 255
    with torch.no_grad():
        ref = omodels.unet_forward(sd, xf)[0]
    _, _, tol = _parity.check(base, ref, omodels.unet_forward, precision, sd, xf, what=f'fuzz {nf} {tile} {batch}')
    assert (base - halo).abs().max().item() <= tol, (nf, tile, batch, precision)   # folded taps vs per-tap halo kernel


@pytest.mark.parametrize('seed', range(6))
def test_fuzz_unet3d_variants(seed):
    from bio_image_unet_b200 import _lib
    from bio_image_unet_b200.unet3d import UNet3D
    lib = _lib.load()
    rng = np.random.default_rng(2000 + seed)
    nf = int(rng.choice([16, 16, 32]))
    tile = (8 * int(rng.integers(1, 5)), 8 * int(rng.integers(2, 9)), 8 * int(rng.integers(1, 19)))
    batch = int(rng.integers(1, 4))
    precision = str(rng.choice(['bf16', 'tf32']))
    torch.manual_seed(seed)
    sd = _randomised_bn(UNet3D(n_filter=nf).state_dict(), seed)
    x = torch.randint(0, 256, (batch, 1, *tile), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
    fwd = lambda sd_, x_: omodels.unet3d_forward(sd_, x_)                                  # noqa: E731
    try:
        base, fb = _engine_out('unet3d', sd, nf, x, precision)
        lib.biu_set_halo_cta2(0)
        single, _ = _engine_out('unet3d', sd, nf, x, precision)
        lib.biu_set_rows_kernel(0)
        halo, _ = _engine_out('unet3d', sd, nf, x, precision)
    finally:
        lib.biu_set_halo_cta2(1)
        lib.biu_set_rows_kernel(1)
    assert torch.equal(base, single), (nf, tile, batch, precision)
    xf = x.float() / 255
    with torch.no_grad():
        ref = fwd(sd, xf)[0]
    _, _, tol = _parity.check(base, ref, fwd, precision, sd, xf, what=f'fuzz3d {nf} {tile} {batch}')
    assert (base - halo).abs().max().item() <= tol, (nf, tile, batch, precision)
    del fb
