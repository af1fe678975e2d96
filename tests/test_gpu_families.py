"""GPU parity of the Siamese, 3D and multi-output 3D predictors against the golden fixtures produced by the
unmodified reference, plus bit-exactness of the 3D / ramp stitch kernels against the CPU oracle.

Integer stages (tile indices, uint8 / float32 tiles, stitches) are bit-exact. The network forward is compared in
the exact-fp32 mode at +-1 LSB of the reference's uint8 quantisation (float heads: 1e-4 of the head's range), and in
the tf32 / bf16 tensor-core modes against 1.5x the error of the reference's own arithmetic at that precision on the
fixture's weights and tiles (tests/_parity.py: torch.autocast(bfloat16) / operand-rounded TF32) - no constants.
"""
import numpy as np
import pytest
import torch

from oracle import models as omodels
from oracle import pipeline as opipe
from tests import _golden, _parity

pytestmark = pytest.mark.gpu



def _u8(t):
    return torch.from_numpy(np.ascontiguousarray(t)).float() / 255


@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
@pytest.mark.parametrize('name', ['siam_concat', 'siam_max', 'siam_control_small', 'siam_corr'])
def test_siam_predict_matches_reference_golden(name, precision, tmp_path):
    from bio_image_unet_b200 import tiff
    from bio_image_unet_b200.siam_unet import Predict
    g = _golden.load(name)
    ckpt = str(tmp_path / 'model.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'mode': str(g['siam_mode'])}, ckpt)
    movie_file = str(tmp_path / 'movie.tif')
    tiff.imwrite(movie_file, g['movie'])
    res_file = str(tmp_path / 'res.tif')
    p = Predict(movie_file, res_file, ckpt, resize_dim=tuple(int(v) for v in g['resize_dim']),
                normalization_mode=str(g['norm_mode']), clip_threshold=tuple(g['clip']), add_tile=int(g['add_tile']),
                show_progress=False, device='cuda:0', precision=precision, keep_intermediates=True)
    assert (p.N_x, p.N_y) == (int(g['N_x']), int(g['N_y']))
    assert np.array_equal(p.X_start, g['X_start']) and np.array_equal(p.Y_start, g['Y_start'])
    assert np.array_equal(p.patches, g['patches'])                    # (T, N, 2, th, tw): ch0 current, ch1 previous
    mode = str(g['siam_mode'])
    th, tw = (int(v) for v in g['resize_dim'])
    cur, prev = _u8(g['patches'][:, :, 0]).reshape(-1, 1, th, tw), _u8(g['patches'][:, :, 1]).reshape(-1, 1, th, tw)
    fwd = lambda sd_, c_, p_: omodels.siam_forward(sd_, c_, p_, mode)                     # noqa: E731
    sd = _golden.state_dict(g)
    lsb = _parity.lsb_bound(fwd, precision, sd, cur, prev)
    d = np.abs(p.result_patches.astype(np.int16) - g['result_patches'].astype(np.int16))
    assert d.max() <= lsb, (d.max(), lsb)
    if precision != 'fp32':       # the error distribution, not only its tail, has to look like the reference arithmetic's
        _, mean_b = _parity.bound(fwd, precision, sd, cur, prev)
        assert d.mean() <= 255 * mean_b + 0.5, (d.mean(), mean_b)
    out = tiff.imread(res_file)
    assert out.dtype == np.uint8 and out.shape == g['result'].shape
    assert np.abs(out.astype(np.int16) - g['result'].astype(np.int16)).max() <= lsb
    # stitch exact from the engine's own tiles
    grid = (p.N_x, p.N_y, p.X_start, p.Y_start)
    for t in range(out.shape[0]):
        st = opipe.stitch_mean_2d(p.result_patches[t], 1, g['movie'].shape[1:], tuple(int(v) for v in g['resize_dim']), grid)
        assert np.array_equal(out[t], st)


@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
@pytest.mark.parametrize('name', ['unet3d_overlap', 'unet3d_disjoint', 'unet3d_trilinear'])
def test_unet3d_predict_matches_reference_golden(name, precision, tmp_path):
    from bio_image_unet_b200 import tiff
    from bio_image_unet_b200.unet3d import Predict
    g = _golden.load(name)
    ckpt = str(tmp_path / 'model.pt')
    params = {'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1, 'out_channels': 1}
    if 'interp' in g and int(g['interp']):
        params['use_interpolation'] = True          # unet3d/predict.py:82-86 reads the flag from the checkpoint
    torch.save(params, ckpt)
    res_file = str(tmp_path / 'res.tif')
    rd = tuple(int(v) for v in g['resize_dim'])
    p = Predict(g['vol'].copy(), res_file, ckpt, resize_dim=rd, clip_threshold=tuple(g['clip']),
                add_patch=int(g['add_patch']), progress_bar=False, device='cuda:0', precision=precision,
                keep_intermediates=True)
    assert (p.N_z, p.N_x, p.N_y) == (int(g['N_z']), int(g['N_x']), int(g['N_y']))
    for a, b in ((p.Z_start, g['Z_start']), (p.X_start, g['X_start']), (p.Y_start, g['Y_start'])):
        assert np.array_equal(a, b) and a.dtype == np.uint16
    assert np.array_equal(p.patches, g['patches'])
    interp = bool('interp' in g and int(g['interp']))
    fwd = lambda sd_, x_: omodels.unet3d_forward(sd_, x_, interp)                          # noqa: E731
    lsb = _parity.lsb_bound(fwd, precision, _golden.state_dict(g), _u8(g['patches'])[:, None])
    d = np.abs(p.result_patches.astype(np.int16) - g['result_patches'].astype(np.int16))
    assert d.max() <= lsb, (d.max(), lsb)
    out = tiff.imread(res_file)
    assert out.dtype == np.float16
    grid = (p.N_z, p.N_x, p.N_y, p.Z_start, p.X_start, p.Y_start)
    assert np.array_equal(out, opipe.stitch_mod3(p.result_patches, g['vol'].shape, rd, grid).astype('float16'))
    assert np.abs(out.astype(np.float32) - g['result_file'].astype(np.float32)).max() <= lsb


@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
@pytest.mark.parametrize('name', ['mo3d_interp', 'mo3d_convt'])
def test_mo3d_predict_matches_reference_golden(name, precision, tmp_path):
    from bio_image_unet_b200.multi_output_unet3d import Predict
    g = _golden.load(name)
    ckpt = str(tmp_path / 'model.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1,
                'output_heads': _golden.MO3D_HEADS, 'use_interpolation': bool(g['interp'])}, ckpt)
    p = Predict(g['imgs'].copy(), ckpt, result_path=None, max_patch_size=tuple(int(v) for v in g['max_patch']),
                overlap_factor=float(g['overlap']), batch_size=2, normalization_mode=str(g['norm_mode']),
                clip_threshold=tuple(g['clip']), show_progress=False, device='cuda:0', precision=precision,
                keep_intermediates=True)
    assert p.patch_size == tuple(int(v) for v in g['patch_size'])
    assert list(p.Z_start) == list(g['Z_start']) and list(p.Y_start) == list(g['Y_start']) and list(p.X_start) == list(g['X_start'])
    assert np.array_equal(p.patches, g['patches'])          # float32 normalisation + patch gather: bit-exact
    ref = _golden.sub(g, 'result')
    assert list(p.result.keys()) == list(_golden.MO3D_HEADS.keys())
    interp = bool(g['interp'])
    fwd = lambda sd_, x_: omodels.mo3d_forward(sd_, x_, _golden.MO3D_HEADS, interp)        # noqa: E731
    x = torch.from_numpy(g['patches']).reshape(-1, 1, *p.patch_size)
    cb = _parity.per_channel_bound(fwd, precision, _golden.state_dict(g), x)               # heads of different scale
    c0 = 0
    for k, cfg in _golden.MO3D_HEADS.items():
        assert p.result[k].shape == ref[k].shape
        err = np.abs(p.result[k] - ref[k]).max()          # the blend is a convex combination: patch bounds carry over
        bound = float(cb[c0:c0 + cfg['channels']].max()) + 2e-4
        assert err <= bound, (k, err, bound)
        c0 += cfg['channels']


def test_norm_f32_single_mode_bit_exact():
    from bio_image_unet_b200 import engine as E
    rng = np.random.default_rng(8)
    vols = rng.integers(0, 3000, (3, 6, 20, 24)).astype('uint16')
    vols[1] = (rng.normal(500, 60, vols[1].shape).clip(0, 65535)).astype('uint16')
    for mode in ('single', 'first', 'all'):
        for clip in ((0., 99.98), (1., 99.)):
            ref = opipe.mo3d_preprocess(vols.astype('float32'), mode, clip)
            dev = torch.from_numpy(vols).cuda().reshape(3, -1)
            hist = E.histogram(dev)
            if mode == 'single':
                lut, _ = E.norm_lut_f32(hist, hist, 3, clip[0], clip[1], 0)
            else:
                b = E.hist_sum(hist) if mode == 'all' else hist[0:1].contiguous()
                lut, _ = E.norm_lut_f32(b, b, 1, clip[0], clip[1], 1)
            got = E.apply_lut_f32(dev, lut).cpu().numpy().reshape(vols.shape)
            assert np.array_equal(got, ref), (mode, clip, np.abs(got - ref).max())


def test_stitch_mod3_bit_exact():
    from bio_image_unet_b200 import engine as E
    from bio_image_unet_b200 import tiling
    rng = np.random.default_rng(5)
    for vol_shape, rd, add in [((12, 40, 40), (8, 16, 16), 1), ((16, 32, 32), (8, 16, 16), 0), ((8, 24, 24), (8, 16, 16), 1),
                               ((5, 20, 36), (8, 16, 16), 0), ((20, 33, 47), (8, 16, 16), 2)]:
        grid = tiling.grid_3d(vol_shape, rd, add)
        n = grid[0] * grid[1] * grid[2]
        tiles = rng.integers(0, 256, (n, *rd)).astype('uint8')
        ref = opipe.stitch_mod3(tiles, vol_shape, rd, grid)
        got = E.stitch_mod3_u8(torch.from_numpy(tiles).cuda(), vol_shape, grid[3], grid[4], grid[5], rd).cpu().numpy()
        assert np.array_equal(np.squeeze(got), ref), (vol_shape, rd, add)


def test_stitch_ramp_bit_exact():
    from bio_image_unet_b200 import engine as E
    from bio_image_unet_b200 import tiling
    rng = np.random.default_rng(6)
    for shape, max_patch, ov, c in [((2, 12, 40, 40), (8, 32, 32), 0.25, 2), ((1, 20, 24, 24), (8, 16, 16), 0.1, 1),
                                    ((1, 24, 50, 70), (8, 32, 32), 0.5, 3), ((1, 6, 30, 30), (8, 32, 32), 0.1, 1)]:
        ps = tuple(min(a, b) for a, b in zip(shape[1:], max_patch))
        zs, ys, xs = (tiling.strided_starts(shape[i + 1], ps[i], ov) for i in range(3))
        n = shape[0] * len(zs) * len(ys) * len(xs)
        tiles = rng.normal(0, 1, (n, c, *ps)).astype('float32')
        ref = opipe.mo3d_stitch(tiles, shape, (ps, zs, ys, xs)).reshape(shape[0], c, *shape[1:])
        got = E.stitch_ramp_f32(torch.from_numpy(tiles).cuda(), shape[0], c, shape[1:], zs, ys, xs, ps).cpu().numpy()
        assert np.array_equal(got, ref), (shape, np.abs(got - ref).max())


def test_modules_eval_forward_on_gpu():
    """nn.Module surface: eval-mode CUDA forward goes through the engine; CPU eval raises (no CPU fallback)."""
    from bio_image_unet_b200.siam_unet import Siam_UNet
    from bio_image_unet_b200.unet import Unet
    from oracle import models as om
    torch.manual_seed(3)
    m = Unet(n_filter=8).eval()
    m.precision = 'fp32'
    x = torch.rand(2, 1, 32, 32)
    with torch.no_grad():
        ref, ref_logits = om.unet_forward(m.state_dict(), x)
        with pytest.raises(RuntimeError):
            m(x)
        sig, logits = m.cuda()(x.cuda())
    assert (logits.cpu() - ref_logits).abs().max() < 1e-4 and (sig.cpu() - ref).abs().max() < 1e-4
    s = Siam_UNet(n_filter=8, mode='concat').eval().cuda()
    s.precision = 'fp32'
    with torch.no_grad():
        ref, _ = om.siam_forward({k: v.cpu() for k, v in s.state_dict().items()}, x, x.flip(0), 'concat')
        sig, _ = s(x.cuda(), x.flip(0).cuda())
    assert (sig.cpu() - ref).abs().max() < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# multi_output_unet.Predict (2D, several heads; reference: multi_output_unet/predict.py)
# ---------------------------------------------------------------------------------------------------------------
MO2D_HEADS = {'seg': {'channels': 1, 'activation': 'sigmoid'}, 'vec': {'channels': 2, 'activation': None},
              'dist': {'channels': 1, 'activation': 'relu'}}


@pytest.mark.parametrize('name', ['mo2d_single_overlap', 'mo2d_all_pad', 'mo2d_first_holes'])
@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
def test_mo2d_predict_matches_reference_golden(name, precision, tmp_path):
    from bio_image_unet_b200.multi_output_unet import MultiOutputUnet, Predict
    from oracle import pipeline as opipe
    g = _golden.load(name)
    ckpt = str(tmp_path / 'model.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1,
                'output_heads': MO2D_HEADS}, ckpt)
    p = Predict(g['imgs'].copy(), ckpt, result_path=None, network=MultiOutputUnet,
                max_patch_size=tuple(int(v) for v in g['max_patch']), batch_size=2, normalization_mode=str(g['norm_mode']),
                clip_threshold=tuple(float(v) for v in g['clip']), add_tile=int(g['add_tile']), show_progress=False,
                device='cuda:0', precision=precision, keep_intermediates=True)
    # indices, float32 normalisation and patches: bit-exact
    assert tuple(p.patch_size) == tuple(g['patch_size']) and (p.N_x, p.N_y) == (int(g['N_x']), int(g['N_y']))
    assert np.array_equal(p.X_start, g['X_start']) and np.array_equal(p.Y_start, g['Y_start'])
    assert np.array_equal(p.norm, g['norm'])
    assert np.array_equal(p.patches, g['patches'])
    fwd = lambda sd_, x_: omodels.mo2d_forward(sd_, x_, MO2D_HEADS)                        # noqa: E731
    cb = _parity.per_channel_bound(fwd, precision, _golden.state_dict(g), torch.from_numpy(g['patches']).float()[:, None])
    c0 = 0
    info = opipe.mo2d_grid(g['imgs'].shape, tuple(int(v) for v in g['max_patch']), int(g['add_tile']))
    for k, cfg in MO2D_HEADS.items():
        c = cfg['channels']
        ref_rp = g[f'result_patches/{k}'].astype(np.float32)
        scale = max(1.0, np.abs(ref_rp).max())
        got_rp = p.result_patches[:, c0:c0 + c]

        def close(a, b):
            # per head: 1.5 x what the reference's arithmetic at this precision does to that head on these weights
            # (linear heads of the stress nets reach the hundreds; the sigmoid head stays in [0, 1])
            return np.abs(a - b).max() <= float(cb[c0:c0 + c].max()) + 2e-3 * scale
        assert close(got_rp, ref_rp), (k, np.abs(got_rp - ref_rp).max(), scale)
        # stitch: the device kernel against the oracle's stitch of the engine's own (float16-rounded) patches
        st = opipe.mo2d_stitch(got_rp.astype(np.float16), c, g['imgs'].shape, info)
        assert p.result[k].shape == g[f'result/{k}'].shape and p.result[k].dtype == np.float32
        assert np.abs(p.result[k] - st).max() <= 1e-3 * scale, k
        assert close(p.result[k], g[f'result/{k}']), k
        c0 += c


@pytest.mark.parametrize('name', ['nested_single_overlap', 'nested3l_ds_all'])
@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
def test_nested_predict_matches_reference_golden(name, precision, tmp_path):
    """multi_output_unet.Predict with the reference's default network, the nested U-Net++ (and its 3-level variant
    with deep supervision), against fixtures produced by the unmodified reference."""
    from bio_image_unet_b200 import multi_output_unet as mo
    from oracle import pipeline as opipe
    g = _golden.load(name)
    ckpt = str(tmp_path / 'model.pt')
    torch.save({'state_dict': _golden.state_dict(g), 'n_filter': int(g['n_filter']), 'in_channels': 1,
                'output_heads': MO2D_HEADS, 'deep_supervision': bool(int(g['deep_supervision']))}, ckpt)
    p = mo.Predict(g['imgs'].copy(), ckpt, result_path=None, network=getattr(mo, str(g['network'])),
                   max_patch_size=tuple(int(v) for v in g['max_patch']), batch_size=2,
                   normalization_mode=str(g['norm_mode']), clip_threshold=tuple(float(v) for v in g['clip']),
                   add_tile=int(g['add_tile']), show_progress=False, device='cuda:0', precision=precision,
                   keep_intermediates=True)
    assert tuple(p.patch_size) == tuple(g['patch_size']) and (p.N_x, p.N_y) == (int(g['N_x']), int(g['N_y']))
    assert np.array_equal(p.norm, g['norm']) and np.array_equal(p.patches, g['patches'])
    depth = 3 if '3' in str(g['network']) else 4
    ds = bool(int(g['deep_supervision']))
    fwd = lambda sd_, x_: omodels.nested_forward(sd_, x_, MO2D_HEADS, depth, deep_supervision=ds)   # noqa: E731
    cb = _parity.per_channel_bound(fwd, precision, _golden.state_dict(g), torch.from_numpy(g['patches']).float()[:, None])
    info = opipe.mo2d_grid(g['imgs'].shape, tuple(int(v) for v in g['max_patch']), int(g['add_tile']))
    c0 = 0
    for k, cfg in MO2D_HEADS.items():
        c = cfg['channels']
        ref_rp = g[f'result_patches/{k}'].astype(np.float32)
        scale = max(1.0, np.abs(ref_rp).max())
        got_rp = p.result_patches[:, c0:c0 + c]

        def close(a, b):      # same criterion as test_mo2d_predict_matches_reference_golden
            return np.abs(a - b).max() <= float(cb[c0:c0 + c].max()) + 2e-3 * scale
        assert close(got_rp, ref_rp), (k, np.abs(got_rp - ref_rp).max(), scale)
        st = opipe.mo2d_stitch(got_rp.astype(np.float16), c, g['imgs'].shape, info)
        assert np.abs(p.result[k] - st).max() <= 1e-3 * scale, k
        assert close(p.result[k], g[f'result/{k}']), k
        c0 += c


@pytest.mark.parametrize('depth', [4, 3])
@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'bf16'])
def test_nested_nodes_match_oracle(depth, precision):
    """Every node x{l}_{j} of the dense skip pathways (conv outputs written at channel offsets of the per-level
    buffers, bilinear align_corners=True up-sampling) against the oracle, n_filter 16 so the tcgen05 kernels run."""
    from bio_image_unet_b200.engine import Engine
    from bio_image_unet_b200.multi_output_unet import MultiOutputNestedUNet, MultiOutputNestedUNet_3Levels
    from oracle import models as om
    torch.manual_seed(depth)
    heads = {'a': {'channels': 1, 'activation': 'sigmoid'}, 'b': {'channels': 2, 'activation': 'tanh'}}
    nf = 16
    cls = MultiOutputNestedUNet if depth == 4 else MultiOutputNestedUNet_3Levels
    m = cls(1, heads, nf).eval()
    sd = m.state_dict()
    x = torch.rand(2, 1, 64, 96)
    nodes = {}
    with torch.no_grad():
        ref = om.nested_forward(sd, x, heads, depth, collect=nodes)
    eng = Engine('nested2d' if depth == 4 else 'nested2d_3l', sd, nf, 1,
                 [(k, v['channels'], v['activation']) for k, v in heads.items()], precision=precision, device='cuda:0')
    eng.plan(2, (64, 96))
    eng.set_profile(True)
    val, _ = eng.forward(x.cuda(), want_val=True, want_u8=False)
    kinds, _ = eng.read_profile()
    tol = {'fp32': 1e-4, 'tf32': 5e-3, 'bf16': 4e-2}[precision]
    for l in range(depth + 1):
        c = nf * 2 ** l
        upw = 2 * c if l < depth else 0                      # the level's up-sampling slot comes first (net.cu)
        ctot = upw + c * (depth - l + (0 if l == 0 else 1))
        buf = eng.debug_activation(f'x{l}', ctot, l)[:, 0]   # (B, H_l, W_l, ctot) NHWC
        for j in range(depth - l + 1):
            if l == 0 and j == depth:
                continue                                     # x0_{depth} feeds the fused heads only
            want = nodes[f'x{l}_{j}'].permute(0, 2, 3, 1).numpy()
            got = buf[..., upw + j * c:upw + (j + 1) * c]
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got - want).max() <= tol * scale * (1 + l + j), (l, j, np.abs(got - want).max(), scale)
    got = val.cpu().numpy()
    want = np.concatenate([ref[k].numpy() for k in heads], 1)
    assert np.abs(got - want).max() <= tol * 4, np.abs(got - want).max()
    assert precision == 'fp32' or not any(16 <= k < 32 for k in kinds), kinds   # no CUDA-core fallback
    eng.close()


def test_nested_module_eval_forward_and_dilation():
    from bio_image_unet_b200.multi_output_unet import MultiOutputNestedUNet
    from oracle import models as om
    torch.manual_seed(5)
    heads = {'seg': {'channels': 1, 'activation': 'sigmoid'}}
    m = MultiOutputNestedUNet(1, heads, 8, deep_supervision=True, train_mode=False).eval()
    m.precision = 'fp32'
    x = torch.rand(1, 1, 32, 48)
    with torch.no_grad():
        ref = om.nested_forward(m.state_dict(), x, heads, 4, deep_supervision=True)
        out = m.cuda()(x.cuda())
    assert list(out.keys()) == ['seg'] and (out['seg'].cpu() - ref['seg']).abs().max() < 1e-4
    d = MultiOutputNestedUNet(1, heads, 8, dilation=(1, 2, 1, 1, 1)).eval().cuda()
    with pytest.raises(NotImplementedError):
        d(x.cuda())


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('mode', ['concat', 'max', 'corr'])
def test_siam_shared_encoder_is_bit_identical(mode, precision):
    """'single' normalisation: frame t's encoder pass is the same in pair t (current) and pair t + 1 (previous), so the
    Session runs the shared-weight encoder once per frame (biu_net_set_siam_shared). Against the two-pass form: the
    same stitched pages, bit for bit - including frame 0, which is paired with frame 1 (siam_unet/predict.py:108-112),
    chunk boundaries and a zero-padded tail batch."""
    from bio_image_unet_b200.siam_unet import Session, Siam_UNet
    from bio_image_unet_b200.siam_unet.predict import _ArraySource
    torch.manual_seed(4)
    sd = Siam_UNet(n_filter=16, mode=mode).state_dict()
    g = torch.Generator().manual_seed(2)
    for k, v in sd.items():
        if k.endswith('running_var') or k.endswith('.1.weight'):
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean') or k.endswith('.1.bias'):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
    movie = np.random.default_rng(9).integers(0, 4096, (7, 96, 160)).astype('uint16')
    outs = {}
    for shared in (True, False):
        ses = Session({'state_dict': sd, 'n_filter': 16, 'mode': mode}, resize_dim=(64, 64), add_tile=1, device='cuda:0',
                      precision=precision, workspace_gb=4.0)
        assert ses.shared_encoder
        ses.shared_encoder = shared
        out = np.zeros((7, 96, 160), dtype='uint8')

        def sink(first, pages):
            out[first:first + len(pages)] = pages
        ses.predict_stream(_ArraySource(movie), 0, 7, sink, chunk_pairs=3)
        outs[shared] = out.copy()
        ses.close()
    assert outs[True].std() > 0 and np.array_equal(outs[True], outs[False])
    ref = opipe.siam_predict(movie.copy(), sd, mode, (64, 64), False, 'single', (0.0, 99.98), 1)
    if precision == 'fp32':
        assert np.abs(outs[True].astype(np.int16) - ref.astype(np.int16)).max() <= 1


@pytest.mark.parametrize('add_patch', [0, 1, 2])
def test_unet3d_pipelined_groups_equal_whole_volume(add_patch):
    """unet3d.Session.predict without intermediates runs the z-rows in groups from the far end of the volume (stitching
    and copying each group's planes back while the next group computes); with keep=True it runs the whole volume in one
    go. Same result bit for bit - also with overlapping patches, where a group borrows rows of the groups before it -
    and equal to the oracle's whole-volume pipeline."""
    from bio_image_unet_b200.unet3d import Session
    g = _golden.load('unet3d_overlap')
    sd, rd = _golden.state_dict(g), tuple(int(v) for v in g['resize_dim'])
    base = g['vol']
    reps = (-(-(5 * rd[0] + 3) // base.shape[0]), 1, 1)
    vol = np.tile(base, reps)[:5 * rd[0] + 3].copy()
    vol[::2] = vol[::2, ::-1]                                   # no two z-rows alike
    ses = Session({'state_dict': sd, 'n_filter': int(g['n_filter']), 'in_channels': 1, 'out_channels': 1}, rd,
                  add_patch=add_patch, clip_threshold=tuple(g['clip']), device='cuda:0', precision='fp32')
    piped = np.array(ses.predict(vol.copy()))
    assert ses.N_z >= 2
    plain = np.array(ses.predict(vol.copy(), keep=True))
    assert piped.std() > 0 and np.array_equal(piped, plain)
    ref = opipe.unet3d_predict(vol.copy(), sd, rd, False, tuple(g['clip']), add_patch)
    assert np.abs(piped.astype(np.int16) - ref.astype(np.int16)).max() <= 1
    ses.close()
