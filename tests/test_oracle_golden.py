"""Pin the CPU oracle (oracle/) against outputs of the unmodified reference (tests/golden/*.npz).

Integer/byte stages must match bit-for-bit; the float forward runs through the same torch CPU kernels as the
reference, so its uint8-quantised tiles are also compared exactly.

Float stages (the network forward) are compared after the reference's own uint8 quantisation with a +-1 LSB
allowance: oneDNN's reduction order depends on the thread count, and trunc(x*255) flips on 1e-7 differences.
The integer stages downstream are then re-checked exactly from the golden tiles.
"""
import numpy as np
import pytest


def close_u8(a, b, max_frac=2e-3):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return d.max() <= 1 and (d > 0).mean() <= max_frac

from oracle import pipeline
from tests import _golden


@pytest.mark.parametrize('name', ['unet_single', 'unet_all_invert', 'unet_first_u8', 'unet_small_reflect',
                                  'attunet_single', 'unetv0_all', 'unet_f32_single', 'unet_f32_first_invert',
                                  'unet_f32_all'])
def test_unet_pipeline_matches_reference(name):
    g = _golden.load(name)
    imgs = g['imgs'].copy()
    stages = {}
    network = str(g['network']) if 'network' in g else 'Unet'      # 'AttentionUnet' / 'Unet_v0' fixtures
    out = pipeline.unet_predict(imgs, _golden.state_dict(g), tuple(g['resize_dim']), bool(g['invert']), str(g['mode']),
                                tuple(float(v) for v in g['clip']), int(g['add_tile']), stages, network=network)
    n_x, n_y, xs, ys = stages['grid']
    assert (n_x, n_y) == (int(g['N_x']), int(g['N_y']))
    assert np.array_equal(xs, g['X_start']) and np.array_equal(ys, g['Y_start'])
    assert xs.dtype == np.uint16
    if str(g['mode']) == 'single':
        assert np.array_equal(imgs, g['imgs_after'])      # the caller's array is overwritten (unet/predict.py:131)
    assert np.array_equal(stages['patches'], g['patches'])
    assert close_u8(stages['result_patches'], g['result_patches'])
    assert close_u8(out, g['result'])
    # integer stitch, exact, from the reference's own tiles
    st = pipeline.stitch_mean_2d(g['result_patches'], g['imgs'].shape[0], g['imgs'].shape[1:], tuple(g['resize_dim']),
                                 stages['grid'])
    assert np.array_equal(st, g['result'])
    assert np.array_equal(st.astype('float16'), g['result_file'])


@pytest.mark.parametrize('name', ['siam_concat', 'siam_max', 'siam_control_small', 'siam_corr'])
def test_siam_pipeline_matches_reference(name):
    g = _golden.load(name)
    out = pipeline.siam_predict(g['movie'].copy(), _golden.state_dict(g), str(g['siam_mode']), tuple(g['resize_dim']),
                                False, str(g['norm_mode']), tuple(g['clip']), int(g['add_tile']))
    assert out.dtype == np.uint8
    assert close_u8(out, g['result'])
    # per-frame tiles
    movie = g['movie']
    pair = pipeline.siam_preprocess_pair(np.array([movie[1], movie[0]]), str(g['norm_mode']), tuple(g['clip']), False)
    patches, grid = pipeline.siam_split(pair, tuple(g['resize_dim']), int(g['add_tile']))
    assert np.array_equal(patches, g['patches'][0])
    assert np.array_equal(grid[2], g['X_start']) and np.array_equal(grid[3], g['Y_start'])


@pytest.mark.parametrize('name', ['unet3d_overlap', 'unet3d_disjoint', 'unet3d_trilinear'])
def test_unet3d_pipeline_matches_reference(name):
    g = _golden.load(name)
    stages = {}
    out = pipeline.unet3d_predict(g['vol'].copy(), _golden.state_dict(g), tuple(int(v) for v in g['resize_dim']), False,
                                  tuple(g['clip']), int(g['add_patch']), stages,
                                  use_interpolation=bool(g['interp']) if 'interp' in g else False)
    n_z, n_x, n_y, zs, xs, ys = stages['grid']
    assert (n_z, n_x, n_y) == (int(g['N_z']), int(g['N_x']), int(g['N_y']))
    assert np.array_equal(zs, g['Z_start']) and np.array_equal(xs, g['X_start']) and np.array_equal(ys, g['Y_start'])
    assert np.array_equal(stages['patches'], g['patches'])
    assert close_u8(stages['result_patches'], g['result_patches'])
    assert close_u8(out, g['result'])
    st = pipeline.stitch_mod3(g['result_patches'], g['vol'].shape, tuple(int(v) for v in g['resize_dim']), stages['grid'])
    assert np.array_equal(st, g['result'])


@pytest.mark.parametrize('name', ['mo3d_interp', 'mo3d_convt'])
def test_mo3d_pipeline_matches_reference(name):
    g = _golden.load(name)
    out = pipeline.mo3d_predict(g['imgs'].copy(), _golden.state_dict(g), _golden.MO3D_HEADS, bool(g['interp']),
                                tuple(int(v) for v in g['max_patch']), float(g['overlap']), 2, str(g['norm_mode']),
                                tuple(g['clip']))
    ref = _golden.sub(g, 'result')
    assert set(out) == set(ref)
    for k in ref:
        assert out[k].shape == ref[k].shape
        assert np.allclose(out[k], ref[k], rtol=1e-4, atol=1e-5), (k, np.abs(out[k] - ref[k]).max())
    # exact blend from the reference's own patch predictions
    imgs = g['imgs'].astype('float32')[None] if g['imgs'].ndim == 3 else g['imgs'].astype('float32')
    _, info = pipeline.mo3d_split(imgs, tuple(int(v) for v in g['max_patch']), float(g['overlap']))
    for k, rp in _golden.sub(g, 'result_patches').items():
        assert np.array_equal(pipeline.mo3d_stitch(rp, imgs.shape, info), ref[k]), k


def test_mo3d_split_and_weights():
    g = _golden.load('mo3d_interp')
    imgs = pipeline.mo3d_preprocess(g['imgs'].astype('float32'), str(g['norm_mode']), tuple(g['clip']))
    patches, info = pipeline.mo3d_split(imgs, tuple(int(v) for v in g['max_patch']), float(g['overlap']))
    assert np.array_equal(patches, g['patches'])
    assert list(info[1]) == list(g['Z_start']) and list(info[2]) == list(g['Y_start']) and list(info[3]) == list(g['X_start'])
    # survey appendix B: blend profile of a middle patch along x
    w = pipeline.mo3d_patch_weight((1, 4, 40, 40), (0, 1, 1), (1, 3, 3))
    assert w[0, 0, 20, 0] == np.float32(15 / 16) and w[0, 0, 20, 1] == np.float32(1 / 16) and w[0, 0, 20, 39] == 1


MO2D_HEADS = {'seg': {'channels': 1, 'activation': 'sigmoid'}, 'vec': {'channels': 2, 'activation': None},
              'dist': {'channels': 1, 'activation': 'relu'}}


@pytest.mark.parametrize('name', ['mo2d_single_overlap', 'mo2d_all_pad', 'mo2d_first_holes'])
def test_mo2d_pipeline_matches_reference(name):
    """multi_output_unet.Predict(network=MultiOutputUnet): normalisation and patches bit-exact (float32), result
    patches after the reference's own float16 storage within 1 fp16 ulp, stitch exact from the reference's patches."""
    g = _golden.load(name)
    stages = {}
    out = pipeline.mo2d_predict(g['imgs'].copy(), _golden.state_dict(g), MO2D_HEADS, tuple(int(v) for v in g['max_patch']),
                                2, str(g['norm_mode']), tuple(float(v) for v in g['clip']), int(g['add_tile']), stages)
    (ph, pw), n_x, n_y, xs, ys, wx, wy = stages['info']
    assert (ph, pw) == tuple(g['patch_size']) and (n_x, n_y) == (int(g['N_x']), int(g['N_y']))
    assert np.array_equal(xs, g['X_start']) and np.array_equal(ys, g['Y_start'])
    assert stages['norm'].dtype == g['norm'].dtype and np.array_equal(stages['norm'], g['norm'])
    assert np.array_equal(stages['patches'], g['patches'])
    shape = g['imgs'].shape
    for k, cfg in MO2D_HEADS.items():
        rp, ref_rp = stages['result_patches'][k], g[f'result_patches/{k}']
        assert rp.dtype == np.float16 and rp.shape == ref_rp.shape
        assert np.abs(rp.astype(np.float32) - ref_rp.astype(np.float32)).max() <= 2e-3 * max(1.0, np.abs(ref_rp).max())
        st = pipeline.mo2d_stitch(ref_rp, cfg['channels'], shape, stages['info'])
        assert st.dtype == np.float32 and np.array_equal(st, g[f'result/{k}'])
        assert np.abs(out[k] - g[f'result/{k}']).max() <= 2e-3 * max(1.0, np.abs(ref_rp).max())


@pytest.mark.parametrize('name', ['nested_single_overlap', 'nested3l_ds_all'])
def test_nested_pipeline_matches_reference(name):
    """multi_output_unet.Predict with its default network, the nested U-Net++ (MultiOutputNestedUNet, and the 3-level
    variant with deep supervision): the oracle's restatement of the dense skip pathways + bilinear up-sampling
    against the reference's float16 result patches, and the stitched result."""
    g = _golden.load(name)
    stages = {}
    out = pipeline.mo2d_predict(g['imgs'].copy(), _golden.state_dict(g), MO2D_HEADS, tuple(int(v) for v in g['max_patch']),
                                2, str(g['norm_mode']), tuple(float(v) for v in g['clip']), int(g['add_tile']), stages,
                                network=str(g['network']), deep_supervision=bool(int(g['deep_supervision'])))
    assert np.array_equal(stages['norm'], g['norm']) and np.array_equal(stages['patches'], g['patches'])
    for k, cfg in MO2D_HEADS.items():
        rp, ref_rp = stages['result_patches'][k], g[f'result_patches/{k}']
        assert rp.dtype == np.float16 and rp.shape == ref_rp.shape
        assert np.abs(rp.astype(np.float32) - ref_rp.astype(np.float32)).max() <= 2e-3 * max(1.0, np.abs(ref_rp).max())
        assert np.abs(out[k] - g[f'result/{k}']).max() <= 2e-3 * max(1.0, np.abs(ref_rp).max())
