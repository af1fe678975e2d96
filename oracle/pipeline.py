"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's Predict pipelines (normalise, split, per-tile
forward, stitch). Each function cites the reference lines it follows; pinned by tests/test_oracle_golden.py.
"""
import numpy as np
import torch

from . import models


# ------------------------------------------------------------------------------------------------------------
# intensity normalisation
# ------------------------------------------------------------------------------------------------------------
def _norm255(img, lo, hi, invert):
    """clip -> minus min -> / max * 255 [-> 255 - x], float64 (unet/predict.py:125-130)."""
    img = np.clip(img, a_min=lo, a_max=hi)
    img = img - np.min(img)
    img = img / np.max(img) * 255
    if invert:
        img = 255 - img
    return img


def preprocess_stack(imgs, mode, clip_threshold, invert):
    """unet.Predict.__preprocess (unet/predict.py:122-150) == siam 3D branch (siam_unet/predict.py:126-150).
    'single' writes each normalised frame back into the stack => cast to the input dtype (truncation) and the
    caller's array is modified; the other modes return a new float64 array."""
    if mode == 'single':
        for i in range(len(imgs)):
            img = imgs[i]
            imgs[i] = _norm255(img, np.nanpercentile(img, clip_threshold[0]), np.percentile(img, clip_threshold[1]),
                               invert)
        return imgs
    if mode == 'first':
        lo, hi = np.nanpercentile(imgs[0], clip_threshold[0]), np.percentile(imgs[0], clip_threshold[1])
        return _norm255(imgs, lo, hi, invert)
    if mode == 'all':
        lo, hi = np.nanpercentile(imgs, clip_threshold[0]), np.percentile(imgs, clip_threshold[1])
        return _norm255(imgs, lo, hi, invert)
    raise ValueError(f'normalization_mode {mode} not valid!')


def preprocess_volume(vol, clip_threshold, invert):
    """unet3d.Predict.__preprocess: global percentiles, float64 result (unet3d/predict.py:109-117)."""
    return _norm255(vol, np.nanpercentile(vol, clip_threshold[0]), np.percentile(vol, clip_threshold[1]), invert)


# ------------------------------------------------------------------------------------------------------------
# tiling
# ------------------------------------------------------------------------------------------------------------
def tile_starts(extent, tile, count):
    """np.linspace(0, extent - tile, count).astype('uint16') (unet/predict.py:171-172)."""
    return np.linspace(0, extent - tile, count).astype('uint16')


def grid_2d(shape_hw, resize_dim, add_tile):
    """N_x, N_y, X_start, Y_start (unet/predict.py:154-155,171-172; siam_unet/predict.py:92-95,183-184)."""
    n_x = int(np.ceil(shape_hw[0] / resize_dim[0])) + add_tile
    n_y = int(np.ceil(shape_hw[1] / resize_dim[1])) + add_tile
    return n_x, n_y, tile_starts(shape_hw[0], resize_dim[0], n_x), tile_starts(shape_hw[1], resize_dim[1], n_y)


def split_2d(imgs, resize_dim, add_tile, pad_mode='reflect'):
    """unet.Predict.__split (unet/predict.py:152-182): uint8 tiles (N,1,th,tw), order frame -> x -> y."""
    t, h, w = imgs.shape
    th, tw = resize_dim
    n_x, n_y, xs, ys = grid_2d((h, w), resize_dim, add_tile)
    if th > h:
        imgs = np.pad(imgs, ((0, 0), (0, th - h), (0, 0)), pad_mode)
    if tw > w:
        imgs = np.pad(imgs, ((0, 0), (0, 0), (0, tw - w)), pad_mode)
    patches = np.zeros((t * n_x * n_y, 1, th, tw), dtype='uint8')
    n = 0
    for img in imgs:
        for j in range(n_x):
            for k in range(n_y):
                patches[n, 0] = img[xs[j]:xs[j] + th, ys[k]:ys[k] + tw]   # truncating cast to uint8
                n += 1
    return patches, (n_x, n_y, xs, ys)


def grid_3d(vol_shape, resize_dim, add_patch):
    """unet3d.Predict.__split counts incl. the as-written N_x double increment (unet3d/predict.py:121-126)."""
    n_z = int(np.ceil(vol_shape[0] / resize_dim[0])) + add_patch
    n_x = int(np.ceil(vol_shape[1] / resize_dim[1])) + add_patch
    n_y = int(np.ceil(vol_shape[2] / resize_dim[2])) + add_patch
    n_x += add_patch if n_z > 1 else 0
    n_x += add_patch if n_x > 1 else 0
    n_y += add_patch if n_y > 1 else 0
    return (n_z, n_x, n_y, tile_starts(vol_shape[0], resize_dim[0], n_z), tile_starts(vol_shape[1], resize_dim[1], n_x),
            tile_starts(vol_shape[2], resize_dim[2], n_y))


def split_3d(vol, resize_dim, add_patch):
    """unet3d.Predict.__split (unet3d/predict.py:119-153): uint8 patches (N,d,h,w), order z -> x -> y."""
    n_z, n_x, n_y, zs, xs, ys = grid_3d(vol.shape, resize_dim, add_patch)
    gaps = [max(0, resize_dim[i] - vol.shape[i]) for i in range(3)]
    vol = np.pad(vol, ((0, gaps[0]), (0, gaps[1]), (0, gaps[2])), 'reflect')
    d, h, w = resize_dim
    patches = np.zeros((n_z * n_x * n_y, d, h, w), dtype='uint8')
    n = 0
    for j in range(n_z):
        for k in range(n_x):
            for p in range(n_y):
                patches[n] = vol[zs[j]:zs[j] + d, xs[k]:xs[k] + h, ys[p]:ys[p] + w]
                n += 1
    return patches, (n_z, n_x, n_y, zs, xs, ys)


# ------------------------------------------------------------------------------------------------------------
# stitching
# ------------------------------------------------------------------------------------------------------------
def stitch_mean_2d(result_patches, n_frames, shape_hw, resize_dim, grid):
    """unet.Predict.__stitch (unet/predict.py:204-229): float64 NaN stack, nanmean, truncating uint8 store, crop,
    squeeze. result_patches: (N, C, th, tw) uint8."""
    n_x, n_y, xs, ys = grid
    th, tw = resize_dim
    h, w = shape_hw
    c = result_patches.shape[1]
    hh, ww = max(th, h), max(tw, w)
    out = np.zeros((n_frames, c, hh, ww), dtype='uint8')
    per = n_x * n_y
    for i in range(n_frames):
        stack = np.full((per, c, hh, ww), np.nan)
        n = 0
        for j in range(n_x):
            for k in range(n_y):
                stack[n, :, xs[j]:xs[j] + th, ys[k]:ys[k] + tw] = result_patches[i * per + n]
                n += 1
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                out[i] = np.nanmean(stack, axis=0)
    return np.squeeze(out[:, :, :h, :w])


def stitch_mod3(result_patches, vol_shape, resize_dim, grid):
    """unet3d.Predict.__stitch (unet3d/predict.py:173-195): float16 (3,Z,X,Y) NaN buffer, patch n -> slot n % 3
    (overwriting), nanmean over slots, uint8, crop, squeeze."""
    n_z, n_x, n_y, zs, xs, ys = grid
    d, h, w = resize_dim
    buf = np.zeros((3, max(vol_shape[0], d), max(vol_shape[1], h), max(vol_shape[2], w)), dtype='float16') * np.nan
    n = 0
    for i in range(n_z):
        for j in range(n_x):
            for k in range(n_y):
                buf[n % 3, zs[i]:zs[i] + d, xs[j]:xs[j] + h, ys[k]:ys[k] + w] = result_patches[n]
                n += 1
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = np.nanmean(buf, axis=0).astype('uint8')
    return np.squeeze(res[:vol_shape[0], :vol_shape[1], :vol_shape[2]])


# ------------------------------------------------------------------------------------------------------------
# per-tile prediction
# ------------------------------------------------------------------------------------------------------------
def _to_sd(state_dict):
    return {k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in state_dict.items()}


def predict_tiles_unet(sd, patches, network='Unet'):
    """unet.Predict.__predict (unet/predict.py:184-202): batch 1, float32(u8)/255, (sigmoid*255).astype(uint8).
    `network`: 'Unet' | 'AttentionUnet' | 'Unet_v0' (unet/predict.py:89-97)."""
    sd = _to_sd(sd)
    forward = models.FORWARD_2D[network]
    out_ch = sd['final.0.weight'].shape[0]
    res = np.zeros((patches.shape[0], out_ch, *patches.shape[2:]), dtype='uint8')
    with torch.no_grad():
        for i, p in enumerate(patches):
            x = torch.from_numpy(p.astype('float32') / 255).view(1, p.shape[0], *p.shape[1:])
            r = forward(sd, x)[0].view(out_ch, *p.shape[1:]).numpy()
            res[i] = (r * 255).astype('uint8')
    return res


def unet_predict(imgs, sd, resize_dim=(512, 512), invert=False, normalization_mode='single',
                 clip_threshold=(0., 99.8), add_tile=0, stages=None, network='Unet'):
    """unet.Predict end to end (unet/predict.py:54-113) minus file I/O. Returns the stitched uint8 result (the
    reference then stores it as float16, utils/utils.py:21). `stages` (dict) receives intermediates."""
    if imgs.ndim == 2:
        imgs = np.expand_dims(imgs, 0)
    shape = imgs.shape
    imgs = preprocess_stack(imgs, normalization_mode, clip_threshold, invert)
    patches, grid = split_2d(imgs, resize_dim, add_tile)
    res = predict_tiles_unet(sd, patches, network)
    out = stitch_mean_2d(res, shape[0], shape[1:], resize_dim, grid)
    if stages is not None:
        stages.update(patches=patches, result_patches=res, grid=grid)
    return out


def siam_preprocess_pair(pair, mode, clip_threshold, invert):
    """siam.Predict.__preprocess on the (prev, curr) stack + final uint8 cast (siam_unet/predict.py:125-162)."""
    return preprocess_stack(pair, mode, clip_threshold, invert).astype('uint8')


def siam_split(pair_u8, resize_dim, add_tile):
    """siam.Predict.__split (siam_unet/predict.py:164-197): zero padding, ch0 = current, ch1 = previous."""
    _, h, w = pair_u8.shape
    th, tw = resize_dim
    n_x, n_y, xs, ys = grid_2d((h, w), resize_dim, add_tile)
    if h < th:
        pair_u8 = np.pad(pair_u8, ((0, 0), (0, th - h), (0, 0)), 'constant')
    if w < tw:
        pair_u8 = np.pad(pair_u8, ((0, 0), (0, 0), (0, tw - w)), 'constant')
    patches = np.zeros((n_x * n_y, 2, th, tw), dtype='uint8')
    n = 0
    for j in range(n_x):
        for k in range(n_y):
            patches[n, 0] = pair_u8[1][xs[j]:xs[j] + th, ys[k]:ys[k] + tw]
            patches[n, 1] = pair_u8[0][xs[j]:xs[j] + th, ys[k]:ys[k] + tw]
            n += 1
    return patches, (n_x, n_y, xs, ys)


def predict_tiles_siam(sd, patches, mode):
    """siam.Predict.__predict (siam_unet/predict.py:199-215)."""
    sd = _to_sd(sd)
    res = np.zeros((patches.shape[0], 1, patches.shape[2], patches.shape[3]), dtype='uint8')
    with torch.no_grad():
        for i, p in enumerate(patches):
            cur = torch.from_numpy(p[0].astype('float32') / 255).view(1, 1, *p.shape[1:])
            prev = torch.from_numpy(p[1].astype('float32') / 255).view(1, 1, *p.shape[1:])
            r = models.siam_forward(sd, cur, prev, mode)[0].view(1, *p.shape[1:]).numpy() * 255
            res[i] = r.astype('uint8')
    return res


def siam_predict(movie, sd, mode, resize_dim=(512, 512), invert=False, normalization_mode='single',
                 clip_threshold=(0.0, 99.98), add_tile=0):
    """siam.Predict frame loop (siam_unet/predict.py:102-123): pair (prev, curr); frame 0 pairs with frame 1
    (or itself for a single-frame file). Returns the (T,H,W) uint8 stack the TiffWriter receives."""
    t = movie.shape[0]
    if resize_dim is None:
        resize_dim = movie.shape[1:]
    frames = []
    cur = None
    for i in range(t):
        if i == 0:
            prev = np.array(movie[0] if t == 1 else movie[1])
        else:
            prev = cur
        cur = np.array(movie[i])
        pair = siam_preprocess_pair(np.array([prev, cur]), normalization_mode, clip_threshold, invert)
        patches, grid = siam_split(pair, resize_dim, add_tile)
        res = predict_tiles_siam(sd, patches, mode)
        out = stitch_mean_2d(res, 1, movie.shape[1:], resize_dim, grid)
        frames.append(out.astype('uint8').reshape(movie.shape[1:]) if out.ndim != 2 else out)
    return np.stack(frames) if t > 1 else frames[0]


def predict_patches_unet3d(sd, patches, use_interpolation=False):
    """unet3d.Predict.__predict (unet3d/predict.py:155-171)."""
    sd = _to_sd(sd)
    res = np.zeros_like(patches, dtype='uint8')
    with torch.no_grad():
        for i, p in enumerate(patches):
            x = torch.from_numpy(p.astype('float32') / 255).view(1, 1, *p.shape)
            r = models.unet3d_forward(sd, x, use_interpolation)[0].view(*p.shape).numpy()
            res[i] = (r * 255).astype('uint8')
    return res


def unet3d_predict(vol, sd, resize_dim, invert=False, clip_threshold=(0., 99.8), add_patch=0, stages=None,
                   use_interpolation=False):
    """unet3d.Predict end to end (unet3d/predict.py:52-100) minus file I/O."""
    if vol.ndim == 2:
        vol = np.expand_dims(vol, 0)
    shape = vol.shape
    vol = preprocess_volume(vol, clip_threshold, invert)
    patches, grid = split_3d(vol, resize_dim, add_patch)
    res = predict_patches_unet3d(sd, patches, use_interpolation)
    out = stitch_mod3(res, shape, resize_dim, grid)
    if stages is not None:
        stages.update(patches=patches, result_patches=res, grid=grid)
    return out


# ------------------------------------------------------------------------------------------------------------
# multi-output 3D
# ------------------------------------------------------------------------------------------------------------
def mo3d_preprocess(imgs, mode, clip_threshold):
    """multi_output_unet3d.Predict.__preprocess (multi_output_unet3d/predict.py:104-125), float32 in place.
    'single' uses ndarray.ptp() in the reference (removed in numpy 2); np.ptp is the same arithmetic."""
    if mode == 'single':
        for i in range(len(imgs)):
            img = imgs[i]
            c = np.clip(img, np.percentile(img, clip_threshold[0]), np.percentile(img, clip_threshold[1]))
            imgs[i] = (c - c.min()) / (np.ptp(c) + 1e-8)
    elif mode in ('first', 'all'):
        ref = imgs[0] if mode == 'first' else imgs
        lo, hi = np.percentile(ref, [clip_threshold[0], clip_threshold[1]])
        imgs[:] = (np.clip(imgs, lo, hi) - lo) / (hi - lo + 1e-8)
    else:
        raise ValueError(f'Invalid normalization mode: {mode}')
    return imgs


def mo3d_starts(extent, patch, overlap_factor):
    """Start list of one axis (multi_output_unet3d/predict.py:134-147)."""
    stride = max(1, int(patch * (1 - overlap_factor)))
    starts = list(range(0, max(extent - patch + 1, 1), stride))
    if starts[-1] + patch < extent:
        starts.append(extent - patch)
    return starts


def mo3d_split(imgs, max_patch_size, overlap_factor):
    """multi_output_unet3d.Predict.__split (:127-174): (N,1,D,H,W) float32, order vol -> z -> y -> x."""
    n, d, h, w = imgs.shape
    pd, ph, pw = [min(a, b) for a, b in zip((d, h, w), max_patch_size)]
    zs, ys, xs = mo3d_starts(d, pd, overlap_factor), mo3d_starts(h, ph, overlap_factor), mo3d_starts(w, pw, overlap_factor)
    out = []
    for v in range(n):
        for z in zs:
            for y in ys:
                for x in xs:
                    out.append(imgs[v, z:z + pd, y:y + ph, x:x + pw])
    return np.stack(out)[:, None], ((pd, ph, pw), zs, ys, xs)


def mo3d_patch_weight(shape_cdhw, idx, counts, blend_margin=16):
    """Blend weights of one patch as written (:246-272): rules overwrite in z -> y -> x order; the far-side
    loops all write index 0."""
    w = np.ones(shape_cdhw, dtype='float32')
    z_idx, y_idx, x_idx = idx
    n_z, n_y, n_x = counts
    if z_idx > 0:
        for i in range(min(blend_margin, n_z)):
            w[:, i, :, :] = i / blend_margin
    if z_idx < n_z - 1:
        for i in range(min(blend_margin, n_z)):
            w[:, max(-(i + 1), 0), :, :] = i / blend_margin
    if y_idx > 0:
        for i in range(blend_margin):
            w[:, :, i, :] = i / blend_margin
    if y_idx < n_y - 1:
        for i in range(blend_margin):
            w[:, :, max(-(i + 1), 0), :] = i / blend_margin
    if x_idx > 0:
        for i in range(blend_margin):
            w[:, :, :, i] = i / blend_margin
    if x_idx < n_x - 1:
        for i in range(blend_margin):
            w[:, :, :, max(-(i + 1), 0)] = i / blend_margin
    return w


def mo3d_stitch(patches, imgs_shape, split_info, blend_margin=16):
    """multi_output_unet3d.Predict.__stitch for one head (:203-307). patches: (N,C,pd,ph,pw) float32."""
    (pd, ph, pw), zs, ys, xs = split_info
    n_vol, depth, height, width = imgs_shape
    c = patches.shape[1]
    per = len(zs) * len(ys) * len(xs)
    res = np.zeros((n_vol, c, depth, height, width), dtype='float32')
    wmap = np.zeros_like(res)
    for v in range(n_vol):
        vp = patches[v * per:(v + 1) * per].reshape(len(zs), len(ys), len(xs), c, pd, ph, pw)
        for zi, z0 in enumerate(zs):
            for yi, y0 in enumerate(ys):
                for xi, x0 in enumerate(xs):
                    p = vp[zi, yi, xi]
                    w = mo3d_patch_weight(p.shape, (zi, yi, xi), (len(zs), len(ys), len(xs)), blend_margin)
                    z1, y1, x1 = min(z0 + pd, depth), min(y0 + ph, height), min(x0 + pw, width)
                    sl = (slice(None), slice(0, z1 - z0), slice(0, y1 - y0), slice(0, x1 - x0))
                    res[v, :, z0:z1, y0:y1, x0:x1] += p[sl] * w[sl]
                    wmap[v, :, z0:z1, y0:y1, x0:x1] += w[sl]
    mask = wmap > 0
    res[mask] = res[mask] / wmap[mask]
    res[~mask] = 0
    return np.squeeze(res)


def mo3d_predict(imgs, sd, output_heads, use_interpolation=True, max_patch_size=(64, 256, 256), overlap_factor=0.1,
                 batch_size=1, normalization_mode='single', clip_threshold=(0., 99.98)):
    """multi_output_unet3d.Predict end to end (:16-88) minus file I/O. Returns {head: ndarray}."""
    sd = _to_sd(sd)
    imgs = imgs.astype('float32')
    if imgs.ndim == 3:
        imgs = np.expand_dims(imgs, 0)
    elif imgs.ndim != 4:
        raise ValueError(f'Unsupported input shape: {imgs.shape}')
    shape = imgs.shape
    imgs = mo3d_preprocess(imgs, normalization_mode, clip_threshold)
    patches, info = mo3d_split(imgs, max_patch_size, overlap_factor)
    results = {k: [] for k in output_heads}
    with torch.no_grad():
        for i in range(int(np.ceil(len(patches) / batch_size))):
            batch = torch.tensor(patches[i * batch_size:(i + 1) * batch_size], dtype=torch.float32)
            preds = models.mo3d_forward(sd, batch, output_heads, use_interpolation)
            for k in results:
                results[k].append(preds[k].numpy())
    return {k: mo3d_stitch(np.concatenate(results[k]), shape, info) for k in results}


# ---------------------------------------------------------------------------------------------------------------
# multi_output_unet.Predict (2D, several heads): multi_output_unet/predict.py:13-285
# ---------------------------------------------------------------------------------------------------------------
def mo2d_preprocess(imgs, mode, clip_threshold):
    """Predict.__preprocess (:128-151) on the float32 stack (T, H, W); numpy's own promotion rules apply."""
    if mode == 'single':
        for i, img in enumerate(imgs):
            img = np.clip(img, a_min=np.nanpercentile(img, clip_threshold[0]), a_max=np.percentile(img, clip_threshold[1]))
            img = img - np.min(img)
            img = img / np.max(img)
            imgs[i] = img
    elif mode in ('first', 'all'):
        src = imgs[0] if mode == 'first' else imgs
        ct = (np.nanpercentile(src, clip_threshold[0]), np.percentile(src, clip_threshold[1]))
        imgs = np.clip(imgs, ct[0], ct[1])
        imgs = imgs - np.min(imgs)
        imgs = imgs / np.max(imgs)
    else:
        raise ValueError(f'normalization_mode {mode} not valid!')
    return imgs


def mo2d_grid(shape, max_patch_size, add_tile):
    """Patch size (rounded up to a multiple of 16), tile counts, linspace starts and the sliding-window starts the
    patches are REALLY taken at (:153-184: windows every X_start[1] pixels, which can differ from the linspace)."""
    _, h, w = shape
    ph = ((min(h, max_patch_size[0]) + 15) // 16) * 16
    pw = ((min(w, max_patch_size[1]) + 15) // 16) * 16
    n_x = int(np.ceil(h / ph)) + add_tile
    n_y = int(np.ceil(w / pw)) + add_tile
    hp, wp = h + max(ph - h, 0), w + max(pw - w, 0)
    xs = np.linspace(0, hp - ph, n_x).astype('uint16')
    ys = np.linspace(0, wp - pw, n_y).astype('uint16')
    sx = int(xs[1]) if n_x > 1 else 1
    sy = int(ys[1]) if n_y > 1 else 1
    wx = np.arange(0, hp - ph + 1, sx)
    wy = np.arange(0, wp - pw + 1, sy)
    return (ph, pw), n_x, n_y, xs, ys, wx, wy


def mo2d_split(imgs, max_patch_size, add_tile):
    (ph, pw), n_x, n_y, xs, ys, wx, wy = mo2d_grid(imgs.shape, max_patch_size, add_tile)
    imgs = np.pad(imgs, ((0, 0), (0, max(ph - imgs.shape[1], 0)), (0, max(pw - imgs.shape[2], 0))), 'reflect')
    patches = np.lib.stride_tricks.sliding_window_view(imgs, (ph, pw), axis=(1, 2))
    patches = patches[:, ::int(xs[1]) if n_x > 1 else 1, ::int(ys[1]) if n_y > 1 else 1]
    return patches.reshape(-1, ph, pw), ((ph, pw), n_x, n_y, xs, ys, wx, wy)


def mo2d_stitch(result_patches, channels, shape, info, safe_margin=20):
    """Predict.__stitch (:230-285) for one head: margin-weighted mean, holes filled with the global mean of the
    (float16) result patches. result_patches: float16 (N, C, ph, pw)."""
    (ph, pw), n_x, n_y, xs, ys, _, _ = info
    t, h, w = shape
    n_per_img = n_x * n_y
    res = np.zeros((t, channels, max(ph, h), max(pw, w)), dtype='float32')
    weight = np.zeros_like(res)
    for i in range(t):
        stack = result_patches[i * n_per_img:(i + 1) * n_per_img].reshape(n_x, n_y, *result_patches.shape[1:])
        for j, x0 in enumerate(xs):
            for k, y0 in enumerate(ys):
                patch = stack[j, k]
                pwt = np.ones_like(patch)
                if j > 0:
                    pwt[..., :safe_margin, :] = 0
                if j < n_x - 1:
                    pwt[..., -safe_margin:, :] = 0
                if k > 0:
                    pwt[..., :safe_margin] = 0
                if k < n_y - 1:
                    pwt[..., -safe_margin:] = 0
                res[i, :, int(x0):int(x0) + ph, int(y0):int(y0) + pw] += patch * pwt
                weight[i, :, int(x0):int(x0) + ph, int(y0):int(y0) + pw] += pwt
    np.divide(res, weight, out=res, where=weight > 0)
    res[weight == 0] = result_patches.mean()
    return np.squeeze(res[:, :, :h, :w])


def mo2d_predict(imgs, sd, output_heads, max_patch_size=(1024, 1024), batch_size=1, normalization_mode='single',
                 clip_threshold=(0., 99.98), add_tile=0, stages=None, network='MultiOutputUnet', deep_supervision=False):
    """multi_output_unet.Predict end to end (:16-126) on the CPU (float32 model, float16 result patches) minus file
    I/O; network = 'MultiOutputUnet' | 'MultiOutputNestedUNet' | 'MultiOutputNestedUNet_3Levels'.
    Returns {head: ndarray}."""
    sd = _to_sd(sd)
    imgs = imgs.astype('float32')
    if imgs.ndim == 2:
        imgs = np.expand_dims(imgs, 0)
    shape = imgs.shape
    imgs = mo2d_preprocess(imgs, normalization_mode, clip_threshold)
    patches, info = mo2d_split(imgs, max_patch_size, add_tile)
    ph, pw = info[0]
    results = {k: np.zeros((patches.shape[0], cfg['channels'], ph, pw), dtype='float16') for k, cfg in output_heads.items()}
    with torch.no_grad():
        for i in range(int(np.ceil(len(patches) / batch_size))):
            batch = torch.tensor(patches[i * batch_size:(i + 1) * batch_size], dtype=torch.float32).view(-1, 1, ph, pw)
            if network == 'MultiOutputUnet':
                preds = models.mo2d_forward(sd, batch, output_heads)
            else:
                preds = models.nested_forward(sd, batch, output_heads, 4 if network == 'MultiOutputNestedUNet' else 3,
                                              deep_supervision)
            for k in results:
                results[k][i * batch_size:(i + 1) * batch_size] = preds[k].numpy()
    if stages is not None:
        stages.update(norm=imgs, patches=patches, info=info, result_patches=results)
    return {k: mo2d_stitch(results[k], output_heads[k]['channels'], shape, info) for k in results}
