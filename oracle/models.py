"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch.nn.functional, fp32) of the reference's model forwards.

Each function takes the reference's ``state_dict`` (name -> tensor) and follows the cited reference lines.
Pinned against the unmodified reference by tests/test_oracle_golden.py (fixtures: tests/golden/*.npz).
"""
import torch
import torch.nn.functional as F


def _block(sd, name, x):
    """Conv(k=3,pad=1) -> BatchNorm(eval, eps=1e-5) -> LeakyReLU(0.1) -> Dropout(p=0)
    (unet/unet.py:54-60, unet3d/unet3d.py:52-58)."""
    w, b = sd[f'{name}.0.weight'], sd[f'{name}.0.bias']
    conv = F.conv3d if w.dim() == 5 else F.conv2d
    y = conv(x, w, b, padding=1)
    y = F.batch_norm(y, sd[f'{name}.1.running_mean'], sd[f'{name}.1.running_var'], sd[f'{name}.1.weight'],
                     sd[f'{name}.1.bias'], training=False, eps=1e-5)
    return F.leaky_relu(y, 0.1)


def _up(sd, name, x):
    """ConvTranspose(k=2,s=2), no BN / activation (unet/unet.py:38-47,87)."""
    w, b = sd[f'{name}.weight'], sd[f'{name}.bias']
    return (F.conv_transpose3d if w.dim() == 5 else F.conv_transpose2d)(x, w, b, stride=2)


def unet_forward(sd, x, collect=None):
    """Unet.forward, unet/unet.py:69-104. Returns (sigmoid(logits), logits)."""
    acts = {}
    e1 = _block(sd, 'encode1', x); e2 = _block(sd, 'encode2', e1); m1 = F.max_pool2d(e2, 2, 2)
    e3 = _block(sd, 'encode3', m1); e4 = _block(sd, 'encode4', e3); m2 = F.max_pool2d(e4, 2, 2)
    e5 = _block(sd, 'encode5', m2); e6 = _block(sd, 'encode6', e5); m3 = F.max_pool2d(e6, 2, 2)
    e7 = _block(sd, 'encode7', m3); e8 = _block(sd, 'encode8', e7); m4 = F.max_pool2d(e8, 2, 2)
    mid1 = _block(sd, 'middle_conv1', m4); mid2 = _block(sd, 'middle_conv2', mid1)
    u1 = _up(sd, 'up1', mid2); d1 = _block(sd, 'decode1', torch.cat((u1, e8), 1)); d2 = _block(sd, 'decode2', d1)
    u2 = _up(sd, 'up2', d2); d3 = _block(sd, 'decode3', torch.cat((u2, e6), 1)); d4 = _block(sd, 'decode4', d3)
    u3 = _up(sd, 'up3', d4); d5 = _block(sd, 'decode5', torch.cat((u3, e4), 1)); d6 = _block(sd, 'decode6', d5)
    u4 = _up(sd, 'up4', d6); d7 = _block(sd, 'decode7', torch.cat((u4, e2), 1)); d8 = _block(sd, 'decode8', d7)
    logits = F.conv2d(d8, sd['final.0.weight'], sd['final.0.bias'])
    if collect is not None:
        collect.update(dict(e1=e1, e2=e2, m1=m1, e3=e3, e4=e4, m2=m2, e5=e5, e6=e6, m3=m3, e7=e7, e8=e8, m4=m4,
                            mid1=mid1, mid2=mid2, u1=u1, d1=d1, d2=d2, u2=u2, d3=d3, d4=d4, u3=u3, d5=d5, d6=d6,
                            u4=u4, d7=d7, d8=d8))
    return torch.sigmoid(logits), logits


def _block_relu(sd, name, x):
    """Unet_v0's block: Conv(k=3,pad=1) -> BatchNorm(eval) -> ReLU -> Dropout (identity in eval),
    unet/unet_v0.py:56-63."""
    w, b = sd[f'{name}.0.weight'], sd[f'{name}.0.bias']
    y = F.conv2d(x, w, b, padding=1)
    y = F.batch_norm(y, sd[f'{name}.1.running_mean'], sd[f'{name}.1.running_var'], sd[f'{name}.1.weight'],
                     sd[f'{name}.1.bias'], training=False, eps=1e-5)
    return F.relu(y)


def unet_v0_forward(sd, x):
    """Unet_v0.forward, unet/unet_v0.py:72-106: ReLU blocks, skips taken after the FIRST conv of each level
    (e7/e5/e3/e1), an extra 3x3 block decode9 (n_filter -> 1) before the 1x1 head."""
    B = _block_relu
    e1 = B(sd, 'encode1', x); e2 = B(sd, 'encode2', e1); m1 = F.max_pool2d(e2, 2, 2)
    e3 = B(sd, 'encode3', m1); e4 = B(sd, 'encode4', e3); m2 = F.max_pool2d(e4, 2, 2)
    e5 = B(sd, 'encode5', m2); e6 = B(sd, 'encode6', e5); m3 = F.max_pool2d(e6, 2, 2)
    e7 = B(sd, 'encode7', m3); e8 = B(sd, 'encode8', e7); m4 = F.max_pool2d(e8, 2, 2)
    mid2 = B(sd, 'middle_conv2', B(sd, 'middle_conv1', m4))
    d2 = B(sd, 'decode2', B(sd, 'decode1', torch.cat((_up(sd, 'up1', mid2), e7), 1)))
    d4 = B(sd, 'decode4', B(sd, 'decode3', torch.cat((_up(sd, 'up2', d2), e5), 1)))
    d6 = B(sd, 'decode6', B(sd, 'decode5', torch.cat((_up(sd, 'up3', d4), e3), 1)))
    d8 = B(sd, 'decode8', B(sd, 'decode7', torch.cat((_up(sd, 'up4', d6), e1), 1)))
    d9 = B(sd, 'decode9', d8)
    logits = F.conv2d(d9, sd['final.0.weight'], sd['final.0.bias'])
    return torch.sigmoid(logits), logits


def _attention(sd, name, gate, skip):
    """AttentionBlock.forward, unet/attention_unet.py:163-181: psi = sigmoid(BN(conv1x1(relu(BN(conv1x1(gate)) +
    BN(conv1x1(skip)))))), out = skip * psi."""
    def cb(prefix, t):
        y = F.conv2d(t, sd[f'{prefix}.0.weight'], sd[f'{prefix}.0.bias'])
        return F.batch_norm(y, sd[f'{prefix}.1.running_mean'], sd[f'{prefix}.1.running_var'], sd[f'{prefix}.1.weight'],
                            sd[f'{prefix}.1.bias'], training=False, eps=1e-5)
    psi = F.relu(cb(f'{name}.W_gate', gate) + cb(f'{name}.W_x', skip))
    psi = torch.sigmoid(cb(f'{name}.psi', psi))
    return skip * psi


def attention_unet_forward(sd, x, collect=None):
    """AttentionUnet.forward, unet/attention_unet.py:69-108: Unet encoder; in the decoder the skip tensor is gated
    by the up-sampled tensor and concatenated FIRST: cat((attention(u, e), u))."""
    e1 = _block(sd, 'encode1', x); e2 = _block(sd, 'encode2', e1); m1 = F.max_pool2d(e2, 2, 2)
    e3 = _block(sd, 'encode3', m1); e4 = _block(sd, 'encode4', e3); m2 = F.max_pool2d(e4, 2, 2)
    e5 = _block(sd, 'encode5', m2); e6 = _block(sd, 'encode6', e5); m3 = F.max_pool2d(e6, 2, 2)
    e7 = _block(sd, 'encode7', m3); e8 = _block(sd, 'encode8', e7); m4 = F.max_pool2d(e8, 2, 2)
    x = _block(sd, 'middle_conv2', _block(sd, 'middle_conv1', m4))
    for k, e in enumerate((e8, e6, e4, e2)):
        u = _up(sd, f'up{k + 1}', x)
        a = _attention(sd, f'attention{k + 1}', u, e)
        if collect is not None:
            collect[f'a{k + 1}'] = a
        x = _block(sd, f'decode{2 * k + 2}', _block(sd, f'decode{2 * k + 1}', torch.cat((a, u), 1)))
    logits = F.conv2d(x, sd['final.0.weight'], sd['final.0.bias'])
    return torch.sigmoid(logits), logits


def mo2d_forward(sd, x, output_heads):
    """MultiOutputUnet.forward, multi_output_unet/multi_output_unet.py:88-134: Unet body + one 1x1 conv per head."""
    acts = {}
    unet_body = dict(sd)
    unet_body['final.0.weight'] = torch.zeros(1, sd['decode8.0.weight'].shape[0], 1, 1)
    unet_body['final.0.bias'] = torch.zeros(1)
    unet_forward(unet_body, x, collect=acts)
    d8 = acts['d8']
    out = {}
    for name, cfg in output_heads.items():
        logits = F.conv2d(d8, sd[f'output_layers.{name}.weight'], sd[f'output_layers.{name}.bias'])
        act = cfg.get('activation')
        if act == 'sigmoid':
            logits = torch.sigmoid(logits)
        elif act == 'tanh':
            logits = torch.tanh(logits)
        elif act == 'relu':
            logits = torch.relu(logits)
        out[name] = logits
    return out


def _apply_heads(sd, feat, output_heads, suffix=''):
    out = {}
    for name, cfg in output_heads.items():
        logits = F.conv2d(feat, sd[f'output_layers.{name}{suffix}.weight'], sd[f'output_layers.{name}{suffix}.bias'])
        act = cfg.get('activation')
        if act == 'sigmoid':
            logits = torch.sigmoid(logits)
        elif act == 'tanh':
            logits = torch.tanh(logits)
        elif act == 'relu':
            logits = torch.relu(logits)
        out[name] = logits
    return out


def _vgg(sd, name, x):
    """VGGBlock.forward (dilation 1, eval), multi_output_unet/multi_output_nested_unet.py:33-55:
    (Conv3x3 -> BatchNorm -> LeakyReLU(0.1) -> Dropout(identity)) x 2 with parameters conv1/bn1/conv2/bn2."""
    for k in (1, 2):
        x = F.conv2d(x, sd[f'{name}.conv{k}.weight'], sd[f'{name}.conv{k}.bias'], padding=1)
        x = F.batch_norm(x, sd[f'{name}.bn{k}.running_mean'], sd[f'{name}.bn{k}.running_var'], sd[f'{name}.bn{k}.weight'],
                         sd[f'{name}.bn{k}.bias'], training=False, eps=1e-5)
        x = F.leaky_relu(x, 0.1)
    return x


def nested_forward(sd, x, output_heads, depth=4, deep_supervision=False, collect=None):
    """MultiOutputNestedUNet.forward (depth=4, multi_output_nested_unet.py:112-148) and
    MultiOutputNestedUNet_3Levels.forward (depth=3, :207-240) in inference configuration (train_mode=False): nodes
    x{l}_{j} = VGG(cat(x{l}_0 .. x{l}_{j-1}, up(x{l+1}_{j-1}))), up = bilinear x2 align_corners=True; heads on
    x0_{depth} ('<head>_{depth}' with deep supervision)."""
    def up(t):
        return F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=True)
    nodes = {}
    for s in range(depth + 1):
        nodes[(s, 0)] = _vgg(sd, f'conv{s}_0', x if s == 0 else F.max_pool2d(nodes[(s - 1, 0)], 2, 2))
        for j in range(1, s + 1):
            l = s - j
            cat = [nodes[(l, k)] for k in range(j)] + [up(nodes[(l + 1, j - 1)])]
            nodes[(l, j)] = _vgg(sd, f'conv{l}_{j}', torch.cat(cat, 1))
    if collect is not None:
        collect.update({f'x{l}_{j}': v for (l, j), v in nodes.items()})
    return _apply_heads(sd, nodes[(0, depth)], output_heads, f'_{depth}' if deep_supervision else '')


FORWARD_2D = {'Unet': unet_forward, 'Unet_v0': unet_v0_forward, 'AttentionUnet': attention_unet_forward}


def _siam_encoder(sd, x):
    """One pass of the shared-weight encoder, siam_unet/siam_unet.py:87-98 (== :101-112)."""
    e1 = _block(sd, 'encode1', x); e2 = _block(sd, 'encode2', e1); m1 = F.max_pool2d(e2, 2, 2)
    e3 = _block(sd, 'encode3', m1); e4 = _block(sd, 'encode4', e3); m2 = F.max_pool2d(e4, 2, 2)
    e5 = _block(sd, 'encode5', m2); e6 = _block(sd, 'encode6', e5); m3 = F.max_pool2d(e6, 2, 2)
    e7 = _block(sd, 'encode7', m3); e8 = _block(sd, 'encode8', e7); m4 = F.max_pool2d(e8, 2, 2)
    return (e2, e4, e6, e8), m4


def siam_forward(sd, x, prev_x, mode='concat'):
    """Siam_UNet.forward, siam_unet/siam_unet.py:85-148."""
    (e2, e4, e6, e8), m4 = _siam_encoder(sd, x)
    _, mm4 = _siam_encoder(sd, prev_x)
    if mode == 'corr':      # depthwise_xcorr, :75-83
        b, c = mm4.size(0), mm4.size(1)
        out = F.conv2d(m4.reshape(1, b * c, m4.size(2), m4.size(3)), mm4.reshape(b * c, 1, mm4.size(2), mm4.size(3)),
                       groups=b * c, padding='same')
        join = out.view(b, c, out.size(2), out.size(3))
    elif mode == 'max':
        join = torch.maximum(m4, mm4)
    elif mode == 'concat':
        join = _block(sd, 'conv_concat', torch.cat((m4, mm4), 1))
    elif mode == 'control':
        join = m4
    else:
        raise NotImplementedError('Unknown mode: {}'.format(mode))
    mid1 = _block(sd, 'middle_conv1', join); mid2 = _block(sd, 'middle_conv2', mid1)
    u1 = _up(sd, 'up1', mid2); d1 = _block(sd, 'decode1', torch.cat((u1, e8), 1)); d2 = _block(sd, 'decode2', d1)
    u2 = _up(sd, 'up2', d2); d3 = _block(sd, 'decode3', torch.cat((u2, e6), 1)); d4 = _block(sd, 'decode4', d3)
    u3 = _up(sd, 'up3', d4); d5 = _block(sd, 'decode5', torch.cat((u3, e4), 1)); d6 = _block(sd, 'decode6', d5)
    u4 = _up(sd, 'up4', d6); d7 = _block(sd, 'decode7', torch.cat((u4, e2), 1)); d8 = _block(sd, 'decode8', d7)
    logits = F.conv2d(d8, sd['final.0.weight'], sd['final.0.bias'])
    return torch.sigmoid(logits), logits


def _body3d(sd, x, interp_updown):
    """Shared body of UNet3D (unet3d/unet3d.py:63-97, transposed-conv path) and MultiOutputUnet3D
    (multi_output_unet3d/multi_output_unet3d.py:106-161; interp_updown=True is its default path)."""
    def down(t):
        return F.interpolate(t, scale_factor=0.5, mode='nearest') if interp_updown else F.max_pool3d(t, 2, 2)

    def up(name, t):
        if interp_updown:
            return _block(sd, name + '_conv', F.interpolate(t, scale_factor=2, mode='nearest'))
        return _up(sd, name, t)

    e1 = _block(sd, 'encode1', x); e2 = _block(sd, 'encode2', e1); m1 = down(e2)
    e3 = _block(sd, 'encode3', m1); e4 = _block(sd, 'encode4', e3); m2 = down(e4)
    e5 = _block(sd, 'encode5', m2); e6 = _block(sd, 'encode6', e5); m3 = down(e6)
    mid1 = _block(sd, 'middle_conv1', m3); mid2 = _block(sd, 'middle_conv2', mid1)
    u1 = up('up1', mid2); d1 = _block(sd, 'decode1', torch.cat((u1, e6), 1)); d2 = _block(sd, 'decode2', d1)
    u2 = up('up2', d2); d3 = _block(sd, 'decode3', torch.cat((u2, e4), 1)); d4 = _block(sd, 'decode4', d3)
    u3 = up('up3', d4); d5 = _block(sd, 'decode5', torch.cat((u3, e2), 1)); d6 = _block(sd, 'decode6', d5)
    return d6


def unet3d_forward(sd, x, use_interpolation=False):
    """UNet3D.forward, unet3d/unet3d.py:63-99 (use_interpolation=True: trilinear x2 instead of the transposed
    convolutions, max-pooling stays)."""
    if use_interpolation:
        def up(t):
            return F.interpolate(t, scale_factor=2, mode='trilinear', align_corners=False)
        e1 = _block(sd, 'encode1', x); e2 = _block(sd, 'encode2', e1); m1 = F.max_pool3d(e2, 2, 2)
        e3 = _block(sd, 'encode3', m1); e4 = _block(sd, 'encode4', e3); m2 = F.max_pool3d(e4, 2, 2)
        e5 = _block(sd, 'encode5', m2); e6 = _block(sd, 'encode6', e5); m3 = F.max_pool3d(e6, 2, 2)
        mid2 = _block(sd, 'middle_conv2', _block(sd, 'middle_conv1', m3))
        d2 = _block(sd, 'decode2', _block(sd, 'decode1', torch.cat((up(mid2), e6), 1)))
        d4 = _block(sd, 'decode4', _block(sd, 'decode3', torch.cat((up(d2), e4), 1)))
        d6 = _block(sd, 'decode6', _block(sd, 'decode5', torch.cat((up(d4), e2), 1)))
        logits = F.conv3d(d6, sd['final.weight'], sd['final.bias'])
        return torch.sigmoid(logits), logits
    d6 = _body3d(sd, x, False)
    logits = F.conv3d(d6, sd['final.weight'], sd['final.bias'])
    return torch.sigmoid(logits), logits


def mo3d_forward(sd, x, output_heads, use_interpolation=True):
    """MultiOutputUnet3D.forward, multi_output_unet3d/multi_output_unet3d.py:106-170."""
    d6 = _body3d(sd, x, use_interpolation)
    out = {}
    for name, cfg in output_heads.items():
        logits = F.conv3d(d6, sd[f'output_layers.{name}.weight'], sd[f'output_layers.{name}.bias'])
        act = cfg.get('activation')
        if act == 'sigmoid':
            logits = torch.sigmoid(logits)
        elif act == 'tanh':
            logits = torch.tanh(logits)
        elif act == 'relu':
            logits = F.relu(logits)
        out[name] = logits
    return out
