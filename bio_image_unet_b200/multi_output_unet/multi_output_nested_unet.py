"""Nested U-Net (U-Net++) with several named output heads (reference:
multi_output_unet/multi_output_nested_unet.py:6-240). Same constructors, parameter names / shapes
(``conv{l}_{j}.conv1|bn1|conv2|bn2``, ``output_layers.<head>[_k]``) and forward contract (dict of activated outputs);
eval-mode CUDA forwards run on the B200 engine: every ``torch.cat`` of the dense skip pathways is a channel prefix of
one per-level buffer, the bilinear ``align_corners=True`` up-sampling is a CUDA kernel writing into that buffer and
all heads are fused into the last block's epilogue."""
from typing import Dict, Tuple, Union

import torch
from torch import nn

from ..nn_base import EngineModule


class FirstVGGBlock(nn.Module):
    """VGG block with InstanceNorm (multi_output_nested_unet.py:6-30). Defined by the reference but not used by
    either network; kept for API parity (plain PyTorch)."""

    def __init__(self, in_channels, middle_channels, out_channels, dropout=0.):
        super().__init__()
        self.relu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.conv1 = nn.Conv2d(in_channels, middle_channels, 3, padding=1)
        self.in1 = nn.InstanceNorm2d(middle_channels)
        self.conv2 = nn.Conv2d(middle_channels, out_channels, 3, padding=1)
        self.in2 = nn.InstanceNorm2d(out_channels)
        self.dropout = nn.Dropout2d(dropout)

    def forward(self, x):
        x = self.dropout(self.relu(self.in1(self.conv1(x))))
        return self.dropout(self.relu(self.in2(self.conv2(x))))


class VGGBlock(nn.Module):
    """(Conv3x3 -> BatchNorm -> LeakyReLU(0.1) -> Dropout) x 2 (multi_output_nested_unet.py:33-55)."""

    def __init__(self, in_channels, middle_channels, out_channels, dropout=0., dilation=1):
        super().__init__()
        self.relu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.conv1 = nn.Conv2d(in_channels, middle_channels, kernel_size=3, padding=dilation, dilation=dilation)
        self.bn1 = nn.BatchNorm2d(middle_channels)
        self.conv2 = nn.Conv2d(middle_channels, out_channels, kernel_size=3, padding=dilation, dilation=dilation)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.dropout = nn.Dropout2d(dropout)

    def forward(self, x):
        x = self.dropout(self.relu(self.bn1(self.conv1(x))))
        return self.dropout(self.relu(self.bn2(self.conv2(x))))


class _NestedBase(EngineModule):
    """Shared body of the 4-pool and 3-pool U-Net++ variants; ``depth`` = number of pooling levels."""

    depth = 4
    _engine_kind = 'nested2d'

    def __init__(self, in_channels=1, output_heads: Dict[str, dict] = None, n_filter: int = 32,
                 deep_supervision: bool = False, dilation: Union[bool, Tuple[int, ...]] = False,
                 train_mode: bool = True, **kwargs):
        super().__init__()
        depth = self.depth
        self.output_heads = output_heads or {'default': {'channels': 1, 'activation': 'sigmoid'}}
        self.deep_supervision = deep_supervision
        self.train_mode = train_mode
        self.dilation = dilation if dilation is not False else (1,) * (depth + 1)
        self.in_channels, self.n_filter = in_channels, n_filter
        nb = [n_filter << l for l in range(depth + 1)]

        self.pool = nn.MaxPool2d(2, 2)
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        # registration order = the reference's (column by column), so state_dict() lists the same keys in the same order
        for l in range(depth + 1):
            setattr(self, f'conv{l}_0', self._backbone_block(in_channels if l == 0 else nb[l - 1], nb[l], self.dilation[l]))
        for j in range(1, depth + 1):
            for l in range(depth + 1 - j):
                setattr(self, f'conv{l}_{j}', VGGBlock(nb[l] * j + nb[l + 1], nb[l], nb[l]))
        self.output_layers = nn.ModuleDict()
        for name, config in self.output_heads.items():
            if self.deep_supervision:
                for k in range(1, depth + 1):
                    self.output_layers[f'{name}_{k}'] = nn.Conv2d(nb[0], config['channels'], kernel_size=1)
            else:
                self.output_layers[name] = nn.Conv2d(nb[0], config['channels'], kernel_size=1)

    def _backbone_block(self, cin, cout, dilation):
        return VGGBlock(cin, cout, cout, dilation=dilation)

    def _conv_dilations(self):
        return tuple(int(d) for d in self.dilation)

    def apply_activation(self, x, activation):
        if activation == 'sigmoid':
            return torch.sigmoid(x)
        if activation == 'tanh':
            return torch.tanh(x)
        if activation == 'relu':
            return torch.relu(x)
        return x

    def _engine_spec(self):
        if any(d != 1 for d in self._conv_dilations()):
            raise NotImplementedError('the B200 engine runs the nested U-Net with dilation 1 only')
        suffix = f'_{self.depth}' if self.deep_supervision else ''
        heads = []
        for name, cfg in self.output_heads.items():
            act = cfg.get('activation')
            heads.append((name + suffix, cfg['channels'], act if act in ('sigmoid', 'tanh', 'relu') else None))
        return dict(kind=self._engine_kind, n_filter=self.n_filter, in_channels=self.in_channels, heads=heads)

    def _torch_nodes(self, x):
        """x{0}_{1..depth} of the dense skip pathways (multi_output_nested_unet.py:113-130)."""
        depth = self.depth
        nodes = {}
        for s in range(depth + 1):
            nodes[(s, 0)] = getattr(self, f'conv{s}_0')(x if s == 0 else self.pool(nodes[(s - 1, 0)]))
            for j in range(1, s + 1):
                l = s - j
                cat = [nodes[(l, k)] for k in range(j)] + [self.up(nodes[(l + 1, j - 1)])]
                nodes[(l, j)] = getattr(self, f'conv{l}_{j}')(torch.cat(cat, 1))
        return [nodes[(0, j)] for j in range(1, depth + 1)]

    def forward(self, x):
        """Dict head name -> activated output (multi_output_nested_unet.py:112-148). With deep supervision and
        ``train_mode`` the outputs of all supervision levels are returned as '<head>_<k>' as well."""
        depth = self.depth
        if self.training or (self.deep_supervision and self.train_mode):
            tops = self._torch_nodes(x)
            outputs = {}
            for name, cfg in self.output_heads.items():
                act = cfg.get('activation')
                if self.deep_supervision:
                    if self.train_mode:
                        for k in range(1, depth + 1):
                            outputs[f'{name}_{k}'] = self.apply_activation(self.output_layers[f'{name}_{k}'](tops[k - 1]), act)
                        outputs[name] = outputs[f'{name}_{depth}']
                    else:
                        outputs[name] = self.apply_activation(self.output_layers[f'{name}_{depth}'](tops[-1]), act)
                else:
                    outputs[name] = self.apply_activation(self.output_layers[name](tops[-1]), act)
            return outputs
        val = self._engine_forward(x, raw_logits=False)
        out, c0 = {}, 0
        for name, cfg in self.output_heads.items():
            out[name] = val[:, c0:c0 + cfg['channels']]
            c0 += cfg['channels']
        return out


class MultiOutputNestedUNet(_NestedBase):
    """U-Net++ with four pooling levels (multi_output_nested_unet.py:58-148).

    Parameters
    ----------
    in_channels : int
    output_heads : Dict[str, dict]
        e.g. {'target1': {'channels': 1, 'activation': 'sigmoid'}, 'target2': {'channels': 2, 'activation': None}}
    n_filter : int
    deep_supervision : bool
        one 1x1 head per supervision level ('<head>_1' .. '<head>_4'); inference uses '<head>_4'
    dilation : False or tuple of 5 ints (backbone blocks conv0_0 .. conv4_0)
    train_mode : bool
        with deep supervision, return all supervision outputs (plain PyTorch path)
    """
    depth = 4
    _engine_kind = 'nested2d'


class MultiOutputNestedUNet_3Levels(_NestedBase):
    """U-Net++ with three pooling levels (multi_output_nested_unet.py:151-240)."""
    depth = 3
    _engine_kind = 'nested2d_3l'

    def _backbone_block(self, cin, cout, dilation):
        # as written in the reference (:167-170) the dilation value is passed POSITIONALLY and lands in VGGBlock's
        # `dropout` parameter: the convolutions of this variant always have dilation 1 (and, with the default tuple
        # (1, 1, 1, 1), the backbone trains with Dropout2d(p=1)); inference is unaffected
        return VGGBlock(cin, cout, cout, dilation)

    def _conv_dilations(self):
        return (1,) * (self.depth + 1)
