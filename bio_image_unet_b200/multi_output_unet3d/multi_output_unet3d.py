"""Multi-output 3D U-Net (reference: multi_output_unet3d/multi_output_unet3d.py:7-170): UNet3D body with a
dictionary of 1x1x1 output heads, each with its own activation; by default nearest-neighbour resampling replaces
max-pooling / transposed convolutions (use_interpolation=True)."""
from typing import Dict

import torch
import torch.nn.functional as F
from torch import nn

from ..nn_base import EngineModule, conv_block
from ..unet3d.unet3d import body3d_channels


class MultiOutputUnet3D(EngineModule):
    def __init__(self, in_channels: int = 1, output_heads: Dict[str, dict] = None, n_filter: int = 16,
                 use_interpolation: bool = True):
        super().__init__()
        self.output_heads = output_heads or {'default': {'channels': 1, 'activation': 'sigmoid'}}
        self.use_interpolation = use_interpolation
        self.in_channels, self.n_filter = in_channels, n_filter
        enc, mid, dec, ups = body3d_channels(n_filter)
        for i, (cin, cout) in enumerate(enc):
            setattr(self, f'encode{i + 1}', conv_block(3, in_channels if cin is None else cin, cout))
            if i % 2 == 1 and not use_interpolation:
                setattr(self, f'maxpool{i // 2 + 1}', nn.MaxPool3d(kernel_size=2, stride=2))
        self.middle_conv1 = conv_block(3, *mid[0])
        self.middle_conv2 = conv_block(3, *mid[1])
        for k, c in enumerate(ups):
            if use_interpolation:
                setattr(self, f'up{k + 1}_conv', conv_block(3, c, c))
            else:
                setattr(self, f'up{k + 1}', nn.ConvTranspose3d(c, c, kernel_size=2, stride=2))
        for i, (cin, cout) in enumerate(dec):
            setattr(self, f'decode{i + 1}', conv_block(3, cin, cout))
        self.output_layers = nn.ModuleDict()
        for name, cfg in self.output_heads.items():
            self.output_layers[name] = nn.Conv3d(n_filter // 2, cfg['channels'], kernel_size=1)

    def _engine_spec(self):
        return dict(kind='mo3d', n_filter=self.n_filter, in_channels=self.in_channels,
                    heads=[(n, c['channels'], c.get('activation')) for n, c in self.output_heads.items()],
                    use_interpolation=self.use_interpolation)

    @staticmethod
    def apply_activation(x, activation):
        if activation == 'sigmoid':
            return torch.sigmoid(x)
        if activation == 'tanh':
            return torch.tanh(x)
        if activation == 'relu':
            return F.relu(x)
        return x

    def _torch_body(self, x):
        skips = []
        for level in range(3):
            x = getattr(self, f'encode{2 * level + 2}')(getattr(self, f'encode{2 * level + 1}')(x))
            skips.append(x)
            x = F.interpolate(x, scale_factor=0.5, mode='nearest') if self.use_interpolation else \
                getattr(self, f'maxpool{level + 1}')(x)
        x = self.middle_conv2(self.middle_conv1(x))
        for k in range(3):
            if self.use_interpolation:
                x = getattr(self, f'up{k + 1}_conv')(F.interpolate(x, scale_factor=2, mode='nearest'))
            else:
                x = getattr(self, f'up{k + 1}')(x)
            x = torch.cat((x, skips[2 - k]), dim=1)
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return x

    def forward(self, x):
        """dict head name -> activated tensor (multi_output_unet3d.py:164-170)."""
        if self.training:
            d6 = self._torch_body(x)
            return {n: self.apply_activation(self.output_layers[n](d6), c.get('activation'))
                    for n, c in self.output_heads.items()}
        val = self._engine_forward(x, raw_logits=False)
        out, c0 = {}, 0
        for n, c in self.output_heads.items():
            out[n] = val[:, c0:c0 + c['channels']]
            c0 += c['channels']
        return out
