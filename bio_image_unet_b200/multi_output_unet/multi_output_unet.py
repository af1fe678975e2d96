"""2D U-Net with several named output heads (reference: multi_output_unet/multi_output_unet.py:6-134). Same
constructor, parameter names / shapes (``output_layers.<head>``) and forward contract (dict of activated outputs);
eval-mode CUDA forwards run on the B200 engine with all heads fused into the last block's epilogue."""
from typing import Dict

import torch
from torch import nn

from ..nn_base import EngineModule, conv_block


class MultiOutputUnet(EngineModule):
    """Multi-output U-Net.

    Parameters
    ----------
    in_channels : int
    output_heads : Dict[str, dict]
        e.g. {'target1': {'channels': 1, 'activation': 'sigmoid'}, 'target2': {'channels': 2, 'activation': None}};
        activations 'sigmoid' | 'tanh' | 'relu' | None (multi_output_unet.py:79-86; anything else is the identity)
    n_filter : int
    **kwargs : ignored (``Predict`` passes deep_supervision / train_mode, multi_output_unet/predict.py:90-94)
    """

    def __init__(self, in_channels=1, output_heads: Dict[str, dict] = None, n_filter=32, **kwargs):
        super().__init__()
        self.output_heads = output_heads or {'default': {'channels': 1, 'activation': 'sigmoid'}}
        self.deep_supervision = False
        self.in_channels, self.n_filter = in_channels, n_filter
        widths = [n_filter * 2 ** i for i in range(5)]
        prev = in_channels
        for level in range(4):
            setattr(self, f'encode{2 * level + 1}', conv_block(2, prev, widths[level]))
            setattr(self, f'encode{2 * level + 2}', conv_block(2, widths[level], widths[level]))
            setattr(self, f'maxpool{level + 1}', nn.MaxPool2d(kernel_size=2, stride=2))
            prev = widths[level]
        self.middle_conv1 = conv_block(2, widths[3], widths[4])
        self.middle_conv2 = conv_block(2, widths[4], widths[4])
        for k in range(4):
            cin, cout = widths[4 - k], widths[3 - k]
            setattr(self, f'up{k + 1}', nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))
            setattr(self, f'decode{2 * k + 1}', conv_block(2, 2 * cout, cout))
            setattr(self, f'decode{2 * k + 2}', conv_block(2, cout, cout))
        self.output_layers = nn.ModuleDict()
        for name, config in self.output_heads.items():
            self.output_layers[name] = nn.Conv2d(n_filter, config['channels'], kernel_size=1, padding=0)

    def _engine_spec(self):
        heads = []
        for name, cfg in self.output_heads.items():
            act = cfg.get('activation')
            heads.append((name, cfg['channels'], act if act in ('sigmoid', 'tanh', 'relu') else None))
        return dict(kind='mo2d', n_filter=self.n_filter, in_channels=self.in_channels, heads=heads)

    @staticmethod
    def concat(x1, x2):
        if x1.shape == x2.shape:
            return torch.cat((x1, x2), 1)
        raise ValueError(f'Concatenation failed: wrong dimensions {x1.shape}, {x2.shape}')

    @staticmethod
    def apply_activation(x, activation):
        if activation == 'sigmoid':
            return torch.sigmoid(x)
        if activation == 'tanh':
            return torch.tanh(x)
        if activation == 'relu':
            return torch.relu(x)
        return x

    def _torch_features(self, x):
        skips = []
        for level in range(4):
            x = getattr(self, f'encode{2 * level + 2}')(getattr(self, f'encode{2 * level + 1}')(x))
            skips.append(x)
            x = getattr(self, f'maxpool{level + 1}')(x)
        x = self.middle_conv2(self.middle_conv1(x))
        for k in range(4):
            x = self.concat(getattr(self, f'up{k + 1}')(x), skips[3 - k])
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return x

    def forward(self, x):
        """Dict head name -> activated output, like multi_output_unet.py:88-134."""
        if self.training:
            d8 = self._torch_features(x)
            return {name: self.apply_activation(self.output_layers[name](d8), cfg.get('activation'))
                    for name, cfg in self.output_heads.items()}
        val = self._engine_forward(x, raw_logits=False)
        out, c0 = {}, 0
        for name, cfg in self.output_heads.items():
            out[name] = val[:, c0:c0 + cfg['channels']]
            c0 += cfg['channels']
        return out
