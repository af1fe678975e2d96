"""Siamese U-Net (reference: siam_unet/siam_unet.py:7-148): shared-weight twin encoder for frames T and T-1,
joined at the bottleneck; the decoder uses the current frame's skips only."""
import logging

import torch
import torch.nn.functional as F
from torch import nn

from ..nn_base import EngineModule, conv_block


class Siam_UNet(EngineModule):
    """
    Parameters
    ----------
    n_filter : int
        Base width.
    mode : str
        How T-1 and T are combined at the bottleneck: 'concat' (cat + conv_concat), 'max', 'corr'
        (depth-wise cross-correlation) or 'control' (ignore T-1); all four run on the engine.
    """

    def __init__(self, n_filter=32, mode='concat'):
        super().__init__()
        self.mode, self.n_filter = mode, n_filter
        widths = [n_filter * 2 ** i for i in range(5)]
        prev = 1
        for level in range(4):
            setattr(self, f'encode{2 * level + 1}', conv_block(2, prev, widths[level]))
            setattr(self, f'encode{2 * level + 2}', conv_block(2, widths[level], widths[level]))
            setattr(self, f'maxpool{level + 1}', nn.MaxPool2d(kernel_size=2, stride=2))
            prev = widths[level]
        if mode == 'concat':
            self.conv_concat = conv_block(2, 2 * widths[3], widths[3])
        self.middle_conv1 = conv_block(2, widths[3], widths[4])
        self.middle_conv2 = conv_block(2, widths[4], widths[4])
        for k in range(4):
            cin, cout = widths[4 - k], widths[3 - k]
            setattr(self, f'up{k + 1}', nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))
            setattr(self, f'decode{2 * k + 1}', conv_block(2, 2 * cout, cout))
            setattr(self, f'decode{2 * k + 2}', conv_block(2, cout, cout))
        self.final = nn.Sequential(nn.Conv2d(n_filter, 1, kernel_size=1, padding=0))

    def _engine_spec(self):
        if self.mode not in ('concat', 'max', 'control', 'corr'):
            raise NotImplementedError('Unknown mode: {}'.format(self.mode))
        return dict(kind='siam2d', n_filter=self.n_filter, in_channels=1, heads=[('', 1, 'sigmoid')],
                    siam_mode=self.mode)

    @staticmethod
    def concat(x1, x2):
        if x1.shape == x2.shape:
            return torch.cat((x1, x2), 1)
        logging.critical(f'Shapes: {x1.shape}, {x2.shape}')
        raise ValueError('concatenation failed: wrong dimensions')

    def _encode(self, x):
        skips = []
        for level in range(4):
            x = getattr(self, f'encode{2 * level + 2}')(getattr(self, f'encode{2 * level + 1}')(x))
            skips.append(x)
            x = getattr(self, f'maxpool{min(level + 1, 2)}')(x)      # the reference reuses maxpool2 (:95,98)
        return skips, x

    def _torch_forward(self, x, prev_x):
        skips, m4 = self._encode(x)
        _, mm4 = self._encode(prev_x)
        if self.mode == 'corr':
            b, c = mm4.size(0), mm4.size(1)
            out = F.conv2d(m4.reshape(1, b * c, *m4.shape[2:]), mm4.reshape(b * c, 1, *mm4.shape[2:]), groups=b * c,
                           padding='same')
            join = out.view(b, c, *out.shape[2:])
        elif self.mode == 'max':
            join = torch.maximum(m4, mm4)
        elif self.mode == 'concat':
            join = self.conv_concat(self.concat(m4, mm4))
        elif self.mode == 'control':
            join = m4
        else:
            raise NotImplementedError('Unknown mode: {}'.format(self.mode))
        x = self.middle_conv2(self.middle_conv1(join))
        for k in range(4):
            x = self.concat(getattr(self, f'up{k + 1}')(x), skips[3 - k])
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return self.final(x)

    def forward(self, x, prev_x):
        logits = self._torch_forward(x, prev_x) if self.training else self._engine_forward(x, prev_x)
        return torch.sigmoid(logits), logits
