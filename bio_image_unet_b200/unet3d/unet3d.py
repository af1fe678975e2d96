"""3D U-Net, depth 3 (reference: unet3d/unet3d.py:6-99)."""
import torch
import torch.nn.functional as F
from torch import nn

from ..nn_base import EngineModule, conv_block


def body3d_channels(n_filter):
    """(cin, cout) of encode1..6 / middle_conv1..2 / decode1..6 and the up widths (unet3d/unet3d.py:24-49)."""
    h, n = n_filter // 2, n_filter
    enc = [(None, h), (h, n), (n, n), (n, 2 * n), (2 * n, 2 * n), (2 * n, 4 * n)]
    mid = [(4 * n, 4 * n), (4 * n, 8 * n)]
    dec = [(12 * n, 4 * n), (4 * n, 4 * n), (6 * n, 2 * n), (2 * n, 2 * n), (3 * n, n), (n, h)]
    ups = [8 * n, 4 * n, 2 * n]
    return enc, mid, dec, ups


class UNet3D(EngineModule):
    """3D U-Net for volume / time-consistent segmentation.

    n_filter : base width; use_interpolation : trilinear upsampling instead of transposed convolutions (the
    engine implements the default transposed-convolution variant).
    """

    def __init__(self, in_channels=1, out_channels=1, n_filter=16, use_interpolation=False):
        super().__init__()
        self.in_channels, self.out_channels, self.n_filter = in_channels, out_channels, n_filter
        self.use_interpolation = use_interpolation
        enc, mid, dec, ups = body3d_channels(n_filter)
        for i, (cin, cout) in enumerate(enc):
            setattr(self, f'encode{i + 1}', conv_block(3, in_channels if cin is None else cin, cout))
            if i % 2 == 1:
                setattr(self, f'maxpool{i // 2 + 1}', nn.MaxPool3d(kernel_size=2, stride=2))
        self.middle_conv1 = conv_block(3, *mid[0])
        self.middle_conv2 = conv_block(3, *mid[1])
        if not use_interpolation:
            for k, c in enumerate(ups):
                setattr(self, f'up{k + 1}', nn.ConvTranspose3d(c, c, kernel_size=2, stride=2))
        for i, (cin, cout) in enumerate(dec):
            setattr(self, f'decode{i + 1}', conv_block(3, cin, cout))
        self.final = nn.Conv3d(n_filter // 2, out_channels=out_channels, kernel_size=1, padding=0)

    def _engine_spec(self):
        return dict(kind='unet3d', n_filter=self.n_filter, in_channels=self.in_channels,
                    heads=[('', self.out_channels, 'sigmoid')], use_interpolation=self.use_interpolation)

    def _torch_forward(self, x):
        skips = []
        for level in range(3):
            x = getattr(self, f'encode{2 * level + 2}')(getattr(self, f'encode{2 * level + 1}')(x))
            skips.append(x)
            x = getattr(self, f'maxpool{level + 1}')(x)
        x = self.middle_conv2(self.middle_conv1(x))
        for k in range(3):
            if self.use_interpolation:
                x = F.interpolate(x, scale_factor=2, mode='trilinear', align_corners=False)
            else:
                x = getattr(self, f'up{k + 1}')(x)
            x = torch.cat((x, skips[2 - k]), 1)
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return self.final(x)

    def forward(self, x):
        logits = self._torch_forward(x) if self.training else self._engine_forward(x)
        return torch.sigmoid(logits), logits
