"""2D U-Net, depth 4 (reference: unet/unet.py:5-104). Same constructor, parameter names and forward contract."""
import torch
from torch import nn

from ..nn_base import EngineModule, conv_block


class Unet(EngineModule):
    """U-Net for semantic segmentation (Falk et al., Nat Methods 2019).

    Parameters
    ----------
    in_channels, out_channels : int
    n_filter : int
        Base width (channels double per level: n, 2n, 4n, 8n, 16n).
    dilation : int
        Dilation (and padding) of the encoder / bottleneck convolutions (unet/unet.py:20-35). The engine implements
        dilation 1, which is also what ``Predict`` always instantiates (unet/predict.py:98-99).
    """

    def __init__(self, in_channels=1, out_channels=1, n_filter=32, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.n_filter, self.dilation = in_channels, out_channels, n_filter, dilation
        widths = [n_filter * 2 ** i for i in range(5)]
        prev = in_channels
        for level in range(4):                                   # encode1..8 + maxpool1..4
            setattr(self, f'encode{2 * level + 1}', conv_block(2, prev, widths[level], dilation=dilation))
            setattr(self, f'encode{2 * level + 2}', conv_block(2, widths[level], widths[level], dilation=dilation))
            setattr(self, f'maxpool{level + 1}', nn.MaxPool2d(kernel_size=2, stride=2))
            prev = widths[level]
        self.middle_conv1 = conv_block(2, widths[3], widths[4], dilation=dilation)
        self.middle_conv2 = conv_block(2, widths[4], widths[4], dilation=dilation)
        for k in range(4):                                       # up1..4, decode1..8
            cin, cout = widths[4 - k], widths[3 - k]
            setattr(self, f'up{k + 1}', nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))
            setattr(self, f'decode{2 * k + 1}', conv_block(2, 2 * cout, cout))
            setattr(self, f'decode{2 * k + 2}', conv_block(2, cout, cout))
        self.final = nn.Sequential(nn.Conv2d(n_filter, out_channels, kernel_size=1, padding=0))

    def _engine_spec(self):
        if self.dilation != 1:
            raise NotImplementedError('the B200 engine implements dilation=1 (what unet.Predict instantiates)')
        return dict(kind='unet2d', n_filter=self.n_filter, in_channels=self.in_channels,
                    heads=[('', self.out_channels, 'sigmoid')])

    @staticmethod
    def concat(x1, x2):
        if x1.shape == x2.shape:
            return torch.cat((x1, x2), 1)
        print(x1.shape, x2.shape)
        raise ValueError('concatenation failed: wrong dimensions')

    def _torch_forward(self, x):
        skips = []
        for level in range(4):
            x = getattr(self, f'encode{2 * level + 2}')(getattr(self, f'encode{2 * level + 1}')(x))
            skips.append(x)
            x = getattr(self, f'maxpool{level + 1}')(x)
        x = self.middle_conv2(self.middle_conv1(x))
        for k in range(4):
            x = self.concat(getattr(self, f'up{k + 1}')(x), skips[3 - k])
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return self.final(x)

    def forward(self, x):
        """Returns (sigmoid(logits), logits) like unet/unet.py:104."""
        logits = self._torch_forward(x) if self.training else self._engine_forward(x)
        return torch.sigmoid(logits), logits
