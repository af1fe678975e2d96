"""First-generation 2D U-Net (reference: unet/unet_v0.py:5-106): Conv-BatchNorm-ReLU blocks, skip connections
taken after the FIRST convolution of every level, and an extra 3x3 block ``decode9`` (n_filter -> 1) in front of
the 1x1 head. Same constructor, parameter names / shapes and forward contract; eval-mode CUDA forwards run on the
B200 engine."""
import torch
from torch import nn

from ..nn_base import EngineModule


def _conv_relu(in_channels, out_channels, kernel_size=3, dropout=0.0):
    """unet/unet_v0.py:56-63 (module indices 0 / 1 carry the parameters, like the reference's state_dict)."""
    return nn.Sequential(nn.Conv2d(kernel_size=kernel_size, in_channels=in_channels, out_channels=out_channels, padding=1),
                         nn.BatchNorm2d(out_channels), nn.ReLU(), nn.Dropout2d(dropout))


class Unet_v0(EngineModule):
    """U-Net (Falk et al., Nat Methods 16, 67-70 (2019)), first version of the package.

    Parameters
    ----------
    n_filter : int      base width (commonly 16, 32 or 64)
    **kwargs            ignored (``unet.Predict`` passes in_channels / out_channels, unet/predict.py:98-99)
    """

    def __init__(self, n_filter=32, **kwargs):
        super().__init__()
        self.n_filter = n_filter
        widths = [n_filter * 2 ** i for i in range(5)]
        prev = 1
        for level in range(4):
            setattr(self, f'encode{2 * level + 1}', _conv_relu(prev, widths[level]))
            setattr(self, f'encode{2 * level + 2}', _conv_relu(widths[level], widths[level]))
            setattr(self, f'maxpool{level + 1}', nn.MaxPool2d(kernel_size=2, stride=2))
            prev = widths[level]
        self.middle_conv1 = _conv_relu(widths[3], widths[4])
        self.middle_conv2 = _conv_relu(widths[4], widths[4], dropout=0.5)
        for k in range(4):
            cin, cout = widths[4 - k], widths[3 - k]
            setattr(self, f'up{k + 1}', nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))
            setattr(self, f'decode{2 * k + 1}', _conv_relu(2 * cout, cout))
            setattr(self, f'decode{2 * k + 2}', _conv_relu(cout, cout))
        self.decode9 = _conv_relu(n_filter, 1)
        self.final = nn.Sequential(nn.Conv2d(1, 1, kernel_size=1, padding=0))

    def _engine_spec(self):
        return dict(kind='unet2d_v0', n_filter=self.n_filter, in_channels=1, heads=[('', 1, 'sigmoid')])

    @staticmethod
    def concat(x1, x2):
        if x1.shape == x2.shape:
            return torch.cat((x1, x2), 1)
        print(x1.shape, x2.shape)
        raise ValueError('concatenation failed: wrong dimensions')

    def _torch_forward(self, x):
        skips = []
        for level in range(4):
            e = getattr(self, f'encode{2 * level + 1}')(x)
            skips.append(e)                                     # e1 / e3 / e5 / e7 (unet/unet_v0.py:91-103)
            x = getattr(self, f'maxpool{level + 1}')(getattr(self, f'encode{2 * level + 2}')(e))
        x = self.middle_conv2(self.middle_conv1(x))
        for k in range(4):
            x = self.concat(getattr(self, f'up{k + 1}')(x), skips[3 - k])
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(x))
        return self.final(self.decode9(x))

    def forward(self, x):
        """Returns (sigmoid(logits), logits) like unet/unet_v0.py:106."""
        logits = self._torch_forward(x) if self.training else self._engine_forward(x)
        return torch.sigmoid(logits), logits
