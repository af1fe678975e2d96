"""2D U-Net with attention-gated skip connections (reference: unet/attention_unet.py:5-181). Same constructor,
parameter names / shapes and forward contract; eval-mode CUDA forwards run on the B200 engine, where each
AttentionBlock is one tcgen05 GEMM over the level's concat buffer (gate and skip 1x1 convs with their BatchNorms
folded, ReLU, psi as the fused 1x1 head) followed by an in-place scaling of the skip half."""
import torch
from torch import nn

from ..nn_base import EngineModule, conv_block


class AttentionBlock(nn.Module):
    """Attention gate (unet/attention_unet.py:111-181): psi = sigmoid(BN(conv(relu(BN(conv(gate)) + BN(conv(skip))))))
    and out = skip * psi.

    Parameters
    ----------
    F_g : int   channels of the gating signal (the up-sampled decoder tensor)
    F_l : int   channels of the encoder tensor arriving through the skip connection
    n_coefficients : int   width of the intermediate attention map
    """

    def __init__(self, F_g, F_l, n_coefficients):
        super().__init__()
        self.W_gate = nn.Sequential(nn.Conv2d(F_g, n_coefficients, kernel_size=1, stride=1, padding=0, bias=True),
                                    nn.BatchNorm2d(n_coefficients))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, n_coefficients, kernel_size=1, stride=1, padding=0, bias=True),
                                 nn.BatchNorm2d(n_coefficients))
        self.psi = nn.Sequential(nn.Conv2d(n_coefficients, 1, kernel_size=1, stride=1, padding=0, bias=True),
                                 nn.BatchNorm2d(1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)

    def forward(self, gate, skip_connection):
        psi = self.psi(self.relu(self.W_gate(gate) + self.W_x(skip_connection)))
        return skip_connection * psi


class AttentionUnet(EngineModule):
    """U-Net with attention mechanism during decoding.

    Parameters
    ----------
    in_channels, out_channels : int
    n_filter : int      base width (commonly 16, 32 or 64; must be even)
    dilation : int      the engine implements dilation 1 (what ``unet.Predict`` instantiates, unet/predict.py:98-99)
    """

    def __init__(self, in_channels=1, out_channels=1, n_filter=32, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.n_filter, self.dilation = in_channels, out_channels, n_filter, dilation
        widths = [n_filter * 2 ** i for i in range(5)]
        prev = in_channels
        for level in range(4):
            setattr(self, f'encode{2 * level + 1}', conv_block(2, prev, widths[level], dilation=dilation))
            setattr(self, f'encode{2 * level + 2}', conv_block(2, widths[level], widths[level], dilation=dilation))
            setattr(self, f'maxpool{level + 1}', nn.MaxPool2d(kernel_size=2, stride=2))
            prev = widths[level]
        self.middle_conv1 = conv_block(2, widths[3], widths[4], dilation=dilation)
        self.middle_conv2 = conv_block(2, widths[4], widths[4], dilation=dilation)
        for k in range(4):
            cin, cout = widths[4 - k], widths[3 - k]
            setattr(self, f'up{k + 1}', nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))
            setattr(self, f'attention{k + 1}', AttentionBlock(cout, cout, n_coefficients=cout // 2))
            setattr(self, f'decode{2 * k + 1}', conv_block(2, 2 * cout, cout))
            setattr(self, f'decode{2 * k + 2}', conv_block(2, cout, cout))
        self.final = nn.Sequential(nn.Conv2d(n_filter, out_channels, kernel_size=1, padding=0))

    def _engine_spec(self):
        if self.dilation != 1:
            raise NotImplementedError('the B200 engine implements dilation=1 (what unet.Predict instantiates)')
        return dict(kind='attunet2d', n_filter=self.n_filter, in_channels=self.in_channels,
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
            u = getattr(self, f'up{k + 1}')(x)
            a = getattr(self, f'attention{k + 1}')(gate=u, skip_connection=skips[3 - k])
            x = getattr(self, f'decode{2 * k + 2}')(getattr(self, f'decode{2 * k + 1}')(self.concat(a, u)))
        return self.final(x)

    def forward(self, x):
        """Returns (sigmoid(logits), logits) like unet/attention_unet.py:108."""
        logits = self._torch_forward(x) if self.training else self._engine_forward(x)
        return torch.sigmoid(logits), logits
