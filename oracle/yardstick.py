"""TEST INFRASTRUCTURE ONLY — reduced-precision yardsticks for the network parity gates.

A reduced-precision mode cannot be held to a constant: how far bf16 / TF32 arithmetic moves the sigmoid depends on
the weights. The gates therefore compare the engine's error with what the REFERENCE's arithmetic
does at the same precision, on the same weights and input:

* ``autocast_bf16(fn, ...)``   — the oracle forward under ``torch.autocast(dtype=torch.bfloat16)``, i.e. what a user
  of the reference gets from PyTorch's own mixed precision (the conv / BN modules of unet/unet.py:54-60 under
  autocast).
* ``operand_rounded(fn, ..., 'tf32' | 'bf16')`` — the oracle forward in fp32 with the two operands of every
  convolution / transposed convolution rounded to the format first (fp32 accumulation, fp32 BatchNorm): the
  arithmetic of a TF32 / bf16 tensor-core GEMM with fp32 epilogue, which is also what cuDNN does for the reference
  when ``torch.backends.cudnn.allow_tf32`` is on (PyTorch's default for convolutions).

``yardstick(...)`` returns the max-abs error of such a run against the plain fp32 oracle run.
"""
import torch
from torch.overrides import TorchFunctionMode

_CONVS = {torch.conv2d, torch.conv3d, torch.conv_transpose2d, torch.conv_transpose3d,
          torch.nn.functional.conv2d, torch.nn.functional.conv3d, torch.nn.functional.conv_transpose2d,
          torch.nn.functional.conv_transpose3d}


def round_tf32(t):
    """Round-to-nearest (ties away, like cvt.rna.tf32.f32) to TF32's 10 explicit mantissa bits."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _RoundConvOperands(TorchFunctionMode):
    def __init__(self, q):
        super().__init__()
        self.q = q

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _CONVS:
            args = (self.q(args[0]), self.q(args[1]), *args[2:])
        return func(*args, **kwargs)


def operand_rounded(fn, fmt, *args, **kwargs):
    """fn(*args) with every convolution's input and weight rounded to `fmt` ('tf32' | 'bf16'), fp32 accumulate."""
    q = {'tf32': round_tf32, 'bf16': round_bf16}[fmt]
    with torch.no_grad(), _RoundConvOperands(q):
        return fn(*args, **kwargs)


def autocast_bf16(fn, *args, **kwargs):
    """fn(*args) under torch.autocast(bfloat16) on the device of the first tensor argument."""
    dev = next((a.device.type for a in args if torch.is_tensor(a)), 'cpu')
    with torch.no_grad(), torch.autocast(dev, dtype=torch.bfloat16):
        return fn(*args, **kwargs)


def _first(out):
    """The tensor the parity gate looks at: the sigmoid output (first element of the reference's tuple), or a dict
    of heads concatenated along the channel axis."""
    if isinstance(out, dict):
        return torch.cat([v.float().reshape(v.shape[0], -1, *v.shape[-3:] if v.dim() == 5 else v.shape[-2:])
                          for v in out.values()], 1)
    if isinstance(out, (tuple, list)):
        return out[0].float()
    return out.float()


def yardstick(fn, precision, *args, ref=None, **kwargs):
    """max |reduced-precision reference - fp32 reference| for `precision` in {'tf32', 'bf16'}.

    bf16: the larger of the reference's autocast run and the operand-rounded emulation (autocast rounds more often,
    but on some nets its errors cancel; either is 'what the reference's arithmetic gives at bf16').
    tf32: the operand-rounded emulation."""
    with torch.no_grad():
        if ref is None:
            ref = _first(fn(*args, **kwargs))
        if precision == 'tf32':
            return (_first(operand_rounded(fn, 'tf32', *args, **kwargs)) - ref).abs().max().item()
        if precision == 'bf16':
            a = (_first(autocast_bf16(fn, *args, **kwargs)) - ref).abs().max().item()
            b = (_first(operand_rounded(fn, 'bf16', *args, **kwargs)) - ref).abs().max().item()
            return max(a, b)
    raise ValueError(precision)
