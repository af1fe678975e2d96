"""Parity gates of the reduced-precision modes.

north_star: fp32/TF32 within max-abs 1e-3 and bf16 within 1e-2 of the reference's fp32 forward "on identical
random-init weights" - asserted as such on PyTorch's default init. On stress-initialised nets (Kaiming weights,
randomised BatchNorm statistics, logits spread over the steep part of the sigmoid) no constant is meaningful, so the
engine is held to the REFERENCE'S OWN arithmetic at that precision, evaluated on the same weights and inputs
(oracle/yardstick.py): torch.autocast(bfloat16) for bf16, operand-rounded TF32 for tf32.

    engine_err <= YARD_FACTOR * yardstick + FLOOR

FLOOR covers the regime where both are at fp32 round-off (default init: errors of 1e-6).
"""
import numpy as np
import torch

from oracle import yardstick as ys

YARD_FACTOR = 1.5
FLOOR = 2e-5
FP32_TOL = 1e-4          # exact-fp32 mode (CUDA-core kernels): only the summation order differs from oneDNN
NORTH_STAR = {'fp32': 1e-3, 'tf32': 1e-3, 'bf16': 1e-2}


def errors(fn, precision, *args, ref=None):
    """(max, mean) abs error of the reference's reduced-precision arithmetic against its fp32 run."""
    with torch.no_grad():
        if ref is None:
            ref = ys._first(fn(*args))
        if precision == 'tf32':
            runs = [ys._first(ys.operand_rounded(fn, 'tf32', *args))]
        else:
            runs = [ys._first(ys.autocast_bf16(fn, *args)), ys._first(ys.operand_rounded(fn, 'bf16', *args))]
        d = [(r - ref).abs() for r in runs]
    return max(x.max().item() for x in d), max(x.mean().item() for x in d)


def bound(fn, precision, *args, ref=None):
    """(max-abs bound, mean-abs bound) for the engine's output in `precision`."""
    if precision == 'fp32':
        return FP32_TOL, FP32_TOL
    mx, mean = errors(fn, precision, *args, ref=ref)
    return YARD_FACTOR * mx + FLOOR, YARD_FACTOR * mean + FLOOR


def lsb_bound(fn, precision, *args, ref=None, scale=255.0):
    """Bound on |trunc(engine * 255) - trunc(reference * 255)| in LSB: one LSB for the truncation itself plus the float
    bound."""
    mx, _ = bound(fn, precision, *args, ref=ref)
    return 1 + int(np.ceil(mx * scale))


def check(val, ref, fn, precision, *args, what=''):
    """Assert the engine output `val` against the fp32 reference `ref` (tensors of equal shape); returns the errors."""
    val, ref = val.detach().float().cpu(), ref.detach().float().cpu()
    mx_b, mean_b = bound(fn, precision, *args, ref=ref)
    d = (val - ref).abs()
    mx, mean = d.max().item(), d.mean().item()
    assert mx <= mx_b, f'{what} {precision}: max-abs error {mx:.3e} > {mx_b:.3e} (= {YARD_FACTOR} x reference arithmetic + floor)'
    assert mean <= mean_b, f'{what} {precision}: mean-abs error {mean:.3e} > {mean_b:.3e}'
    return mx, mean, mx_b


def per_channel_bound(fn, precision, *args, ref=None):
    """(C,) tensor: bound on the max-abs error of every output channel (heads of different scale: linear heads of a
    stress net reach the hundreds, a sigmoid head stays in [0, 1])."""
    with torch.no_grad():
        if ref is None:
            ref = ys._first(fn(*args))
        red = [d for d in range(ref.dim()) if d != 1]
        scale = ref.abs().amax(dim=red).clamp(min=1.0)
        if precision == 'fp32':
            return FP32_TOL * scale
        if precision == 'tf32':
            runs = [ys._first(ys.operand_rounded(fn, 'tf32', *args))]
        else:
            runs = [ys._first(ys.autocast_bf16(fn, *args)), ys._first(ys.operand_rounded(fn, 'bf16', *args))]
        yard = torch.stack([(r - ref).abs().amax(dim=red) for r in runs]).amax(0)
    return YARD_FACTOR * yard + FLOOR * scale
