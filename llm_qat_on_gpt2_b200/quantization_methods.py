"""Min-max and log fake quantisation on the B200 kernels, with the reference's call signatures.

Drop-in for part1_switchable_precision/quantization_methods.py of the reference
(`apply_minmax_quantization`, `apply_log_quantization`, :92-98; the two autograd Functions,
:5-90).  Forward is one launch of `spq_fake_quantize` (csrc/spq_quantize.cu) instead of 4 / ~25
eager kernels; backward is the straight-through estimator (`spq_ste_backward`): identity for
min-max, clamp to [-10, 10] for log, no gradient to scale / zero-point.
"""
from __future__ import annotations

import torch

from . import _lib


def _view2d(x: torch.Tensor, param: torch.Tensor):
    """Find the [rows, cols] view of x and the broadcast mode under which `param` (the
    scale-shaped tensor) lines up with it.  Supports the layouts the reference produces:
    scalar, per-last-dim ([1,..,1,C]) and per-first-dim ([N,1,..,1])."""
    n = param.numel()
    if x.dim() == 0:
        raise ValueError("cannot quantise a 0-d tensor")
    last = x.shape[-1]
    if n == 1:
        return x.reshape(-1, last), _lib.PER_TENSOR
    pshape = tuple(param.shape)
    if n == last and pshape[-1] == last:
        return x.reshape(-1, last), _lib.PER_COL
    if n == x.shape[0] and len(pshape) >= 1 and pshape[0] == x.shape[0] and (param.dim() == x.dim() or param.dim() == 1):
        return x.reshape(x.shape[0], -1), _lib.PER_ROW
    raise NotImplementedError(
        f"scale of shape {pshape} does not broadcast per-tensor, per-first-dim or per-last-dim over "
        f"an input of shape {tuple(x.shape)}")


def _prep(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("fake quantisation runs on the CUDA kernels only (no CPU fallback); "
                           f"got a tensor on {x.device}")
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


def quantize_forward(x, scale, zero_point, num_bits, symmetric, qtype, want_codes=False):
    """Run the quantise kernel; returns dequant (and int32 codes, int8 sign when asked)."""
    xc = _prep(x)
    sc = _prep(scale)
    zp = _prep(zero_point)
    x2d, bcast = _view2d(xc, sc)
    if zp.numel() != sc.numel():
        zp = zp.expand_as(sc).contiguous()
    out = torch.empty_like(x2d)
    codes = torch.empty(x2d.shape, dtype=torch.int32, device=x2d.device) if want_codes else None
    sign = torch.empty(x2d.shape, dtype=torch.int8, device=x2d.device) if (want_codes and qtype == _lib.LOG) else None
    if x2d.numel():
        _lib.fake_quantize(x2d, sc, zp, bcast, qtype, int(num_bits), bool(symmetric), dequant=out, codes=codes, sign=sign)
    out = out.view(x.shape)
    if want_codes:
        return out, codes.view(x.shape), (None if sign is None else sign.view(x.shape))
    return out


class MinMaxQuantizationFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, scale, zero_point, num_bits, symmetric):
        return quantize_forward(input, scale, zero_point, num_bits, symmetric, _lib.MINMAX)

    @staticmethod
    def backward(ctx, grad_output):
        # identity STE (reference: grad_output.clone(), :25-28) -- nothing to launch
        return grad_output, None, None, None, None


class LogQuantizationFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, log_min, log_range, num_bits, symmetric):
        # the kernel takes (scale, zero_point) = (log_range, log_min), as the module stores them
        return quantize_forward(input, log_range, log_min, num_bits, symmetric, _lib.LOG)

    @staticmethod
    def backward(ctx, grad_output):
        return _lib.ste_backward(grad_output, _lib.LOG), None, None, None, None


def apply_minmax_quantization(x, scale, zero_point, num_bits, symmetric=True):
    return MinMaxQuantizationFunction.apply(x, scale, zero_point, num_bits, symmetric)


def apply_log_quantization(x, log_min, log_range, num_bits, symmetric=True):
    return LogQuantizationFunction.apply(x, log_min, log_range, num_bits, symmetric)


def quantize_codes(x, scale, zero_point, num_bits, symmetric=True, quantizer_type="minmax"):
    """Integer view of the quantiser (not in the reference API; used for parity tests and for
    exporting real integer weights): returns (dequant, int32 codes, int8 sign-or-None)."""
    return quantize_forward(x, scale, zero_point, num_bits, symmetric, _lib.QTYPE[quantizer_type], want_codes=True)
