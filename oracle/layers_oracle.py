"""numpy restatement of the reference layers on the path (TEST INFRASTRUCTURE ONLY).

Follows:
  p1/lora.py:13-54                   LoRALayer
  p1/lora.py:56-150                  SPLinearWithLoRA
  p1/switchable_batchnorm.py:7-109   SwitchableLayerNorm
Backward formulas are what torch autograd derives for those forwards with the
straight-through estimators of p1/quantization_methods.py:25-28, 82-90.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .quant_oracle import F32, QuantizerState, fake_quantize, ste_backward


def _mm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.matmul(a.astype(F32, copy=False), b.astype(F32, copy=False))


def lora_forward(x, lora_A, lora_B, q_A: Optional[QuantizerState],
                 q_B: Optional[QuantizerState], scaling: float, out_features: int):
    """p1/lora.py:45-54.  Disabled adapters (bits>=32 or rank<=0) return zeros."""
    if q_A is None or q_B is None or scaling == 0:
        return np.zeros(x.shape[:-1] + (out_features,), dtype=F32)
    a_q = fake_quantize(q_A, lora_A)
    b_q = fake_quantize(q_B, lora_B)
    return (_mm(_mm(x, a_q), b_q) * F32(scaling)).astype(F32)


def sp_linear_forward(x, weight, bias, bits: int,
                      q_in: Optional[QuantizerState], q_w: Optional[QuantizerState],
                      lora: Optional[Dict] = None, calibration_mode: bool = False):
    """p1/lora.py:127-150.

    ``lora`` is ``{'A','B','q_A','q_B','scaling'}`` for the active bit-width.
    In collecting mode ``q_in`` records statistics and passes x through
    (p1/quantization.py:214-216), exactly as the reference's calibration pass.
    """
    if bits >= 32:
        return _mm(x, weight.T) + bias
    x_q = fake_quantize(q_in, x)
    w_q = fake_quantize(q_w, weight)
    base = _mm(x_q, w_q.T) + bias
    if calibration_mode or lora is None:
        return base.astype(F32)
    lo = lora_forward(x, lora["A"], lora["B"], lora["q_A"], lora["q_B"],
                      lora["scaling"], weight.shape[0])
    return (base + lo).astype(F32)


def sp_linear_backward(grad_y, x, weight, bits: int,
                       q_in: Optional[QuantizerState], q_w: Optional[QuantizerState],
                       lora: Optional[Dict] = None):
    """Gradients of sp_linear_forward w.r.t. x, weight, bias, lora A/B under the STE.

    x / grad_y are flattened to 2-D [M, K] / [M, N].
    """
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K).astype(F32)
    g2 = grad_y.reshape(-1, N).astype(F32)
    out = {}
    if bits >= 32:
        out["x"] = _mm(g2, weight).reshape(x.shape)
        out["weight"] = _mm(g2.T, x2)
        out["bias"] = g2.sum(axis=0)
        return out
    x_q = fake_quantize(q_in, x2.reshape(x.shape)).reshape(-1, K)
    w_q = fake_quantize(q_w, weight)
    gx = ste_backward(_mm(g2, w_q), q_in.quantizer_type)
    out["weight"] = ste_backward(_mm(g2.T, x_q), q_w.quantizer_type)
    out["bias"] = g2.sum(axis=0)
    if lora is not None and lora.get("q_A") is not None and lora["scaling"] != 0:
        s = F32(lora["scaling"])
        a_q = fake_quantize(lora["q_A"], lora["A"])
        b_q = fake_quantize(lora["q_B"], lora["B"])
        t = _mm(x2, a_q)                       # [M, r]
        gs = (g2 * s).astype(F32)
        out["lora_B"] = ste_backward(_mm(t.T, gs), lora["q_B"].quantizer_type)
        gt = _mm(gs, b_q.T)                    # [M, r]
        out["lora_A"] = ste_backward(_mm(x2.T, gt), lora["q_A"].quantizer_type)
        gx = gx + _mm(gt, a_q.T)
    out["x"] = gx.reshape(x.shape).astype(F32)
    return out


def switchable_layernorm_forward(x, weight, bias, eps: float = 1e-5):
    """p1/switchable_batchnorm.py:102-109 (biased variance, eps inside the sqrt).

    Returns (y, mean, rstd); weight/bias are the pair of the current precision.
    """
    x = x.astype(F32, copy=False)
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    var = np.mean((x - mean) ** 2, axis=-1, keepdims=True, dtype=F32)
    rstd = (F32(1) / np.sqrt(var + F32(eps))).astype(F32)
    y = weight * ((x - mean) * rstd) + bias
    return y.astype(F32), mean, rstd


def switchable_layernorm_backward(grad_y, x, weight, mean, rstd):
    """dx, dweight, dbias of the LayerNorm above."""
    C = x.shape[-1]
    xh = ((x - mean) * rstd).astype(F32)
    g2 = grad_y.reshape(-1, C)
    dweight = (g2 * xh.reshape(-1, C)).sum(axis=0).astype(F32)
    dbias = g2.sum(axis=0).astype(F32)
    gxh = (grad_y * weight).astype(F32)
    m1 = gxh.mean(axis=-1, keepdims=True, dtype=F32)
    m2 = (gxh * xh).mean(axis=-1, keepdims=True, dtype=F32)
    dx = ((gxh - m1 - xh * m2) * rstd).astype(F32)
    return dx, dweight, dbias
