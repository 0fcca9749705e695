"""numpy restatement of the reference fake-quantiser (TEST INFRASTRUCTURE ONLY).

Follows, function by function:
  p1/quantization.py:15-239          LearnableFakeQuantize (state machine, calibration)
  p1/quantization_methods.py:5-98    MinMax / Log quantisation autograd Functions

All arithmetic is float32 with IEEE semantics, in the reference's operation
order; nothing is algebraically simplified, because the integer codes must be
reproduced bit for bit (division is a true division, rounding is half-to-even,
``a*b+c`` is two roundings).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import numpy as np

F32 = np.float32
# p1/quantization_methods.py:35 -- the log quantiser ignores the module eps and
# hard-codes 1e-5.
LOG_EPS = F32(1e-5)


def log2_cr(a: np.ndarray) -> np.ndarray:
    """float32 log2 with (essentially) correct rounding: via float64.

    torch-CPU's log2 agrees with this on 99.987 % of inputs (measured in the
    build container); the CUDA kernels evaluate the same definition.
    """
    return np.log2(np.asarray(a, dtype=np.float64)).astype(F32)


def exp2_cr(a: np.ndarray) -> np.ndarray:
    return np.exp2(np.asarray(a, dtype=np.float64)).astype(F32)


@dataclass
class QuantizerState:
    """Host state of one LearnableFakeQuantize (p1/quantization.py:16-38)."""

    num_bits: int
    channel_dim: Optional[int] = 0
    quantizer_type: str = "minmax"
    eps: float = 1e-5
    symmetric: bool = True
    per_channel: bool = True
    is_input: bool = False
    scale: np.ndarray = field(default_factory=lambda: np.ones(1, F32))
    zero_point: np.ndarray = field(default_factory=lambda: np.zeros(1, F32))
    running_min: np.ndarray = field(default_factory=lambda: np.zeros(1, F32))
    running_max: np.ndarray = field(default_factory=lambda: np.zeros(1, F32))
    calibrated: bool = False
    collecting_stats: bool = False
    num_batches_collected: int = 0
    temp_min: Optional[np.ndarray] = None
    temp_max: Optional[np.ndarray] = None

    def __post_init__(self):
        # p1/quantization.py:19,22
        self.num_bits = max(1, min(int(self.num_bits), 32))
        if not self.per_channel:
            self.channel_dim = None

    # p1/quantization.py:96-102
    def start_calibration(self):
        self.collecting_stats = True
        self.calibrated = False
        self.num_batches_collected = 0
        self.temp_min = None
        self.temp_max = None

    # p1/quantization.py:77-85
    def set_num_bits(self, value: int):
        old = self.num_bits
        self.num_bits = max(1, min(int(value), 32))
        if old != self.num_bits:
            self.calibrated = False


# --------------------------------------------------------------------------
# calibration
# --------------------------------------------------------------------------

def _reduction_dims(q: QuantizerState, ndim: int):
    """p1/quantization.py:141-150."""
    dims = list(range(ndim))
    if q.per_channel and q.channel_dim is not None:
        keep = q.channel_dim if q.channel_dim >= 0 else ndim + q.channel_dim
        if keep in dims:
            dims.remove(keep)
    return tuple(dims)


def reduce_min_max(x: np.ndarray, dims) -> Tuple[np.ndarray, np.ndarray]:
    """p1/quantization.py:152-162 (keepdim reductions; NaN propagates as in torch)."""
    if not dims:
        return x, x
    return (np.min(x, axis=tuple(dims), keepdims=True),
            np.max(x, axis=tuple(dims), keepdims=True))


def _default_shape(q: QuantizerState, x: np.ndarray, value) -> np.ndarray:
    """p1/quantization.py:164-172.  Note the quirk: the *channel* dim is the one
    set to 1, every other dim keeps x's extent."""
    if q.per_channel and q.channel_dim is not None:
        shape = list(x.shape)
        keep = q.channel_dim if q.channel_dim >= 0 else len(shape) + q.channel_dim
        shape[keep] = 1
        return np.full(shape, value, dtype=F32)
    return np.asarray(value, dtype=F32)


def collect_statistics(q: QuantizerState, x: np.ndarray,
                       log2_fn: Callable = log2_cr) -> None:
    """One calibration batch (p1/quantization.py:174-209)."""
    x = np.asarray(x, dtype=F32)
    eps = F32(q.eps)
    if q.quantizer_type == "log":
        ax = np.abs(x)
        if np.any(ax > eps):
            lx = log2_fn(np.maximum(ax, eps))
            mn, mx = reduce_min_max(lx, _reduction_dims(q, lx.ndim))
            if q.num_batches_collected == 0:
                q.temp_min, q.temp_max = mn.copy(), mx.copy()
            else:
                q.temp_min = np.minimum(q.temp_min, mn)
                q.temp_max = np.maximum(q.temp_max, mx)
        elif q.num_batches_collected == 0:
            log_eps = log2_fn(np.asarray(eps, dtype=F32))
            q.temp_min = _default_shape(q, x, log_eps)
            q.temp_max = _default_shape(q, x, log_eps)
    else:
        mn, mx = reduce_min_max(x, _reduction_dims(q, x.ndim))
        if q.num_batches_collected == 0:
            q.temp_min, q.temp_max = mn.copy(), mx.copy()
        else:
            q.temp_min = np.minimum(q.temp_min, mn)
            q.temp_max = np.maximum(q.temp_max, mx)
    q.num_batches_collected += 1


def finish_calibration(q: QuantizerState) -> None:
    """p1/quantization.py:104-139."""
    if q.num_batches_collected > 0 and q.temp_min is not None:
        q.running_min = np.array(q.temp_min, dtype=F32, copy=True)
        q.running_max = np.array(q.temp_max, dtype=F32, copy=True)
        if q.quantizer_type == "log":
            q.zero_point = q.running_min.copy()                      # log_min
            q.scale = (q.running_max - q.running_min).astype(F32)     # log_range
        elif q.symmetric:
            amax = np.maximum(np.abs(q.running_min), np.abs(q.running_max))
            amax = np.maximum(amax, F32(q.eps))
            q.scale = (amax / F32(2 ** (q.num_bits - 1) - 1)).astype(F32)
            q.zero_point = np.zeros_like(amax)
        else:
            rng = np.maximum(q.running_max - q.running_min, F32(q.eps))
            q.scale = (rng / F32(2 ** q.num_bits - 1)).astype(F32)
            q.zero_point = np.rint((-q.running_min) / q.scale).astype(F32)
        q.calibrated = True
        q.temp_min = q.temp_max = None
    q.collecting_stats = False


# --------------------------------------------------------------------------
# quantise / dequantise
# --------------------------------------------------------------------------

def minmax_quantize(x, scale, zero_point, num_bits: int, symmetric: bool = True):
    """p1/quantization_methods.py:8-22.  Returns (dequantised float32, integer codes)."""
    x = np.asarray(x, dtype=F32)
    scale = np.asarray(scale, dtype=F32)
    zero_point = np.asarray(zero_point, dtype=F32)
    if symmetric:
        n = F32(2 ** (num_bits - 1) - 1)
        q = np.clip(np.rint(x / scale), -n, n)
        out = q * scale
    else:
        hi = F32(2 ** num_bits - 1)
        q = np.clip(np.rint((x / scale) + zero_point), F32(0), hi)
        out = (q - zero_point) * scale
    return out.astype(F32), q.astype(np.int32)


def log_quantize(x, log_min, log_range, num_bits: int, symmetric: bool = True,
                 log2_fn: Callable = log2_cr, exp2_fn: Callable = exp2_cr):
    """p1/quantization_methods.py:33-79.

    Returns (dequantised float32, level index L, sign in {-1,0,1}, zero mask).
    ``L`` is the integer the reference rounds to (``quantized`` after the clamp
    at :55/:60); it and the zero mask are the parts that must match exactly.
    """
    x = np.asarray(x, dtype=F32)
    log_min = np.asarray(log_min, dtype=F32)
    log_range = np.asarray(log_range, dtype=F32)
    zero_mask = np.abs(x) < LOG_EPS
    sgn = np.sign(x).astype(F32)
    ax = np.maximum(np.abs(x), LOG_EPS)
    lx = log2_fn(ax)
    ln = (lx - log_min) / np.maximum(log_range, LOG_EPS)
    ln = np.clip(ln, F32(0), F32(1))
    if symmetric:
        n = 2 ** (num_bits - 1) - 1
        centered = ln - F32(0.5)
        level = np.rint((centered * F32(2)) * F32(n))
        level = np.clip(level, F32(-n), F32(n))
        qv = ((level / F32(2 * n)) + F32(0.5)) * F32(2 ** num_bits - 1)
        qn = qv / F32(2 ** num_bits - 1)
    else:
        n = 2 ** num_bits - 1
        level = np.clip(np.rint(ln * F32(n)), F32(0), F32(n))
        qn = level / F32(n)
    x_hat = (qn * log_range).astype(F32) + log_min
    mag = exp2_fn(x_hat)
    out = np.where(zero_mask, F32(0), (mag * sgn).astype(F32)).astype(F32)
    return out, level.astype(np.int32), sgn.astype(np.int8), zero_mask


def _chunked(fn, x, params, threads):
    """Evaluate the elementwise fn(x, *params) over row chunks on a thread pool (numpy ufuncs
    release the GIL).  Bit-identical to the single call; only used to make the CPU baseline use
    the host cores the way the reference's torch-CPU kernels do."""
    shape = x.shape
    last = shape[-1]
    per_last = all(p.size == 1 or (p.size == last and p.shape[-1] == last) for p in params)
    per_first = all(p.size == 1 or (p.ndim == x.ndim and p.shape[0] == shape[0] and p.size == shape[0]) for p in params)
    if per_last:
        x2 = x.reshape(-1, last)
        ps = [p.reshape(1, -1) if p.size > 1 else p.reshape(1, 1) for p in params]
        slicer = lambda p, a, b: p
    elif per_first:
        x2 = x.reshape(shape[0], -1)
        ps = [p.reshape(-1, 1) if p.size > 1 else p.reshape(1, 1) for p in params]
        slicer = lambda p, a, b: p[a:b] if p.shape[0] > 1 else p
    else:
        return fn(x, *params)
    n = x2.shape[0]
    k = min(threads, n)
    if k <= 1:
        return fn(x, *params)
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=threads)
    bounds = [(i * n // k, (i + 1) * n // k) for i in range(k)]
    parts = list(_POOL.map(lambda ab: fn(x2[ab[0]:ab[1]], *[slicer(p, ab[0], ab[1]) for p in ps]), bounds))
    return np.concatenate(parts, axis=0).reshape(shape)


_POOL = None


ORACLE_THREADS = int(__import__("os").environ.get("SPQ_ORACLE_THREADS", "0")) or (__import__("os").cpu_count() or 1)


def fake_quantize(q: QuantizerState, x: np.ndarray, **kw) -> np.ndarray:
    """LearnableFakeQuantize.forward (p1/quantization.py:211-226)."""
    if q.num_bits >= 32:
        return x
    if q.collecting_stats:
        collect_statistics(q, x, **({"log2_fn": kw["log2_fn"]} if "log2_fn" in kw else {}))
        return x
    if not q.calibrated:
        raise RuntimeError(
            f"Quantizer not calibrated. Please run calibration first for {q.quantizer_type} quantizer.")
    big = np.size(x) >= (1 << 20) and ORACLE_THREADS > 1 and not kw
    if q.quantizer_type == "minmax":
        if big:
            return _chunked(lambda a, s, z: minmax_quantize(a, s, z, q.num_bits, q.symmetric)[0],
                            np.asarray(x, F32), [np.asarray(q.scale, F32), np.asarray(q.zero_point, F32)], ORACLE_THREADS)
        return minmax_quantize(x, q.scale, q.zero_point, q.num_bits, q.symmetric)[0]
    if q.quantizer_type == "log":
        if big:
            return _chunked(lambda a, lmin, lrng: log_quantize(a, lmin, lrng, q.num_bits, q.symmetric)[0],
                            np.asarray(x, F32), [np.asarray(q.zero_point, F32), np.asarray(q.scale, F32)], ORACLE_THREADS)
        return log_quantize(x, q.zero_point, q.scale, q.num_bits, q.symmetric, **kw)[0]
    raise ValueError(f"Unknown quantizer type: {q.quantizer_type}. Supported types: 'minmax', 'log'")


def ste_backward(grad_out: np.ndarray, quantizer_type: str) -> np.ndarray:
    """Straight-through estimator (p1/quantization_methods.py:25-28, 82-90):
    identity for min-max, clamp to [-10, 10] for log; no gradient to scale/zp."""
    g = np.array(grad_out, dtype=F32, copy=True)
    if quantizer_type == "log":
        g = np.clip(g, F32(-10), F32(10))
    return g
