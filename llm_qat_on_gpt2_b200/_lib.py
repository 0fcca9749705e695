"""ctypes binding of libspq_b200.so (the C ABI declared in include/spq_b200.h).

PyTorch is used for device memory and streams only: every wrapper takes torch CUDA tensors,
passes raw device pointers plus the current CUDA stream, and raises on any non-zero status.
There is no CPU path: calling a compute wrapper with the library missing, or with a tensor that
is not on a CUDA device, raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspq_b200.so")

PER_TENSOR, PER_ROW, PER_COL = 0, 1, 2
MINMAX, LOG = 0, 1
OPERAND_CODE, OPERAND_DEQUANT, OPERAND_RAW, OPERAND_CODE_E4M3 = 0, 1, 2, 3
QTYPE = {"minmax": MINMAX, "log": LOG}
ABI_VERSION = 6

# name -> (restype, argtypes); must list every SPQ_API symbol of include/spq_b200.h
SIGNATURES = {
    "spq_abi_version": (c_int, []),
    "spq_last_error": (c_char_p, []),
    "spq_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "spq_debug_status": (c_int, [POINTER(c_int)]),
    "spq_launch_count": (c_int64, []),
    "spq_stats_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "spq_minmax_stats": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "spq_calibrate_many": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "spq_finish_calibration": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p,
                                       c_void_p, c_void_p]),
    "spq_fake_quantize": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float,
                                  c_int, c_int64, c_void_p]),
    "spq_quantize_act": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "spq_prep_linear_scales": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_int64,
                                       c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "spq_cpt_lora_scales": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "spq_ste_backward": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "spq_qgemm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                          c_void_p, c_int64, c_void_p, c_int64, c_int64,
                          c_float, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64,
                          c_void_p, c_int64, c_int, c_int, c_void_p]),
    "spq_qgemm_f8": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                             c_void_p, c_int64, c_void_p, c_int64, c_int64,
                             c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "spq_gemm_tn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "spq_gemm_tn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p,
                            c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p, c_size_t,
                            c_void_p]),
    "spq_rowscale_f16_max": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "spq_rowscale_dgelu_f16_max": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "spq_lora_bwd_prep": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "spq_softmax_loss_grad16": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_float,
                                        c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "spq_layernorm_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "spq_ln_quantize_act": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "spq_ln_rowscale_stats_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "spq_ln_rowscale_stats": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_float,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "spq_layernorm_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "spq_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "spq_layernorm_bwd_finalize": (c_int, [c_void_p, c_size_t, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "spq_qgemm_lse_parts": (c_int64, [c_int64, c_int64]),
    "spq_qgemm_lse": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "spq_cross_entropy_from_parts": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                             c_void_p, c_void_p, c_void_p]),
    "spq_mse_select_workspace_bytes": (c_size_t, []),
    "spq_mse_select": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "spq_distill_kl": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_float, c_int64, c_float, c_void_p,
                               c_void_p, c_void_p]),
    "spq_cross_entropy_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "spq_rowscale_f16": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "spq_sumsq_workspace_bytes": (c_size_t, []),
    "spq_grad_sumsq": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "spq_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_float,
                               c_int64, c_float, c_void_p, c_float, c_void_p]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and type every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if path == LIB_PATH and os.environ.get("SPQ_LIB"):
        path = os.environ["SPQ_LIB"]             # same-box A/B of two builds (tools/ab.sh)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m llm_qat_on_gpt2_b200.build` "
            "(there is no CPU fallback for the fake-quant linear path)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.spq_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libspq_b200 ABI version {lib.spq_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().spq_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    # the raw handle of torch's current stream; the private accessor skips building a Stream object
    # (this runs once per launch and the training step is host-bound)
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _req_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libspq_b200 operates on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


_ws_cache = {}


def _workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    """Per (device, stream, tag) scratch buffer, grown on demand (stream-ordered reuse)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def launch_count() -> int:
    return int(load_library().spq_launch_count())


def debug_status() -> int:
    v = c_int(0)
    _check(load_library().spq_debug_status(ctypes.byref(v)), "spq_debug_status")
    return v.value


def device_info():
    a, b, c = c_int(0), c_int(0), c_int(0)
    _check(load_library().spq_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "spq_device_info")
    return a.value, b.value, c.value


# ------------------------------------------------------------------------------------------
# thin typed wrappers (shapes are validated here; the kernels trust them)
# ------------------------------------------------------------------------------------------

def minmax_stats(x2d: torch.Tensor, bcast: int, log_mode: bool, eps: float, stat_min: torch.Tensor,
                 stat_max: torch.Tensor, accumulate: bool, state: Optional[torch.Tensor]):
    lib = load_library()
    _req_cuda(x2d, stat_min, stat_max, state)
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype in (torch.float32, torch.float16)
    rows, cols = x2d.shape
    n = {PER_TENSOR: 1, PER_ROW: rows, PER_COL: cols}[bcast]
    assert stat_min.numel() == n and stat_max.numel() == n and stat_min.is_contiguous() and stat_max.is_contiguous()
    nbytes = lib.spq_stats_workspace_bytes(rows, cols, bcast)
    ws = _workspace(nbytes, x2d.device, "stats")
    _check(lib.spq_minmax_stats(x2d.data_ptr(), int(x2d.dtype == torch.float16), rows, cols, bcast, int(log_mode), float(eps), stat_min.data_ptr(),
                                stat_max.data_ptr(), int(accumulate), _ptr(state), ws.data_ptr(), ws.numel(),
                                _stream()), "spq_minmax_stats")


CALIB_JOB_FORMAT = "QqqQQQQiiiffi"        # struct SpqCalibJob (include/spq_b200.h), 80 bytes


def calibrate_many(jobs_dev: torch.Tensor, n_jobs: int, max_blocks: int, flags_dev: torch.Tensor):
    """One launch calibrating n_jobs small tensors; jobs_dev is the packed SpqCalibJob table (uint8, on device)."""
    _req_cuda(jobs_dev, flags_dev)
    assert jobs_dev.dtype == torch.uint8 and jobs_dev.numel() >= 80 * n_jobs and flags_dev.dtype == torch.int32
    _check(load_library().spq_calibrate_many(jobs_dev.data_ptr(), n_jobs, max_blocks, flags_dev.data_ptr(), _stream()),
           "spq_calibrate_many")


def finish_calibration(rmin, rmax, qtype: int, symmetric: bool, bits: int, eps: float, scale, zp):
    _req_cuda(rmin, rmax, scale, zp)
    n = rmin.numel()
    assert rmax.numel() == n and scale.numel() == n and zp.numel() == n
    _check(load_library().spq_finish_calibration(rmin.data_ptr(), rmax.data_ptr(), n, qtype, int(symmetric), bits,
                                                 float(eps), scale.data_ptr(), zp.data_ptr(), _stream()),
           "spq_finish_calibration")


def fake_quantize(x2d, scale, zp, bcast, qtype, bits, symmetric, dequant=None, codes=None, sign=None,
                  operand=None, operand_kind=OPERAND_DEQUANT, row_mul=None, col_mul=None, mul=1.0,
                  operand_transposed=False):
    """`operand` may be a column-sliced view of a wider buffer (padded leading dimension)."""
    _req_cuda(x2d, scale, zp, dequant, codes, sign, operand, row_mul, col_mul)
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype == torch.float32
    rows, cols = x2d.shape
    need = {PER_TENSOR: 1, PER_ROW: rows, PER_COL: cols}[bcast]
    assert scale.numel() == need and zp.numel() == need, (scale.shape, zp.shape, bcast, x2d.shape)
    assert row_mul is None or row_mul.numel() == rows
    assert col_mul is None or col_mul.numel() == cols
    assert operand is None or (operand.dtype == torch.uint8) == (operand_kind == OPERAND_CODE_E4M3)
    _check(load_library().spq_fake_quantize(x2d.data_ptr(), rows, cols, scale.data_ptr(), zp.data_ptr(), bcast, qtype,
                                            bits, int(symmetric), _ptr(dequant), _ptr(codes), _ptr(sign),
                                            _ptr(operand), operand_kind, _ptr(row_mul), _ptr(col_mul), float(mul),
                                            int(operand_transposed), 0 if operand is None else operand.stride(0),
                                            _stream()), "spq_fake_quantize")


def quantize_act(x2d, scale, zp, bcast, qtype, bits, symmetric, operand_kind, col_mul, mul, a_q, a_raw, raw_col_mul):
    _req_cuda(x2d, scale, zp, col_mul, a_q, a_raw, raw_col_mul)
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype in (torch.float32, torch.float16)
    M, K = x2d.shape
    _check(load_library().spq_quantize_act(x2d.data_ptr(), int(x2d.dtype == torch.float16), M, K, _ptr(scale), _ptr(zp), bcast, qtype, bits,
                                           int(symmetric), operand_kind, _ptr(col_mul), float(mul), _ptr(a_q),
                                           _ptr(a_raw), _ptr(raw_col_mul), _stream()), "spq_quantize_act")


def prep_linear_scales(in_scale, in_zp, qtype, bits, symmetric, K, w_rowmax, N, aq_abs, bq, r, lora_scaling,
                       absorb, act_mul, raw_mul, inv_raw_mul, pw, inv_pw, lora_vec):
    _req_cuda(in_scale, in_zp, w_rowmax, aq_abs, bq, absorb, act_mul, raw_mul, inv_raw_mul, pw, inv_pw, lora_vec)
    _check(load_library().spq_prep_linear_scales(in_scale.data_ptr(), in_zp.data_ptr(), in_scale.numel(), qtype, bits,
                                                 int(symmetric), K, w_rowmax.data_ptr(), N, _ptr(aq_abs), _ptr(bq), r,
                                                 float(lora_scaling), absorb.data_ptr(), act_mul.data_ptr(),
                                                 raw_mul.data_ptr(), inv_raw_mul.data_ptr(), pw.data_ptr(),
                                                 inv_pw.data_ptr(), _ptr(lora_vec), _stream()), "spq_prep_linear_scales")


def cpt_lora_scales(aq, bq, absorb, xbound, scaling, out):
    """Scale vectors of a CPTLinear's shared-LoRA level (see include/spq_b200.h): aq [K, r], bq [N, r] -> out [8 r]."""
    _req_cuda(aq, bq, absorb, xbound, out)
    K, r = aq.shape
    N = bq.shape[0]
    assert bq.shape[1] == r and absorb.numel() == K and xbound.numel() == K and out.numel() == 8 * r
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (aq, bq, absorb, xbound, out))
    _check(load_library().spq_cpt_lora_scales(aq.data_ptr(), bq.data_ptr(), absorb.data_ptr(), xbound.data_ptr(), K, N, r,
                                              float(scaling), out.data_ptr(), _stream()), "spq_cpt_lora_scales")


def ste_backward(grad, qtype: int):
    _req_cuda(grad)
    g = grad.contiguous()
    if g.dtype != torch.float32:
        g = g.float()
    out = torch.empty_like(g)
    if g.numel():
        _check(load_library().spq_ste_backward(g.data_ptr(), g.numel(), qtype, out.data_ptr(), _stream()),
               "spq_ste_backward")
    return out


def qgemm(A, B, M, N, K, out, A2=None, B2=None, K2=0, alpha=1.0, row_scale=None, col_scale=None, bias=None,
          clamp_abs=0.0, C=None, activation=0):
    """out[M,N] = epi(A[M,K] B[N,K]^T + A2[M,K2] B2[N,K2]^T); A/B fp16 row-major, out fp32 or fp16."""
    _req_cuda(A, B, out, A2, B2, row_scale, col_scale, bias, C)
    assert A.dtype == torch.float16 and B.dtype == torch.float16
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and out.stride(-1) == 1
    d_half = out.dtype == torch.float16
    assert d_half or out.dtype == torch.float32
    _check(load_library().spq_qgemm(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K,
                                    _ptr(A2), 0 if A2 is None else A2.stride(0), _ptr(B2),
                                    0 if B2 is None else B2.stride(0), K2, float(alpha), _ptr(row_scale),
                                    _ptr(col_scale), _ptr(bias), float(clamp_abs), _ptr(C),
                                    0 if C is None else C.stride(0), out.data_ptr(), out.stride(0), int(d_half),
                                    int(activation), _stream()), "spq_qgemm")
    return out


def qgemm_f8(A8, B8, M, N, K, out, A2=None, B2=None, K2=0, alpha=1.0, row_scale=None, col_scale=None, bias=None, activation=0):
    """out[M,N] = epi(A8[M,K] B8[N,K]^T + A2[M,K2] B2[N,K2]^T): A8 / B8 e4m3 integer codes (uint8 storage, row strides
    multiples of 16 bytes) on tcgen05.mma.kind::f8f6f4, the optional second segment fp16; out fp32 or fp16."""
    _req_cuda(A8, B8, out, A2, B2, row_scale, col_scale, bias)
    assert A8.dtype == torch.uint8 and B8.dtype == torch.uint8 and A8.stride(-1) == 1 and B8.stride(-1) == 1
    assert out.stride(-1) == 1 and out.dtype in (torch.float32, torch.float16)
    _check(load_library().spq_qgemm_f8(A8.data_ptr(), A8.stride(0), B8.data_ptr(), B8.stride(0), M, N, K,
                                       _ptr(A2), 0 if A2 is None else A2.stride(0), _ptr(B2), 0 if B2 is None else B2.stride(0),
                                       K2, float(alpha), _ptr(row_scale), _ptr(col_scale), _ptr(bias), out.data_ptr(),
                                       out.stride(0), int(out.dtype == torch.float16), int(activation), _stream()), "spq_qgemm_f8")
    return out


def gemm_tn(P, Q, out, alpha=1.0, alpha_dev=None, i_scale=None, j_scale=None, transposed_out=False, clamp_abs=0.0,
            gq_scale_i=None, gq_bits=8, accumulate=False):
    """out[I,J] (or out[J,I] when transposed_out) = clamp(alpha * P[Mred,I]^T Q[Mred,J]); fp16 in, fp32 out.
    Deterministic (split reduction folded in a fixed order); clamp_abs > 0 applies the log STE clamp."""
    lib = load_library()
    _req_cuda(P, Q, out, alpha_dev, i_scale, j_scale, gq_scale_i)
    assert P.dtype == torch.float16 and Q.dtype == torch.float16 and out.dtype == torch.float32
    assert P.stride(-1) == 1 and Q.stride(-1) == 1 and out.is_contiguous()
    Mred, I = P.shape
    assert gq_scale_i is None or (gq_scale_i.numel() == I and gq_scale_i.dtype == torch.float32 and gq_scale_i.is_contiguous())
    J = Q.shape[1]
    assert Q.shape[0] == Mred
    if transposed_out:
        assert tuple(out.shape) == (J, I)
        si, sj = 1, I
    else:
        assert tuple(out.shape) == (I, J)
        si, sj = J, 1
    ws = _workspace(lib.spq_gemm_tn_workspace_bytes(Mred, I, J), out.device, "gemm_tn")
    _check(lib.spq_gemm_tn(P.data_ptr(), P.stride(0), Q.data_ptr(), Q.stride(0), Mred, I, J, float(alpha),
                           _ptr(alpha_dev), _ptr(i_scale), _ptr(j_scale), float(clamp_abs), _ptr(gq_scale_i), int(gq_bits),
                           int(accumulate), out.data_ptr(), si, sj,
                           ws.data_ptr(), ws.numel(), _stream()), "spq_gemm_tn")
    return out


def layernorm_fwd(x2d, weight, bias, eps, y, mean, rstd):
    _req_cuda(x2d, weight, bias, y, mean, rstd)
    rows, cols = x2d.shape
    _check(load_library().spq_layernorm_fwd(x2d.data_ptr(), rows, cols, weight.data_ptr(), bias.data_ptr(), float(eps),
                                            y.data_ptr(), _ptr(mean), _ptr(rstd), _stream()), "spq_layernorm_fwd")


def ln_quantize_act(x2d, ln_w, ln_b, ln_eps, scale, zp, bcast, qtype, bits, symmetric, operand_kind, col_mul, mul, a_q, a_raw,
                    raw_col_mul, y_out=None):
    """quantize_act(layernorm(x2d)) without the float32 round trip of the normalised rows."""
    _req_cuda(x2d, ln_w, ln_b, scale, zp, col_mul, a_q, a_raw, raw_col_mul, y_out)
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype == torch.float32
    M, K = x2d.shape
    assert ln_w.numel() == K and ln_b.numel() == K and ln_w.dtype == torch.float32 and ln_b.dtype == torch.float32
    assert a_q.shape == (M, K) and a_q.is_contiguous() and (a_raw is None or (a_raw.shape == (M, K) and a_raw.is_contiguous()))
    _check(load_library().spq_ln_quantize_act(x2d.data_ptr(), M, K, ln_w.data_ptr(), ln_b.data_ptr(), float(ln_eps), scale.data_ptr(),
                                              zp.data_ptr(), bcast, qtype, bits, int(symmetric), operand_kind, _ptr(col_mul),
                                              float(mul), a_q.data_ptr(), _ptr(a_raw), _ptr(raw_col_mul), _ptr(y_out), _stream()),
           "spq_ln_quantize_act")


def ln_rowscale_stats(x2d, ln_w, ln_b, ln_eps, out, row_scale, stats_mode=0, stat_eps=0.0, stat_min=None, stat_max=None,
                      accumulate=False, state=None, y_out=None):
    """rowscale_f16(layernorm(x2d)) and, with stats_mode 1 (min-max) / 2 (log), minmax_stats(layernorm(x2d), PER_COL) too."""
    lib = load_library()
    _req_cuda(x2d, ln_w, ln_b, out, row_scale, stat_min, stat_max, state, y_out)
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype == torch.float32
    M, K = x2d.shape
    assert out.dtype == torch.float16 and out.shape == (M, K) and out.stride(0) == K and row_scale.numel() == M
    ws = None
    nbytes = 0
    if stats_mode:
        assert stat_min.numel() == K and stat_max.numel() == K and stat_min.is_contiguous() and stat_max.is_contiguous()
        nbytes = lib.spq_ln_rowscale_stats_workspace_bytes(M, K)
        ws = _workspace(nbytes, x2d.device, "ln_stats")
    _check(lib.spq_ln_rowscale_stats(x2d.data_ptr(), M, K, ln_w.data_ptr(), ln_b.data_ptr(), float(ln_eps), out.data_ptr(),
                                     row_scale.data_ptr(), int(stats_mode), float(stat_eps), _ptr(stat_min), _ptr(stat_max),
                                     int(accumulate), _ptr(state), _ptr(y_out), _ptr(ws), ws.numel() if ws is not None else 0,
                                     _stream()), "spq_ln_rowscale_stats")


def layernorm_bwd(dy2d, x2d, weight, mean, rstd, dx, dweight, dbias, accumulate_params=False):
    lib = load_library()
    _req_cuda(dy2d, x2d, weight, mean, rstd, dx, dweight, dbias)
    rows, cols = x2d.shape
    nbytes = lib.spq_layernorm_bwd_workspace_bytes(rows, cols)
    ws = _workspace(nbytes, x2d.device, "ln_bwd")
    _check(lib.spq_layernorm_bwd(dy2d.data_ptr(), x2d.data_ptr(), weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                 rows, cols, dx.data_ptr(), _ptr(dweight), _ptr(dbias), int(accumulate_params), ws.data_ptr(), ws.numel(),
                                 _stream()), "spq_layernorm_bwd")


def layernorm_bwd_split(dy2d, x2d, weight, mean, rstd, dx):
    """layernorm_bwd without the parameter-gradient fold: returns the private workspace holding the per-CTA column sums
    (hand it to layernorm_bwd_finalize, possibly on another stream)."""
    lib = load_library()
    _req_cuda(dy2d, x2d, weight, mean, rstd, dx)
    rows, cols = x2d.shape
    ws = torch.empty(max(int(lib.spq_layernorm_bwd_workspace_bytes(rows, cols)), 16), dtype=torch.uint8, device=x2d.device)
    _check(lib.spq_layernorm_bwd(dy2d.data_ptr(), x2d.data_ptr(), weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                 rows, cols, dx.data_ptr(), None, None, 0, ws.data_ptr(), ws.numel(), _stream()),
           "spq_layernorm_bwd")
    return ws


def layernorm_bwd_finalize(ws, rows, cols, dweight, dbias, accumulate_params=False):
    _req_cuda(ws, dweight, dbias)
    _check(load_library().spq_layernorm_bwd_finalize(ws.data_ptr(), ws.numel(), rows, cols, _ptr(dweight), _ptr(dbias),
                                                     int(accumulate_params), _stream()), "spq_layernorm_bwd_finalize")


def cross_entropy_fwd(logits2d, targets, ignore_index=-100):
    """Mean cross-entropy over rows whose target != ignore_index; logits2d may have a padded row stride."""
    _req_cuda(logits2d, targets)
    assert logits2d.dim() == 2 and logits2d.stride(1) == 1 and logits2d.dtype == torch.float32
    M, V = logits2d.shape
    tg = targets.reshape(-1).to(torch.int64).contiguous()
    assert tg.numel() == M
    out = torch.empty((2, M), dtype=torch.float32, device=logits2d.device)
    _check(load_library().spq_cross_entropy_fwd(logits2d.data_ptr(), M, V, logits2d.stride(0), tg.data_ptr(),
                                                int(ignore_index), out[0].data_ptr(), out[1].data_ptr(), _stream()),
           "spq_cross_entropy_fwd")
    sums = out.sum(dim=1)
    return sums[0] / sums[1]


def qgemm_lse(A, B, M, N, K, out, alpha=1.0, row_scale=None, col_scale=None, bias=None):
    """spq_qgemm with float32 `out` plus the per-row (max, sum exp) pairs of every column half-tile: returns the
    [M, P, 2] partials for cross_entropy_from_parts.  `out` needs 16-byte aligned rows padded to a multiple of 4."""
    lib = load_library()
    _req_cuda(A, B, out, row_scale, col_scale, bias)
    assert A.dtype == torch.float16 and B.dtype == torch.float16 and out.dtype == torch.float32
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and out.stride(-1) == 1
    P = int(lib.spq_qgemm_lse_parts(M, N))
    parts = torch.empty((M, P, 2), dtype=torch.float32, device=out.device)
    _check(lib.spq_qgemm_lse(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, float(alpha), _ptr(row_scale),
                             _ptr(col_scale), _ptr(bias), out.data_ptr(), out.stride(0), parts.data_ptr(), P, _stream()),
           "spq_qgemm_lse")
    return parts


def cross_entropy_from_parts(parts, logits2d, targets, ignore_index=-100):
    """Mean cross-entropy from qgemm_lse's partials; reads one logit per row."""
    _req_cuda(parts, logits2d, targets)
    M, V = logits2d.shape
    assert parts.dim() == 3 and parts.shape[0] == M and parts.shape[2] == 2 and parts.is_contiguous()
    assert logits2d.stride(1) == 1 and logits2d.dtype == torch.float32
    tg = targets.reshape(-1).to(torch.int64).contiguous()
    assert tg.numel() == M
    out = torch.empty((2, M), dtype=torch.float32, device=logits2d.device)
    _check(load_library().spq_cross_entropy_from_parts(parts.data_ptr(), parts.shape[1], parts.shape[1], logits2d.data_ptr(),
                                                       M, V, logits2d.stride(0), tg.data_ptr(), int(ignore_index),
                                                       out[0].data_ptr(), out[1].data_ptr(), _stream()),
           "spq_cross_entropy_from_parts")
    sums = out.sum(dim=1)
    return sums[0] / sums[1]


def mse_select(a_list, b_list, select: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[0] = F.mse_loss(a_list[l], b_list[l]) with l = select[0] read on the device (int32 tensor of one element)."""
    lib = load_library()
    n = len(a_list)
    assert n == len(b_list) and 0 < n <= 32 and select.dtype == torch.int32 and select.numel() == 1
    _req_cuda(select, out, *a_list, *b_list)
    numel = a_list[0].numel()
    for t in list(a_list) + list(b_list):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == numel
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=select.device)
    ws = _workspace(lib.spq_mse_select_workspace_bytes(), select.device, "mse")
    pa = (c_void_p * n)(*[t.data_ptr() for t in a_list])
    pb = (c_void_p * n)(*[t.data_ptr() for t in b_list])
    _check(lib.spq_mse_select(pa, pb, n, select.data_ptr(), numel, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
           "spq_mse_select")
    return out


def distill_kl(s2d, t2d, temperature, seq_len, grad_scale, want_grad=True):
    """Per-row KL(softmax(t/T) || softmax(s/T)) and (optionally) the dense gradient w.r.t. s; rows may be padded."""
    _req_cuda(s2d, t2d)
    assert s2d.dim() == 2 and t2d.shape == s2d.shape and s2d.stride(1) == 1 and t2d.stride(1) == 1
    assert s2d.dtype == torch.float32 and t2d.dtype == torch.float32
    M, V = s2d.shape
    row_loss = torch.empty(M, dtype=torch.float32, device=s2d.device)
    grad = torch.empty((M, V), dtype=torch.float32, device=s2d.device) if want_grad else None
    _check(load_library().spq_distill_kl(s2d.data_ptr(), s2d.stride(0), t2d.data_ptr(), t2d.stride(0), M, V,
                                         float(temperature), int(seq_len), float(grad_scale), row_loss.data_ptr(),
                                         _ptr(grad), _stream()), "spq_distill_kl")
    return row_loss, grad


def rowscale_f16(g2d, out, row_scale):
    _req_cuda(g2d, out, row_scale)
    M, N = g2d.shape
    assert out.stride(-1) == 1 and g2d.is_contiguous() and g2d.dtype in (torch.float32, torch.float16)
    _check(load_library().spq_rowscale_f16(g2d.data_ptr(), int(g2d.dtype == torch.float16), M, N, out.data_ptr(), out.stride(0), row_scale.data_ptr(),
                                           _stream()), "spq_rowscale_f16")


def grad_sumsq(flat_grad: torch.Tensor, out: torch.Tensor, scale: float = 1.0):
    """out[0] = scale^2 * sum(flat_grad^2), deterministic."""
    lib = load_library()
    _req_cuda(flat_grad, out)
    assert flat_grad.dtype == torch.float32 and flat_grad.is_contiguous() and out.dtype == torch.float32
    ws = _workspace(lib.spq_sumsq_workspace_bytes(), flat_grad.device, "sumsq")
    _check(lib.spq_grad_sumsq(flat_grad.data_ptr(), flat_grad.numel(), float(scale), out.data_ptr(), ws.data_ptr(), ws.numel(),
                              _stream()), "spq_grad_sumsq")


def adamw_flat(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step, grad_scale=1.0, total_sumsq=None,
               max_norm=0.0):
    """In-place AdamW on one flat float32 segment (see include/spq_b200.h)."""
    _req_cuda(param, grad, exp_avg, exp_avg_sq, total_sumsq)
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (param, grad, exp_avg, exp_avg_sq))
    _check(load_library().spq_adamw_flat(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                         float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                         int(step), float(grad_scale), _ptr(total_sumsq), float(max_norm), _stream()),
           "spq_adamw_flat")


def rowscale_f16_max(g2d, out, row_scale, max_scale):
    """rowscale_f16 that also leaves max(row_scale) in the device scalar `max_scale`."""
    _req_cuda(g2d, out, row_scale, max_scale)
    M, N = g2d.shape
    assert out.stride(-1) == 1 and g2d.is_contiguous() and g2d.dtype in (torch.float32, torch.float16)
    _check(load_library().spq_rowscale_f16_max(g2d.data_ptr(), int(g2d.dtype == torch.float16), M, N, out.data_ptr(), out.stride(0),
                                               row_scale.data_ptr(), max_scale.data_ptr(), _stream()), "spq_rowscale_f16_max")


def rowscale_dgelu_f16_max(g2d, y2d, out, row_scale, max_scale):
    """rowscale_f16_max of g2d * gelu'(y2d) (exact erf GELU): the gradient through nn.GELU() without torch's gelu_backward pass."""
    _req_cuda(g2d, y2d, out, row_scale, max_scale)
    M, N = g2d.shape
    assert g2d.dtype == torch.float32 and y2d.dtype == torch.float32 and g2d.is_contiguous() and y2d.is_contiguous()
    assert y2d.shape == g2d.shape and out.shape == (M, N) and out.stride(0) == N and out.dtype == torch.float16
    _check(load_library().spq_rowscale_dgelu_f16_max(g2d.data_ptr(), y2d.data_ptr(), M, N, out.data_ptr(), row_scale.data_ptr(),
                                                     max_scale.data_ptr(), _stream()), "spq_rowscale_dgelu_f16_max")


def lora_bwd_prep(dtn, t16, row_scale, max_scale, dt_mul, want_dt16=True, want_dt2=True, want_t2=True):
    """(dt16, dt2, t2) fp16 operands of the LoRA gradient GEMMs in one launch (see include/spq_b200.h)."""
    _req_cuda(dtn, t16, row_scale, max_scale)
    src = dtn if dtn is not None else t16
    M, r = src.shape
    dev = src.device
    if r % 4:
        # odd ranks (no reference configuration has one): the same arithmetic with torch device ops
        e = (row_scale / max_scale).unsqueeze(1)
        d = None if dtn is None else dtn * dt_mul
        return (d.half() if (d is not None and want_dt16) else None, (d * e).half() if (d is not None and want_dt2) else None,
                (t16.float() * e).half() if (t16 is not None and want_t2) else None)
    dt16 = torch.empty((M, r), dtype=torch.float16, device=dev) if (dtn is not None and want_dt16) else None
    dt2 = torch.empty((M, r), dtype=torch.float16, device=dev) if (dtn is not None and want_dt2) else None
    t2 = torch.empty((M, r), dtype=torch.float16, device=dev) if (t16 is not None and want_t2) else None
    assert dtn is None or (dtn.dtype == torch.float32 and dtn.is_contiguous())
    assert t16 is None or (t16.dtype == torch.float16 and t16.stride(-1) == 1)
    _check(load_library().spq_lora_bwd_prep(_ptr(dtn), _ptr(t16), 0 if t16 is None else t16.stride(0), row_scale.data_ptr(),
                                            max_scale.data_ptr(), M, r, float(dt_mul), _ptr(dt16), _ptr(dt2), _ptr(t2),
                                            _stream()), "spq_lora_bwd_prep")
    return dt16, dt2, t2


def softmax_loss_grad16(s2d, kind, *, t2d=None, targets=None, temperature=1.0, seq_len=0, ignore_index=-100):
    """kind 'kl' | 'ce' -> (row_loss [M], row_valid [M], g16 [M, V] fp16 view (padded rows), row_scale [M], max_scale [1])."""
    _req_cuda(s2d, t2d, targets)
    assert s2d.dim() == 2 and s2d.stride(1) == 1 and s2d.dtype == torch.float32
    M, V = s2d.shape
    dev = s2d.device
    ld = (V + 7) // 8 * 8
    gbuf = torch.empty((M, ld), dtype=torch.float16, device=dev)
    row_loss = torch.empty(M, dtype=torch.float32, device=dev)
    row_valid = torch.empty(M, dtype=torch.float32, device=dev)
    row_scale = torch.empty(M, dtype=torch.float32, device=dev)
    max_scale = torch.empty(1, dtype=torch.float32, device=dev)
    if kind == 'kl':
        assert t2d is not None and t2d.shape == s2d.shape and t2d.stride(1) == 1 and t2d.dtype == torch.float32
        tg = None
    else:
        tg = targets.reshape(-1).to(torch.int64).contiguous()
        assert tg.numel() == M
    _check(load_library().spq_softmax_loss_grad16(0 if kind == 'kl' else 1, s2d.data_ptr(), s2d.stride(0), _ptr(t2d),
                                                  0 if t2d is None else t2d.stride(0), _ptr(tg), int(ignore_index), M, V,
                                                  float(temperature), int(seq_len), row_loss.data_ptr(), row_valid.data_ptr(),
                                                  gbuf.data_ptr(), ld, row_scale.data_ptr(), max_scale.data_ptr(), _stream()),
           "spq_softmax_loss_grad16")
    return row_loss, row_valid, (gbuf[:, :V] if ld != V else gbuf), row_scale, max_scale


def empty_f16_padded(rows: int, cols: int, device) -> torch.Tensor:
    """[rows, cols] fp16 view whose row stride is a multiple of 8 elements (TMA needs 16-byte strides)."""
    ld = (cols + 7) // 8 * 8
    buf = torch.empty((rows, ld), dtype=torch.float16, device=device)
    return buf if ld == cols else buf[:, :cols]
