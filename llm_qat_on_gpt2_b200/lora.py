"""LoRALayer and SPLinearWithLoRA on the B200 kernels -- drop-ins for the reference classes of the
same names (part1_switchable_precision/lora.py:13-150; byte-identical copy in part5_squad).

Same constructors, attributes, ModuleDict keys ('{bits}bit'), state_dict keys and exceptions.
What changes is where the arithmetic runs:

  reference forward (p1/lora.py:127-150)          this module
  -------------------------------------------     ------------------------------------------------
  q_in(x)            ~4 / ~25 eager kernels   ->   spq_quantize_act: one pass over x writes the fp16
  lora reads x again                               code/dequant operand AND the row-scaled raw operand
  q_w(W) every call                           ->   cached fp16 operand, rebuilt only when W or a
  q_A(A), q_B(B) every call                        calibration changes (input scale folded in per K)
  F.linear + 2 matmuls + mul + add            ->   spq_qgemm (tcgen05): base and LoRA up-projection in
                                                   one TMEM accumulator, scale + bias in the epilogue
  autograd of the above (STE)                 ->   spq_qgemm / spq_gemm_tn on fp16 gradient operands

Numerics: operands are fp16 (integer codes are exact; dequantised values carry 2^-11 relative
rounding), accumulation is fp32; every power-of-two pre-scale is exact.  Outputs agree with the
fp32 reference to rel ~3e-4 (tolerance 1e-3, BASELINE.json).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib
from .quantization import LearnableFakeQuantize, pow2_ceil
from .quantization_methods import _view2d


# ----------------------------------------------------------------------------------------------
# small helpers (device-side plumbing on tiny tensors; never a host sync)
# ----------------------------------------------------------------------------------------------

def _dummy_params(device):
    key = (device.type, device.index)
    p = _dummy_params.cache.get(key)
    if p is None:
        p = (torch.ones(1, device=device), torch.zeros(1, device=device))
        _dummy_params.cache[key] = p
    return p


_dummy_params.cache = {}


def _to_f16_operand(x2d, row_mul=None, col_mul=None, mul=1.0, transposed=False):
    """fp16(x * row_mul[:,None] * col_mul[None,:] * mul) through the quantise kernel's raw mode."""
    one, zero = _dummy_params(x2d.device)
    rows, cols = x2d.shape
    out = _lib.empty_f16_padded(cols, rows, x2d.device) if transposed else _lib.empty_f16_padded(rows, cols, x2d.device)
    _lib.fake_quantize(x2d, one, zero, _lib.PER_TENSOR, _lib.MINMAX, 8, True, operand=out,
                       operand_kind=_lib.OPERAND_RAW, row_mul=row_mul, col_mul=col_mul, mul=mul,
                       operand_transposed=transposed)
    return out


def _quantizer_params(q: LearnableFakeQuantize, x_like: torch.Tensor):
    """(scale, zero_point) flattened + broadcast mode of quantiser q over a tensor shaped like x_like."""
    sc = q.scale.detach().float().contiguous()
    _, bcast = _view2d(x_like, sc)
    zp = q.zero_point.detach().float()
    if zp.numel() != sc.numel():
        zp = zp.expand_as(sc)
    return sc.reshape(-1).contiguous(), zp.reshape(-1).contiguous(), bcast


def _dequant(q: LearnableFakeQuantize, w: torch.Tensor) -> torch.Tensor:
    sc, zp, bcast = _quantizer_params(q, w)
    w2 = w.detach().float().contiguous()
    x2d = w2.reshape(-1, w2.shape[-1]) if bcast != _lib.PER_ROW else w2.reshape(w2.shape[0], -1)
    out = torch.empty_like(x2d)
    _lib.fake_quantize(x2d, sc, zp, bcast, _lib.QTYPE[q.quantizer_type], q.num_bits, q.symmetric, dequant=out)
    return out.view(w.shape)


def _quantized_operand(q, w, row_mul=None, col_mul=None, mul=1.0, transposed=False):
    """fp16(q(w) * row_mul * col_mul * mul) for a 2-D parameter w, straight from the quantise kernel."""
    sc, zp, bcast = _quantizer_params(q, w)
    w2 = w.detach().float().contiguous()
    rows, cols = w2.shape
    out = _lib.empty_f16_padded(cols, rows, w.device) if transposed else _lib.empty_f16_padded(rows, cols, w.device)
    _lib.fake_quantize(w2, sc, zp, bcast, _lib.QTYPE[q.quantizer_type], q.num_bits, q.symmetric, operand=out,
                       operand_kind=_lib.OPERAND_DEQUANT, row_mul=row_mul, col_mul=col_mul, mul=mul,
                       operand_transposed=transposed)
    return out


def _norm_pow2(absmax: torch.Tensor, target_log2: int = 8) -> torch.Tensor:
    """Power of two p with absmax / p in (2^(target-1), 2^target]; p = 1 for all-zero rows."""
    return torch.where(absmax > 0, pow2_ceil(absmax) * (2.0 ** -target_log2), torch.ones_like(absmax))


def _rowscaled_f16(x2d: torch.Tensor):
    """Row-scaled fp16 copy of a float32 matrix: returns (x16, row_scale) with x = x16 * row_scale."""
    M, K = x2d.shape
    x16 = _lib.empty_f16_padded(M, K, x2d.device)
    rs = torch.empty(M, dtype=torch.float32, device=x2d.device)
    _lib.rowscale_f16(x2d, x16, rs)
    return x16, rs


# LayerNorm fused into the consumer's activation-side kernel (SPQ_FUSE_LN=0: A/B switch, read at import)
_FUSE_LN = os.environ.get("SPQ_FUSE_LN", "1") != "0"
# gelu'(y) folded into the backward's row-scaling pass when GELU follows a linear under autograd (SPQ_FUSE_DGELU=0: A/B)
_FUSE_DGELU = os.environ.get("SPQ_FUSE_DGELU", "1") != "0"


def _ln_params(ln):
    """(weight, bias, eps) of the active precision of a SwitchableLayerNorm, as flat float32 vectors."""
    key = str(ln.current_precision)
    return (ln.weights[key].detach().reshape(-1).float().contiguous(), ln.biases[key].detach().reshape(-1).float().contiguous(),
            float(ln.eps))


def _ln_fusable(x: torch.Tensor, ln, K: int) -> bool:
    """The fused kernels take a float32, contiguous [..., K] input under no_grad, K a multiple of 4.  One warp holds a
    row in registers: beyond 1024 columns (8 float4 per lane) occupancy drops to one CTA per SM and the separate
    kernels are faster -- GPT-2 XL (1600) keeps the unfused path."""
    return (_FUSE_LN and ln is not None and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32
            and x.is_contiguous() and x.shape[-1] == K and K % 4 == 0 and K <= 1024 and x.numel() > 0
            and ln.__class__.__name__ == 'SwitchableLayerNorm' and getattr(ln, '_ncols', None) == K
            and x.data_ptr() % 16 == 0)


def _ln_rowscaled_f16(x2d: torch.Tensor, ln, stats=None):
    """(x16, row_scale) of layernorm(x2d), one pass; `stats` = (mode, eps, temp_min, temp_max, accumulate, state)."""
    M, K = x2d.shape
    w, b, eps = _ln_params(ln)
    x16 = torch.empty((M, K), dtype=torch.float16, device=x2d.device)
    rs = torch.empty(M, dtype=torch.float32, device=x2d.device)
    if stats is None:
        _lib.ln_rowscale_stats(x2d, w, b, eps, x16, rs)
    else:
        mode, seps, tmin, tmax, acc, state = stats
        _lib.ln_rowscale_stats(x2d, w, b, eps, x16, rs, stats_mode=mode, stat_eps=seps, stat_min=tmin, stat_max=tmax,
                               accumulate=acc, state=state)
    return x16, rs


def _as_2d_f32(x: torch.Tensor, last: int) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("SPLinearWithLoRA runs on the CUDA kernels only (no CPU fallback); "
                           f"got a tensor on {x.device}")
    x2 = x.reshape(-1, last)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    return x2.contiguous()


def _as_2d_act(x: torch.Tensor, last: int, max_cols: int = 1 << 30) -> torch.Tensor:
    """Layer input as a contiguous 2-D matrix.  float16 inputs (the fp16 attention output feeding c_proj)
    stay float16 -- the activation-side kernels widen them exactly -- everything else becomes float32."""
    if x.is_cuda and x.dtype == torch.float16 and last % 4 == 0 and last <= max_cols:
        x2 = x.reshape(-1, last).contiguous()
        if x2.data_ptr() % 16 == 0:
            return x2
    return _as_2d_f32(x, last)


def _grad_sink(param, shape):
    """The tensor a backward kernel may accumulate this parameter's gradient INTO, or None.  A training driver that owns
    the .grad buffers (training.FlatTrainState: slices of one flat float32 buffer, zeroed once per optimizer step) marks
    its parameters `_spq_accumulate_in_place`; the gradient GEMM's fold pass then adds straight into p.grad and the
    autograd node returns None for it -- the AccumulateGrad add (one small launch per parameter per micro-step) is gone."""
    if not getattr(param, '_spq_accumulate_in_place', False):
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or tuple(g.shape) != tuple(shape):
        return None
    return g


class _GradSide:
    """Optional side stream for the LoRA weight-gradient GEMMs (dA = x^T dT, dB = t^T dY): they hang off the backward
    chain (nothing downstream reads them before the optimizer step), so a training driver that owns the step can let
    them run beside the dX GEMMs of the layers that follow -- parallel branches once the micro-step is captured into a
    CUDA graph.  OFF unless a driver enters `side_stream_grads()`; the driver joins with `.join()` before it reads the
    gradients (training.SPTrainer, cpt.CPTTrainer).  Inputs are `record_stream`-ed: the caching allocator must not hand
    their memory to a later main-stream allocation while the side kernel may still read it."""
    stream = None

    @classmethod
    def fork(cls, *tensors):
        """-> the stream to launch on (the current one when the feature is off)."""
        side = cls.stream
        cur = torch.cuda.current_stream()
        if side is None:
            return cur
        side.wait_stream(cur)
        for t in tensors:
            if t is not None:
                t.record_stream(side)
        return side

    @classmethod
    def join(cls):
        if cls.stream is not None:
            torch.cuda.current_stream().wait_stream(cls.stream)


class side_stream_grads:
    """with side_stream_grads(stream): ... backward ... -- see _GradSide."""

    def __init__(self, stream):
        self.stream, self.prev = stream, None

    def __enter__(self):
        self.prev, _GradSide.stream = _GradSide.stream, self.stream
        return self

    def __exit__(self, *exc):
        _GradSide.join()
        _GradSide.stream = self.prev
        return False


def _as_2d_grad(gy: torch.Tensor, last: int) -> torch.Tensor:
    """Incoming gradient as a contiguous 2-D matrix.  float16 gradients (the fp16 attention backward feeding c_attn)
    go to the row-scaling kernel as they are -- it widens them exactly -- instead of through a float32 copy."""
    return _as_2d_act(gy, last, max_cols=8192)


# ----------------------------------------------------------------------------------------------
# plain (unquantised) linear on the tcgen05 GEMM: 32-bit teacher path, calibration pass, LM head
# ----------------------------------------------------------------------------------------------

class _FpWeightCache:
    """fp16 operands of an unquantised weight, keyed on the tensor's storage and version."""

    def __init__(self):
        self.key = None
        self.fwd = None     # (W16 [N,K], pw [N])
        self.bwd = None     # (WT16 [K,N], pk [K])

    def get(self, w: torch.Tensor, transposed: bool):
        key = (w.data_ptr(), w._version, tuple(w.shape))
        if key != self.key:
            self.key, self.fwd, self.bwd = key, None, None
        w2 = w.detach()
        if not transposed:
            if self.fwd is None:
                wf = w2.float().contiguous()
                pw = _norm_pow2(wf.abs().amax(dim=1))
                self.fwd = (_to_f16_operand(wf, row_mul=1.0 / pw), pw)
            return self.fwd
        if self.bwd is None:
            wf = w2.float().contiguous()
            pk = _norm_pow2(wf.abs().amax(dim=0))
            self.bwd = (_to_f16_operand(wf, col_mul=1.0 / pk, transposed=True), pk)
        return self.bwd


class _LinearFpFn(torch.autograd.Function):
    """y = x W^T + b with fp16 operands / fp32 accumulation on spq_qgemm."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache, activation=0, out_half=False, residual=None, lse_out=None, pre=None):
        N, K = weight.shape
        if pre is not None:
            x16, rs = pre                     # (row-scaled fp16 x, row scales) already built by the caller's pass over x
            M = x16.shape[0]
        else:
            x2d = _as_2d_act(x, K, max_cols=8192)
            M = x2d.shape[0]
            x16, rs = _rowscaled_f16(x2d)
        w16, pw = cache.get(weight, transposed=False)
        # odd widths (N = 50257): rows padded to 128 bytes -- the GEMM then stores through TMA, and every 32 x 32
        # block it stores covers whole 32-byte sectors (rows merely 16-byte aligned made L2 fetch the other
        # half of each edge sector: 8.4 GB of DRAM reads for a 6.6 GB logits store).  The caller gets a
        # [..., N] view of the padded buffer
        align = 64 if out_half else 32
        ld = N if N % 4 == 0 else (N + align - 1) // align * align
        ybuf = torch.empty((M, ld), dtype=torch.float16 if out_half else torch.float32, device=x.device)
        dview = ybuf[:, :N] if ld != N else ybuf
        bias_f = None if bias is None else bias.detach().float().contiguous()
        if lse_out is not None:
            # LM head under no_grad: the epilogue also leaves the per-row log-sum-exp partials (fused CE)
            lse_out.append(_lib.qgemm_lse(x16, w16, M, N, K, dview, row_scale=rs, col_scale=pw, bias=bias_f))
        else:
            _lib.qgemm(x16, w16, M, N, K, dview, row_scale=rs, col_scale=pw, bias=bias_f, activation=activation,
                       C=None if residual is None else residual.reshape(M, N))
        ctx.cache = cache
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x16, rs, weight)
        y = ybuf.view(*x.shape[:-1], ld)
        return y[..., :N] if ld != N else y

    @staticmethod
    def backward(ctx, gy):
        x16, rs, weight = ctx.saved_tensors
        N, K = weight.shape
        g2d = _as_2d_grad(gy, N)
        M = g2d.shape[0]
        g16, eg = _rowscaled_f16(g2d)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt16, pk = ctx.cache.get(weight, transposed=True)
            gx = torch.empty((M, K), dtype=torch.float32, device=gy.device)
            _lib.qgemm(g16, wt16, M, K, N, gx, row_scale=eg, col_scale=pk)
            gx = gx.view(ctx.x_shape).to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            # dW[n,k] = sum_m dY[m,n] x[m,k]: the per-token scales sit inside the reduction, so they
            # are folded into one operand: x2 = x16 * (rs*eg / (max rs * max eg))
            gmax, xmax = eg.max(), rs.max()
            fold = (rs / xmax) * (eg / gmax)
            x2 = _to_f16_operand(x16.float(), row_mul=fold)
            gw = torch.empty((N, K), dtype=torch.float32, device=gy.device)
            _lib.gemm_tn(g16, x2, gw, alpha=1.0, alpha_dev=(gmax * xmax).reshape(1).contiguous())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2d.float().sum(dim=0)
        # the residual added in the epilogue passes its gradient through unchanged
        return gx, gw, gb, None, None, None, (gy if ctx.needs_input_grad[6] else None), None, None


def linear_fp(x, weight, bias=None, cache: _FpWeightCache = None, activation: int = 0, out_half: bool = False,
              residual=None, lse_out=None, pre=None):
    """`out_half` stores float16 from the epilogue and `residual` (float32, contiguous, output-shaped) is added in
    the epilogue -- both differentiable (the backward takes float16 gradients; the residual's gradient is the output
    gradient).  `activation=1` fuses the exact-erf GELU into the epilogue and `lse_out` (a list) receives the per-row
    log-sum-exp partials of the output for `_lib.cross_entropy_from_parts`: no-grad fast paths."""
    if (activation or lse_out is not None) and torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        raise RuntimeError("the fused activation / log-sum-exp epilogues are no-grad fast paths")
    return _LinearFpFn.apply(x, weight, bias, cache if cache is not None else _FpWeightCache(), activation, out_half,
                             residual, lse_out, pre)


# ----------------------------------------------------------------------------------------------
# LoRA adapter
# ----------------------------------------------------------------------------------------------

class LoRALayer(nn.Module):
    def __init__(self, in_features, out_features, rank, alpha, bits, quantizer_type, eps=1e-5, per_channel=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.rank = rank
        self.alpha = alpha
        self.bits = bits

        if bits >= 32 or rank <= 0:
            # reference :23-29 -- a disabled adapter keeps [1,1] buffers so the state_dict keys exist
            self.enabled = False
            self.scaling = 0
            self.register_buffer('lora_A', torch.zeros(1, 1))
            self.register_buffer('lora_B', torch.zeros(1, 1))
            self.quantize_A = None
            self.quantize_B = None
        else:
            self.enabled = True
            self.scaling = alpha / rank
            self.lora_A = nn.Parameter(torch.zeros(in_features, rank))
            self.lora_B = nn.Parameter(torch.zeros(rank, out_features))
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)
            self.quantize_A = LearnableFakeQuantize(num_bits=bits, quantizer_type=quantizer_type, channel_dim=1,
                                                    eps=eps, per_channel=per_channel)
            self.quantize_B = LearnableFakeQuantize(num_bits=bits, quantizer_type=quantizer_type, channel_dim=1,
                                                    eps=eps, per_channel=per_channel)
            # unused by the reference's forward too, but part of its state_dict (:42-43)
            self.register_buffer('lora_A_quantized', torch.empty(in_features, rank))
            self.register_buffer('lora_B_quantized', torch.empty(rank, out_features))

    def forward(self, x):
        """Stand-alone adapter output ((x q(A)) q(B)) * alpha/rank (reference :45-54).  Inside
        SPLinearWithLoRA the fused kernels are used instead of this composition."""
        if not self.enabled or self.scaling == 0:
            batch_shape = x.shape[:-1]
            return torch.zeros(*batch_shape, self.out_features, device=x.device, dtype=x.dtype)
        a_q = self.quantize_A(self.lora_A)
        b_q = self.quantize_B(self.lora_B)
        t = linear_fp(x, a_q.t())
        return linear_fp(t, b_q.t()) * self.scaling


# ----------------------------------------------------------------------------------------------
# the fused quantised linear
# ----------------------------------------------------------------------------------------------

def _act_config(qi: LearnableFakeQuantize, K: int):
    """How the activation operand is formed, and the per-K factor the weight operand must absorb.

    min-max: operand = integer code (exact in fp16 up to 11 bits), weights absorb scale[k].
    log    : operand = dequantised value * 2^-e[k] with e from the calibrated log-range,
             weights absorb 2^e[k]."""
    sc = qi.scale.detach().float().reshape(-1).contiguous()
    zp = qi.zero_point.detach().float().reshape(-1)
    if zp.numel() != sc.numel():
        zp = zp.expand_as(sc)
    zp = zp.contiguous()
    if sc.numel() not in (1, K):
        raise NotImplementedError(f"input quantiser scale of {sc.numel()} elements for in_features={K}")
    bcast = _lib.PER_TENSOR if sc.numel() == 1 else _lib.PER_COL
    if qi.quantizer_type == 'minmax':
        code_mul = 2.0 ** -max(0, qi.num_bits - 11)
        kind, col_mul, mul = _lib.OPERAND_CODE, None, code_mul
        absorb = (sc / code_mul).expand(K).contiguous()
    else:
        lmax = (zp + sc.clamp(min=0)).clamp(-100.0, 100.0)
        amul = torch.exp2(8.0 - torch.ceil(lmax)).expand(K).contiguous()
        kind, col_mul, mul = _lib.OPERAND_DEQUANT, amul, 1.0
        absorb = (1.0 / amul).contiguous()
    return dict(scale=sc, zp=zp, bcast=bcast, kind=kind, col_mul=col_mul, mul=mul, absorb=absorb,
                qtype=_lib.QTYPE[qi.quantizer_type], bits=qi.num_bits, symmetric=qi.symmetric,
                input_qtype=qi.quantizer_type)


class _SPLinearFn(torch.autograd.Function):
    """Fused forward / STE backward of SPLinearWithLoRA at a quantised precision."""

    @staticmethod
    def forward(ctx, x, weight, bias, lora_A, lora_B, mod, bits, out_half=False, activation=0, residual=None,
                grad_mode=True, ln=None, gelu_out=False):
        # ln: a SwitchableLayerNorm to apply to x first, inside the activation-side kernel (no_grad only)
        # gelu_out (autograd on, frozen base weight / bias): the Function returns gelu(y) and keeps y; its backward takes
        # the gradient of the GELU OUTPUT and folds gelu'(y) into the row-scaling pass that builds the fp16 gradient
        # operand (spq_rowscale_dgelu_f16_max) -- torch's gelu_backward pass and its float32 result never exist
        # grad_mode: torch.is_grad_enabled() of the caller (always False inside Function.forward)
        use_lora = lora_A is not None
        base, lo = mod._operands_for(bits, use_lora)
        act = base['act']
        N, K = weight.shape
        x2d = _as_2d_act(x, K)
        M = x2d.shape[0]
        need_wgrad = grad_mode and ctx.needs_input_grad[1]
        f8 = base.get('f8') if (residual is None and not need_wgrad) else None
        a_q = torch.empty((M, K), dtype=torch.uint8 if f8 is not None else torch.float16, device=x.device)
        a_raw = torch.empty((M, K), dtype=torch.float16, device=x.device) if use_lora else None
        if ln is not None:
            lw, lb, leps = _ln_params(ln)
            _lib.ln_quantize_act(x2d, lw, lb, leps, act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'],
                                 act['symmetric'], _lib.OPERAND_CODE_E4M3 if f8 is not None else act['kind'], act['col_mul'],
                                 act['mul'], a_q, a_raw, act['raw_mul'] if use_lora else None)
        else:
            _lib.quantize_act(x2d, act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'], act['symmetric'],
                              _lib.OPERAND_CODE_E4M3 if f8 is not None else act['kind'], act['col_mul'], act['mul'], a_q, a_raw,
                              act['raw_mul'] if use_lora else None)
        y = torch.empty((M, N), dtype=torch.float16 if out_half else torch.float32, device=x.device)
        bias_f = None if bias is None else bias.detach().float().contiguous()
        res2d = None if residual is None else residual.reshape(M, N)
        t16 = None
        if use_lora:
            r = lo['rank']
            # the down-projection stores its fp16 operand directly (pa and tau are powers of two, so
            # fp16(acc * (pa tau)) is the value the GEMM consumes); the same tensor is what dB = t^T dY needs later
            t16 = _lib.empty_f16_padded(M, r, x.device)
            _lib.qgemm(a_raw, lo['A_op'], M, r, K, t16, col_scale=lo['pa_tmul'])
            if f8 is not None:
                # integer codes x integer codes on the fp8 tensor pipe; the scales s_x * s_w[n] in the epilogue
                _lib.qgemm_f8(a_q, f8['B8'], M, N, K, y, A2=t16, B2=lo['Bl_op8'], K2=r, col_scale=f8['cs'], bias=bias_f,
                              activation=activation)
            else:
                _lib.qgemm(a_q, base['B_op'], M, N, K, y, A2=t16, B2=lo['Bl_op'], K2=r, col_scale=base['pw'], bias=bias_f,
                           activation=activation, C=res2d)
        elif f8 is not None:
            _lib.qgemm_f8(a_q, f8['B8'], M, N, K, y, col_scale=f8['cs'], bias=bias_f, activation=activation)
        else:
            _lib.qgemm(a_q, base['B_op'], M, N, K, y, col_scale=base['pw'], bias=bias_f, activation=activation, C=res2d)
        ctx.use_lora = use_lora
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.has_bias = bias is not None
        ctx.dims = (M, N, K)
        ctx.base, ctx.lo = base, lo
        need = ctx.needs_input_grad
        ctx.mod, ctx.bits = mod, bits       # backward operands are built (and cached) when backward first runs
        ctx.weight_qtype = mod.quantizers_weight[f'{bits}bit'].quantizer_type
        keep_t = grad_mode and use_lora and need[4]
        ctx.gelu_out = bool(gelu_out)
        ctx.save_for_backward(a_q if need[1] else None, a_raw if (grad_mode and use_lora and need[3]) else None,
                              t16 if keep_t else None, y if gelu_out else None)
        if gelu_out:
            return torch.nn.functional.gelu(y).view(*x.shape[:-1], N)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        a_q, a_raw, t16, y_pre = ctx.saved_tensors
        base, lo = ctx.base, ctx.lo
        bw = ctx.mod._backward_operands_for(ctx.bits, ctx.use_lora)
        M, N, K = ctx.dims
        g2d = _as_2d_grad(gy, N)
        dev = gy.device
        act = base['act']
        need_x, need_w, need_b, need_A, need_B = ctx.needs_input_grad[:5]
        # dY = g16 * eg[:,None]; gmax1 = max eg (all-zero rows -- positions without a loss -- carry the smallest scale)
        g16 = _lib.empty_f16_padded(M, N, dev)
        eg = torch.empty(M, dtype=torch.float32, device=dev)
        gmax1 = torch.empty(1, dtype=torch.float32, device=dev)
        if ctx.gelu_out:
            # gy is the gradient of gelu(y): dY = gy * gelu'(y), formed inside the row-scaling pass
            _lib.rowscale_dgelu_f16_max(g2d if g2d.dtype == torch.float32 else g2d.float(), y_pre, g16, eg, gmax1)
        else:
            _lib.rowscale_f16_max(g2d, g16, eg, gmax1)
        gx = gw = gb = gA = gB = None
        clamp_in = 10.0 if act['input_qtype'] == 'log' else 0.0

        dt16 = None
        if ctx.use_lora and (need_x or need_A or need_B):
            lb = bw['lora']
            r = lo['rank']
            dtn = None
            if need_x or need_A:
                # dtn[m,r] = dt[m,r] / eg[m],  dt = scaling * dY q(B)^T
                dtn = torch.empty((M, r), dtype=torch.float32, device=dev)
                _lib.qgemm(g16, lb['B_rn_op'], M, r, N, dtn, col_scale=lb['pb'])
            # one pass builds every fp16 operand of the three LoRA gradient GEMMs (token scales relative to the
            # largest one folded into the operands of the two token reductions)
            dt16, dt2, t2 = _lib.lora_bwd_prep(dtn, t16 if need_B else None, eg, gmax1, lb['dt_mul'],
                                               want_dt16=need_x, want_dt2=need_A, want_t2=need_B)
            adapter = ctx.mod.lora_adapters[f'{ctx.bits}bit']
            sink_A = _grad_sink(adapter.lora_A, (K, r)) if need_A else None
            sink_B = _grad_sink(adapter.lora_B, (r, N)) if need_B else None
            # gradients that accumulate straight into a driver-owned buffer may run on the driver's side stream
            on_side = (not need_A or sink_A is not None) and (not need_B or sink_B is not None)
            lane = _GradSide.fork(a_raw, dt2, t2, g16, gmax1) if on_side else torch.cuda.current_stream()
            with torch.cuda.stream(lane):
                if need_A:
                    # dA[k,r] = sum_m x[m,k] dt[m,r],  x[m,k] = a_raw[m,k] / raw_mul[k]; log STE clamp in the reduce pass
                    gA = sink_A if sink_A is not None else torch.empty((K, r), dtype=torch.float32, device=dev)
                    _lib.gemm_tn(a_raw, dt2, gA, alpha=1.0 / lb['dt_mul'], alpha_dev=gmax1, i_scale=act['inv_raw_mul'],
                                 clamp_abs=10.0 if lo['qtype_A'] == 'log' else 0.0, accumulate=sink_A is not None)
                    if sink_A is not None:
                        gA = None
                if need_B:
                    # dB[r,n] = scaling * sum_m t[m,r] dY[m,n],  t[m,r] = t16[m,r] / tau[r]
                    gB = sink_B if sink_B is not None else torch.empty((r, N), dtype=torch.float32, device=dev)
                    _lib.gemm_tn(g16, t2, gB, alpha=lo['scaling'], alpha_dev=gmax1, j_scale=lo['inv_tmul_vec'],
                                 transposed_out=True, clamp_abs=10.0 if lo['qtype_B'] == 'log' else 0.0, accumulate=sink_B is not None)
                    if sink_B is not None:
                        gB = None

        if need_x:
            # the fp16 attention output feeding c_proj takes its gradient in fp16 straight from the epilogue
            gx = torch.empty((M, K), dtype=torch.float16 if ctx.x_dtype == torch.float16 else torch.float32, device=dev)
            if dt16 is not None and clamp_in == 0.0:
                # identity STE: base and LoRA input-gradients share one accumulator
                _lib.qgemm(g16, bw['WT_op'], M, K, N, gx, A2=dt16, B2=bw['lora']['A_kr_op'], K2=lo['rank'],
                           row_scale=eg, col_scale=bw['pk'])
            elif dt16 is not None:
                # log STE clamps only the gradient that flows through q_in(x)
                gl = torch.empty((M, K), dtype=torch.float32, device=dev)
                _lib.qgemm(dt16, bw['lora']['A_kr_op'], M, K, lo['rank'], gl, row_scale=eg, col_scale=bw['pk'])
                _lib.qgemm(g16, bw['WT_op'], M, K, N, gx, row_scale=eg, col_scale=bw['pk'], clamp_abs=clamp_in, C=gl)
            else:
                _lib.qgemm(g16, bw['WT_op'], M, K, N, gx, row_scale=eg, col_scale=bw['pk'], clamp_abs=clamp_in)
            gx = gx.view(ctx.x_shape).to(ctx.x_dtype)
        if need_w:
            # dW[n,k] = sum_m dY[m,n] q(x)[m,k];  q(x)[m,k] = a_q[m,k] * absorb[k]; gG = dY / gmax, |gG| <= 256
            gG = _to_f16_operand(g2d.float(), row_mul=(1.0 / gmax1).expand(M).contiguous())
            gw = torch.empty((N, K), dtype=torch.float32, device=dev)
            _lib.gemm_tn(gG, a_q, gw, alpha=1.0, alpha_dev=gmax1, j_scale=act['absorb'],
                         clamp_abs=10.0 if ctx.weight_qtype == 'log' else 0.0)
        if ctx.has_bias and need_b:
            gb = g2d.float().sum(dim=0)
        return gx, gw, gb, gA, gB, None, None, None, None, (gy if ctx.needs_input_grad[9] else None), None, None, None


class SPLinearWithLoRA(nn.Module):
    def __init__(self, in_features, out_features, bit_widths, lora_rank_per_bit, lora_alpha_per_bit,
                 quantizer_per_bit, eps=1e-5, per_channel=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.bit_widths = bit_widths
        self.lora_rank_per_bit = lora_rank_per_bit
        student_bits = [b for b in bit_widths if b < 32]
        # reference :71 -- starts at the second-largest configured width
        self.current_bits = sorted(bit_widths, reverse=True)[1]

        self.linear = nn.Linear(in_features, out_features, bias=True)
        self.quantizers_weight = nn.ModuleDict({
            f'{bits}bit': LearnableFakeQuantize(num_bits=bits, quantizer_type=quantizer_per_bit[bits],
                                                channel_dim=0, eps=eps, per_channel=per_channel)
            for bits in student_bits})
        self.quantizers_input = nn.ModuleDict({
            f'{bits}bit': LearnableFakeQuantize(num_bits=bits, quantizer_type=quantizer_per_bit[bits],
                                                channel_dim=-1, eps=eps, per_channel=per_channel, is_input=True)
            for bits in student_bits})
        self.lora_adapters = nn.ModuleDict({
            f'{bits}bit': LoRALayer(in_features, out_features, rank=lora_rank_per_bit[bits],
                                    alpha=lora_alpha_per_bit[bits], bits=bits,
                                    quantizer_type=quantizer_per_bit[bits], eps=eps, per_channel=per_channel)
            for bits in student_bits})

        self.register_buffer('weight_quantized', torch.empty(out_features, in_features))
        self.register_buffer('input_quantized', None)
        self.calibration_mode = False

        # operand caches (not module state): bits -> {'base': {...}, 'lora': {...}, 'bwd': {...}}.
        # Entries are replaced, never mutated, so an autograd node can keep the ones it used.
        self._op_cache = {}
        self._fp_cache = _FpWeightCache()   # 32-bit path
        self._calib_cache = {}              # bits -> (key, dequantised weight, _FpWeightCache)

    # ---------------------------------------------------------------- host state (reference :105-125)
    def set_precision(self, bits) -> int:
        if bits >= 32:
            self.current_bits = 32
            return bits
        self.current_bits = bits
        bits_key = f'{bits}bit'
        self.quantizers_weight[bits_key].set_num_bits(bits)
        self.quantizers_input[bits_key].set_num_bits(bits)
        lora = self.lora_adapters[bits_key]
        if lora.quantize_A is not None:
            lora.quantize_A.set_num_bits(bits)
        if lora.quantize_B is not None:
            lora.quantize_B.set_num_bits(bits)
        return self.current_bits

    def get_active_lora(self):
        return self.lora_adapters[f'{self.current_bits}bit']

    # ---------------------------------------------------------------- operand caches
    # Three levels, so that each event only rebuilds what depends on it:
    #   weight level  (W, weight quantiser)          : q(W) fp32, its row absmax; backward: column scale, W^T operand
    #   lora level    (A, B, their quantisers)       : |q(A)|, A operand, q(B); backward: B operand
    #   input level   (+ input quantiser calibration): per-K absorb factor, activation multiplier, row
    #                                                  normaliser, W operand, LoRA-B operand -- 3 launches
    def _weight_level(self, bits):
        qw = self.quantizers_weight[f'{bits}bit']
        W = self.linear.weight
        ent = self._op_cache.setdefault(bits, {})
        key = (W.data_ptr(), W._version, qw.generation)
        wl = ent.get('weight')
        if wl is None or wl['key'] != key:
            with torch.no_grad():
                wq = _dequant(qw, W)
                wl = ent['weight'] = dict(key=key, wq=wq, wmax_row=wq.abs().amax(dim=1).contiguous(), bwd=None)
        return wl

    def _weight_level_bwd(self, bits):
        wl = self._weight_level(bits)
        if wl['bwd'] is None:
            with torch.no_grad():
                pk = _norm_pow2(wl['wq'].abs().amax(dim=0))                                  # [K]
                wl['bwd'] = dict(pk=pk, inv_pk=(1.0 / pk).contiguous(),
                                 WT_op=_to_f16_operand(wl['wq'], col_mul=(1.0 / pk).contiguous(), transposed=True))   # [K, N]
        return wl['bwd']

    def _lora_level(self, bits):
        lo = self.lora_adapters[f'{bits}bit']
        ent = self._op_cache.setdefault(bits, {})
        key = (lo.lora_A.data_ptr(), lo.lora_A._version, lo.lora_B.data_ptr(), lo.lora_B._version,
               lo.quantize_A.generation, lo.quantize_B.generation)
        ll = ent.get('lora')
        if ll is None or ll['key'] != key:
            for qq in (lo.quantize_A, lo.quantize_B):
                if not qq.ready():
                    # same exception the reference raises from quantize_A/B (p1/lora.py:49-50)
                    raise RuntimeError(
                        f"Quantizer not calibrated. Please run calibration first for {qq.quantizer_type} quantizer.")
            with torch.no_grad():
                K, r = lo.lora_A.shape
                aq = _dequant(lo.quantize_A, lo.lora_A)                                      # [K, r]
                bq = _dequant(lo.quantize_B, lo.lora_B)                                      # [r, N]
            ll = ent['lora'] = dict(key=key, rank=r, aq=aq, aq_abs=aq, bq=bq,
                                    scaling=float(lo.scaling), qtype_A=lo.quantize_A.quantizer_type,
                                    qtype_B=lo.quantize_B.quantizer_type, bwd=None, bwd_key=None)
        return ll

    def _operands_for(self, bits, want_lora):
        qi = self.quantizers_input[f'{bits}bit']
        K, N = self.in_features, self.out_features
        wl = self._weight_level(bits)
        ll = self._lora_level(bits) if want_lora else None
        ent = self._op_cache[bits]
        bkey = (wl['key'], qi.generation)
        lkey = None if ll is None else (bkey, ll['key'])
        il = ent.get('input')
        base = il['base'] if (il is not None and il['base']['key'] == bkey) else None
        lora = il['lora'] if (il is not None and il['lora'] is not None and il['lora']['key'] == lkey) else None
        if base is not None and (ll is None or lora is not None):
            return base, (lora if ll is not None else None)
        dev = wl['wq'].device
        with torch.no_grad():
            sc = qi.scale.detach().float().reshape(-1).contiguous()
            zp = qi.zero_point.detach().float().reshape(-1)
            if zp.numel() != sc.numel():
                zp = zp.expand_as(sc)
            zp = zp.contiguous()
            if sc.numel() not in (1, K):
                raise NotImplementedError(f"input quantiser scale of {sc.numel()} elements for in_features={K}")
            qtype = _lib.QTYPE[qi.quantizer_type]
            # one launch computes every scale vector (the base part is recomputed identically when only
            # the LoRA level changed; it is K + N elements)
            vec = torch.empty(4 * K + 2 * N, dtype=torch.float32, device=dev)
            absorb, act_mul, raw_mul, inv_raw_mul = vec[:K], vec[K:2 * K], vec[2 * K:3 * K], vec[3 * K:4 * K]
            pw, inv_pw = vec[4 * K:4 * K + N], vec[4 * K + N:]
            r = 0 if ll is None else ll['rank']
            lora_vec = torch.empty(8 * r, dtype=torch.float32, device=dev) if ll is not None else None
            # pb (normaliser of scaling * q(B)) does not depend on the input quantiser: computed by the first
            # prep after the LoRA level changed, reused by the recalibrations that follow
            want_pb = ll is not None and ll.get('pb') is None
            _lib.prep_linear_scales(sc, zp, qtype, qi.num_bits, qi.symmetric, K, wl['wmax_row'], N,
                                    None if ll is None else ll['aq_abs'], ll['bq'] if want_pb else None, r,
                                    0.0 if ll is None else ll['scaling'],
                                    absorb, act_mul, raw_mul, inv_raw_mul, pw, inv_pw, lora_vec)
            if base is None:
                if qi.quantizer_type == 'minmax':
                    kind, col_mul, mul = _lib.OPERAND_CODE, None, 2.0 ** -max(0, qi.num_bits - 11)
                else:
                    kind, col_mul, mul = _lib.OPERAND_DEQUANT, act_mul, 1.0
                act = dict(scale=sc, zp=zp, bcast=_lib.PER_TENSOR if sc.numel() == 1 else _lib.PER_COL, kind=kind,
                           col_mul=col_mul, mul=mul, absorb=absorb, raw_mul=raw_mul, inv_raw_mul=inv_raw_mul, qtype=qtype,
                           bits=qi.num_bits, symmetric=qi.symmetric, input_qtype=qi.quantizer_type)
                base = dict(key=bkey, act=act, pw=pw, inv_pw=inv_pw,
                            B_op=_to_f16_operand(wl['wq'], row_mul=inv_pw, col_mul=absorb))      # the big one: [N, K]
                base['f8'] = self._f8_level(bits, sc, qi)
            if ll is not None:
                if want_pb:
                    ll['pb'], ll['inv_pb'] = lora_vec[5 * r:6 * r].clone(), lora_vec[6 * r:7 * r].clone()
                tmul_vec, inv_tmul_vec, bl_rowmul = lora_vec[:r], lora_vec[r:2 * r], lora_vec[2 * r:3 * r]
                pa, inv_pa = lora_vec[3 * r:4 * r], lora_vec[4 * r:5 * r]
                # A operand of the down-projection: q(A)[k,j] / (raw_mul[k] pa[j]), stored [r, K]
                A_op = _to_f16_operand(ll['aq'], row_mul=base['act']['inv_raw_mul'], col_mul=inv_pa, transposed=True)
                Bl_op8 = None
                if base.get('f8') is not None:
                    # the epilogue multiplies by cs[n] = s_x * s_w[n] (not a power of two): the LoRA operand carries its inverse
                    b8 = (ll['bq'].t() * ll['scaling']) / (tmul_vec[None, :] * base['f8']['cs'][:, None])
                    Bl_op8 = _lib.empty_f16_padded(N, r, dev)
                    Bl_op8.copy_(b8.clamp(-65504.0, 65504.0))
                lora = dict(key=lkey, rank=r, A_op=A_op, pa=pa, tmul_vec=tmul_vec, inv_tmul_vec=inv_tmul_vec, Bl_op8=Bl_op8,
                            Bl_op=_to_f16_operand(ll['bq'], row_mul=bl_rowmul, col_mul=base['inv_pw'], transposed=True),   # [N, r]
                            scaling=ll['scaling'], qtype_A=ll['qtype_A'], qtype_B=ll['qtype_B'],
                            pb=ll['pb'], inv_pb=ll['inv_pb'], pa_tmul=lora_vec[7 * r:])
        ent['input'] = dict(base=base, lora=lora if ll is not None else (il['lora'] if il is not None and base is il['base'] else None))
        return base, (lora if ll is not None else None)

    def _f8_level(self, bits, in_scale, qi):
        """e4m3 operands of the integer-code GEMM, when the configuration allows one (north_star (b)/(c); the reference's
        evaluation loaders force per-tensor scales: p1/deploy.py:210,238, part3_eval_sp/main_sp_eval.py:60): symmetric
        min-max quantisers of <= 4 bits on both sides (codes in [-7, 7], exact in e4m3) and ONE input scale, so that
        q(x)[m,k] q(W)[n,k] = (s_x s_w[n]) * cx[m,k] cw[n,k] and the scales can wait for the epilogue.  K * 49 < 2^24 keeps
        the fp32 accumulation of the integer dot product exact.  `SPQ_FP8=0` disables the path (A/B switch)."""
        import os
        qw = self.quantizers_weight[f'{bits}bit']
        K, N = self.in_features, self.out_features
        ok = (qi.quantizer_type == 'minmax' and qw.quantizer_type == 'minmax' and qi.symmetric and qw.symmetric
              and qi.num_bits <= 4 and qw.num_bits <= 4 and in_scale.numel() == 1 and K % 16 == 0 and K * 49 < (1 << 24)
              and qw.scale.numel() in (1, N) and os.environ.get('SPQ_FP8', '1') != '0')
        if not ok:
            return None
        W = self.linear.weight.detach().float().contiguous()
        sw = qw.scale.detach().float().reshape(-1)
        zw = qw.zero_point.detach().float().reshape(-1)
        if zw.numel() != sw.numel():
            zw = zw.expand_as(sw)
        B8 = torch.empty((N, K), dtype=torch.uint8, device=W.device)
        _lib.fake_quantize(W, sw.contiguous(), zw.contiguous(), _lib.PER_TENSOR if sw.numel() == 1 else _lib.PER_ROW, _lib.MINMAX,
                           qw.num_bits, True, operand=B8, operand_kind=_lib.OPERAND_CODE_E4M3)
        return dict(B8=B8, cs=(in_scale.reshape(1) * sw).expand(N).contiguous())

    def _backward_operands_for(self, bits, want_lora):
        wb = self._weight_level_bwd(bits)
        bw = dict(pk=wb['pk'], WT_op=wb['WT_op'], lora=None)
        if want_lora:
            _, lora = self._operands_for(bits, True)
            ll = self._lora_level(bits)
            bkey = (self._weight_level(bits)['key'], lora['key'])
            if ll['bwd'] is None or ll['bwd_key'] != bkey:
                with torch.no_grad():
                    N = self.out_features
                    dt_mul = 2.0 ** -max(0, math.ceil(math.log2(max(N, 2))) - 7)
                    ll['bwd'] = dict(
                        pb=lora['pb'], dt_mul=dt_mul,
                        B_rn_op=_to_f16_operand(ll['bq'], row_mul=lora['inv_pb'], mul=ll['scaling']),                 # [r, N]
                        A_kr_op=_to_f16_operand(ll['aq'], row_mul=wb['inv_pk'], mul=1.0 / dt_mul))                      # [K, r]
                    ll['bwd_key'] = bkey
            bw['lora'] = ll['bwd']
        return bw

    def _restamp_lora_keys(self, bits, with_bwd=True):
        """After a CUDA-graph replay rewrote the LoRA-level tensors in place (training.LoRARefresher): mark the
        cached levels as built from the current parameter versions / quantiser generations."""
        lo = self.lora_adapters[f'{bits}bit']
        ent = self._op_cache[bits]
        ll, il = ent['lora'], ent['input']
        ll['key'] = (lo.lora_A.data_ptr(), lo.lora_A._version, lo.lora_B.data_ptr(), lo.lora_B._version,
                     lo.quantize_A.generation, lo.quantize_B.generation)
        il['lora']['key'] = (il['base']['key'], ll['key'])
        if with_bwd and ll['bwd'] is not None:
            ll['bwd_key'] = (self._weight_level(bits)['key'], il['lora']['key'])

    def _calibration_weight(self, bits, weight_quantizer):
        """q_w(W) and its fp16 operand cache for the calibration pass (inputs not quantised yet)."""
        W = self.linear.weight
        key = (W.data_ptr(), W._version, weight_quantizer.generation, weight_quantizer.collecting_stats)
        ent = self._calib_cache.get(bits)
        if ent is None or ent[0] != key:
            with torch.no_grad():
                ent = (key, weight_quantizer(W), _FpWeightCache())
            self._calib_cache[bits] = ent
        return ent[1], ent[2]

    # ---------------------------------------------------------------- forward (reference :127-150)
    def forward(self, x, out_half=False, fuse_gelu=False, residual=None, pre_norm=None):
        """Reference signature is forward(x) -> float32.  Two internal extensions used by the model
        wrapper: `out_half=True` (SPAttention with fp16 attention) stores fp16 from the GEMM epilogue
        instead of float32 followed by a cast; `fuse_gelu=True` (SPMLP under no_grad) applies the exact
        erf GELU in the epilogue instead of a separate elementwise pass; `residual` (SPBlock's residual
        stream) returns residual + forward(x), the add done in the GEMM epilogue (with autograd on, the residual's
        gradient is the output gradient)."""
        # `pre_norm` (SPBlock: the SwitchableLayerNorm in front of c_attn / c_fc): forward(pre_norm(x)), with the
        # normalisation done inside this layer's activation-side kernel whenever the fused kernels apply (no_grad,
        # float32 contiguous input); otherwise it is simply applied first
        if pre_norm is not None and not _ln_fusable(x, pre_norm, self.in_features):
            x, pre_norm = pre_norm(x), None
        if residual is not None:
            fuse_res = (not out_half and not fuse_gelu and residual.is_cuda
                        and residual.dtype == torch.float32 and residual.is_contiguous()
                        and residual.shape == x.shape[:-1] + (self.linear.out_features,))
            if not fuse_res:
                return residual + self.forward(x, out_half=out_half, fuse_gelu=fuse_gelu, pre_norm=pre_norm)
        act = 1 if (fuse_gelu and not torch.is_grad_enabled()) else 0
        post_gelu = fuse_gelu and not act
        if self.current_bits >= 32:
            half_here = out_half and not post_gelu
            pre = _ln_rowscaled_f16(x.reshape(-1, self.in_features), pre_norm) if pre_norm is not None else None
            y = linear_fp(x, self.linear.weight, self.linear.bias, self._fp_cache, activation=act, out_half=half_here,
                          residual=residual, pre=pre)
            if post_gelu:
                y = torch.nn.functional.gelu(y)
            return y.half() if (out_half and not half_here) else y

        bits_key = f'{self.current_bits}bit'
        if bits_key not in self.quantizers_weight or bits_key not in self.quantizers_input:
            raise KeyError(f"No weight quantizer for {bits_key}")
        weight_quantizer = self.quantizers_weight[bits_key]
        input_quantizer = self.quantizers_input[bits_key]
        active_lora = self.lora_adapters[bits_key]

        if input_quantizer.ready() and weight_quantizer.ready():
            lora_on = active_lora.enabled and active_lora.scaling != 0 and not self.calibration_mode
            # GELU under autograd with frozen base weight / bias: inside the Function (gelu'(y) then rides in the backward's
            # row-scaling pass); otherwise torch's GELU follows the call
            N_ = self.linear.out_features
            gelu_in_fn = (post_gelu and _FUSE_DGELU and not out_half and residual is None and N_ % 8 == 0 and N_ <= 8192
                          and not self.linear.weight.requires_grad
                          and (self.linear.bias is None or not self.linear.bias.requires_grad))
            y = _SPLinearFn.apply(x, self.linear.weight, self.linear.bias,
                                  active_lora.lora_A if lora_on else None,
                                  active_lora.lora_B if lora_on else None, self, self.current_bits, out_half, act,
                                  residual, torch.is_grad_enabled(), pre_norm, gelu_in_fn)
            return torch.nn.functional.gelu(y) if (post_gelu and not gelu_in_fn) else y

        # A quantiser is collecting statistics or is uncalibrated: compose the same steps as the
        # reference, module by module (this is the calibration pass; errors surface as upstream).
        pre = None
        if pre_norm is not None:
            # calibration pass behind a LayerNorm: ONE kernel normalises the rows, folds the input quantiser's per-column
            # statistics and writes the fp16 operand of the GEMM below (the float32 normalised rows are never stored)
            # (with LoRA enabled the adapter needs the normalised rows themselves: not fused)
            tg = (input_quantizer._stats_targets_lastdim(tuple(x.shape), x.device)
                  if (self.calibration_mode and input_quantizer.collecting_stats and input_quantizer.num_bits < 32
                      and input_quantizer.quantizer_type in ('minmax', 'log')) else None)
            if tg is None:
                x, pre_norm = pre_norm(x), None
            else:
                mode = 2 if input_quantizer.quantizer_type == 'log' else 1
                pre = _ln_rowscaled_f16(x.reshape(-1, self.in_features), pre_norm,
                                        stats=(mode, input_quantizer.eps, tg[0], tg[1], tg[2], tg[3]))
        x_quantized = input_quantizer(x) if pre is None else x   # collecting: records stats, returns x
        if torch.is_grad_enabled() and self.linear.weight.requires_grad and not weight_quantizer.collecting_stats:
            weight_quantized, cache = weight_quantizer(self.linear.weight), None
        else:
            weight_quantized, cache = self._calibration_weight(self.current_bits, weight_quantizer)
        fuse_here = act and self.calibration_mode            # GELU follows the LoRA add when LoRA is on
        half_here = (out_half and self.calibration_mode and not (fuse_gelu and not fuse_here)
                     and not torch.is_grad_enabled())
        res_here = residual if self.calibration_mode else None      # otherwise after the LoRA add, as upstream
        base_output = linear_fp(x_quantized, weight_quantized, self.linear.bias, cache, activation=1 if fuse_here else 0,
                                out_half=half_here, residual=res_here, pre=pre)
        if not self.calibration_mode:
            base_output = base_output + active_lora(x)
            if residual is not None:
                base_output = residual + base_output
        if fuse_gelu and not fuse_here:
            base_output = torch.nn.functional.gelu(base_output)
        input_quantizer.join_stats()          # a statistics pass issued on a side stream reads x: join before x can go
        return base_output.half() if (out_half and not half_here) else base_output
