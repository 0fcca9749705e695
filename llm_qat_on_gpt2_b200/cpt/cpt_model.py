"""CPTLinear and the CPT GPT-2 wrapper on the B200 kernels -- drop-ins for
part2_cyclic_precision_training/cpt_model.py (LoRAAdapter :11, CPTLinear :37, CPTSelfAttention :116,
CPTBlock :170, CPTModel :206).

CPTLinear differs from part1's SPLinearWithLoRA in four ways, all kept: ONE shared LoRA adapter for
every width (`shared_lora`, B stored [out, rank]); the adapter reads the QUANTISED input
(`x_quant @ q(A) @ q(B).T`, :108-113); one `lora_weight_quantizers['{b}bit']` quantiser object is used
for both A and B; and the LoRA gradients pass through `GradientQuantizer` (8-bit min-max fake
quantisation of dA / dB once calibrated, p2/quantization.py:14-26).  The LM head is a CPTLinear
(768 -> 50257, no bias), LayerNorm is stock nn.LayerNorm.

Fused path: as in lora.py -- one pass over x produces the fp16 code/dequant operand, which here
also feeds the LoRA down-projection (its per-K factor is absorbed into the A operand), the
up-projection rides in the same TMEM accumulator as the base GEMM.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ..lora import (_FpWeightCache, _GradSide, _act_config, _as_2d_f32, _as_2d_grad, _dequant, _grad_sink, _norm_pow2, _quantized_operand,
                    _rowscaled_f16, _to_f16_operand, linear_fp)
from ..quantization import pow2_ceil
from .quantization import GradientQuantizer, LearnableFakeQuantize


class LoRAAdapter(nn.Module):
    def __init__(self, in_features: int, out_features: int, rank: int = 16, alpha: float = 32,
                 num_bits: int = 8, quantizer_type: str = 'log', gradient_bits: int = 8):
        super().__init__()
        self.rank = rank
        self.alpha = alpha
        self.scaling = alpha / rank if rank > 0 else 1.0
        if rank > 0:
            self.lora_A = nn.Parameter(torch.empty(in_features, rank))
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            self.lora_B = nn.Parameter(torch.zeros(out_features, rank))
            self.grad_quantizer_A = LearnableFakeQuantize(num_bits=gradient_bits, quantizer_type='minmax',
                                                          channel_dim=0, per_channel=True)
            self.grad_quantizer_B = LearnableFakeQuantize(num_bits=gradient_bits, quantizer_type='minmax',
                                                          channel_dim=0, per_channel=True)
        else:
            self.lora_A = None
            self.lora_B = None
            self.grad_quantizer_A = None
            self.grad_quantizer_B = None
        self.calibration_mode = False


def _grad_quantize(q: Optional[LearnableFakeQuantize], g: torch.Tensor) -> torch.Tensor:
    """GradientQuantizer.backward (p2/quantization.py:19-26) on an already computed gradient."""
    if q is not None and (q.collecting_stats or q.num_bits in q.calibrated_bits):
        with torch.no_grad():
            return q(g)
    return g


def _grad_quantizer_scale(q: Optional[LearnableFakeQuantize], rows: int):
    """(per-row scale for the fused epilogue or None, whether the whole GradientQuantizer + STE tail is fused).
    Fusable: no gradient quantiser / not active (identity), or calibrated symmetric min-max with one scale per row."""
    if q is None or not (q.collecting_stats or q.num_bits in q.calibrated_bits):
        return None, True
    if q.collecting_stats or q.quantizer_type != 'minmax' or not q.symmetric or q.num_bits >= 32:
        return None, False
    sc = q.scales[q.num_bits]
    if sc.numel() != rows:
        return None, False
    return sc.detach().float().reshape(-1).contiguous(), True


class _CPTLinearFn(torch.autograd.Function):
    """Fused forward / backward of CPTLinear at a quantised width."""

    @staticmethod
    def forward(ctx, x, weight, bias, lora_A, lora_B, mod, bits, ce_targets=None):
        # ce_targets (int64 [M], -100 = not scored): the LM head of CPTModel with labels.  The Function then returns
        # (mean cross-entropy, logits): the loss kernel leaves the fp16 gradient operand of this layer's backward, so
        # torch's log_softmax / nll passes, the float32 [M, V] dlogits and their row-scaling pass never exist
        use_lora = lora_A is not None
        base, lo = mod._operands_for(bits, use_lora)
        act = base['act']
        N, K = weight.shape
        x2d = _as_2d_f32(x, K)
        M = x2d.shape[0]
        a_q = torch.empty((M, K), dtype=torch.float16, device=x.device)
        _lib.quantize_act(x2d, act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'], act['symmetric'],
                          act['kind'], act['col_mul'], act['mul'], a_q, None, None)
        # with the fused loss: odd widths (the vocabulary) get rows padded to 128 bytes so that the GEMM stores through
        # TMA (as lora.linear_fp does); the plain call keeps the dense [.., N] tensor upstream's callers may .view()
        ld = N if (N % 4 == 0 or ce_targets is None) else (N + 31) // 32 * 32
        ybuf = torch.empty((M, ld), dtype=torch.float32, device=x.device)
        y = ybuf[:, :N] if ld != N else ybuf
        bias_f = None if bias is None else bias.detach().float().contiguous()
        t16 = None
        if use_lora:
            r = lo['rank']
            # t = q(x) q(A), stored as the fp16 operand of the up-projection (pa and tau are powers of two)
            t16 = _lib.empty_f16_padded(M, r, x.device)
            _lib.qgemm(a_q, lo['A_op'], M, r, K, t16, col_scale=lo['pa_tmul'])
            _lib.qgemm(a_q, base['B_op'], M, N, K, y, A2=t16, B2=lo['Bl_op'], K2=r, col_scale=base['pw'], bias=bias_f)
        else:
            _lib.qgemm(a_q, base['B_op'], M, N, K, y, col_scale=base['pw'], bias=bias_f)
        need = ctx.needs_input_grad
        ctx.use_lora, ctx.x_shape, ctx.has_bias, ctx.dims = use_lora, x.shape, bias is not None, (M, N, K)
        ctx.base, ctx.lo, ctx.mod = base, lo, mod
        ctx.bits = bits                     # backward operands are built (and cached) when backward first runs
        ctx.weight_qtype = mod.quantizer_weight.quantizer_type
        keep_aq = need[1] or (use_lora and need[3])
        out = ybuf.view(*x.shape[:-1], ld)
        out = out[..., :N] if ld != N else out
        ctx.fused_ce = ce_targets is not None
        if ce_targets is None:
            ctx.save_for_backward(a_q if keep_aq else None, t16 if (use_lora and need[4]) else None)
            return out
        row_loss, row_valid, g16, eg, gmax1 = _lib.softmax_loss_grad16(y, 'ce', targets=ce_targets)
        factor = (1.0 / row_valid.sum()).reshape(1)            # mean over the scored rows, as F.cross_entropy
        loss = row_loss.sum() * factor[0]
        ctx.save_for_backward(a_q if keep_aq else None, t16 if (use_lora and need[4]) else None, g16, eg, gmax1, factor)
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, gy, *unused):
        base, lo, mod = ctx.base, ctx.lo, ctx.mod
        bw = mod._backward_operands_for(ctx.bits, ctx.use_lora)
        M, N, K = ctx.dims
        dev = gy.device
        act = base['act']
        if ctx.fused_ce:
            # gy = d(loss): dY = g16 * eg[:, None] * (gy / n_scored) -- the scalar goes into the row scales
            a_q, t16, g16, eg, gmax1, factor = ctx.saved_tensors
            scal = (gy.reshape(1).float() * factor)
            eg, gmax1 = eg * scal, gmax1 * scal
            g2d = None                                       # (weight / bias gradients are not taken on this path)
        else:
            a_q, t16 = ctx.saved_tensors
            g2d = _as_2d_grad(gy, N)
            g16 = _lib.empty_f16_padded(M, N, dev)
            eg = torch.empty(M, dtype=torch.float32, device=dev)
            gmax1 = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.rowscale_f16_max(g2d, g16, eg, gmax1)       # dY = g16 * eg[:,None], gmax1 = max eg
        gx = gw = gb = gA = gB = None
        need_x, need_w, need_b, need_A, need_B = ctx.needs_input_grad[:5]
        clamp_in = 10.0 if act['input_qtype'] == 'log' else 0.0
        dt16 = None
        if ctx.use_lora and (need_x or need_A or need_B):
            lb = bw['lora']
            r = lo['rank']
            clamp_w = 10.0 if lo['qtype'] == 'log' else 0.0
            dtn = None
            if need_x or need_A:
                dtn = torch.empty((M, r), dtype=torch.float32, device=dev)       # dt / eg,  dt = scaling * dY q(B)
                _lib.qgemm(g16, lb['B_rn_op'], M, r, N, dtn, col_scale=lb['pb'])
            dt16, dt2, t2 = _lib.lora_bwd_prep(dtn, t16 if need_B else None, eg, gmax1, lb['dt_mul'],
                                               want_dt16=need_x, want_dt2=need_A, want_t2=need_B)
            # GradientQuantizer (p2/quantization.py:14-26): calibrated -> fused into the gradient GEMM's fold pass in
            # front of the weight quantiser's STE clamp; collecting statistics -> the module call records them
            sl = mod.shared_lora
            # gradients that accumulate straight into a driver-owned buffer may run on the driver's side stream (lora._GradSide)
            fa = _grad_quantizer_scale(sl.grad_quantizer_A, K)[1] if need_A else True
            fb = _grad_quantizer_scale(sl.grad_quantizer_B, N)[1] if need_B else True
            on_side = ((not need_A or (fa and _grad_sink(sl.lora_A, (K, r)) is not None)) and
                       (not need_B or (fb and _grad_sink(sl.lora_B, (N, r)) is not None)))
            lane = _GradSide.fork(a_q, dt2, t2, g16, gmax1) if on_side else torch.cuda.current_stream()
            with torch.cuda.stream(lane):
                if need_A:
                    # dA[k,j] = sum_m q(x)[m,k] dt[m,j],  q(x)[m,k] = a_q[m,k] * absorb[k]
                    gqs, fused = _grad_quantizer_scale(sl.grad_quantizer_A, K)
                    sink = _grad_sink(sl.lora_A, (K, r)) if fused else None
                    gA = sink if sink is not None else torch.empty((K, r), dtype=torch.float32, device=dev)
                    _lib.gemm_tn(a_q, dt2, gA, alpha=1.0 / lb['dt_mul'], alpha_dev=gmax1, i_scale=act['absorb'],
                                 clamp_abs=clamp_w if fused else 0.0, gq_scale_i=gqs,
                                 gq_bits=sl.grad_quantizer_A.num_bits if gqs is not None else 8, accumulate=sink is not None)
                    if sink is not None:
                        gA = None
                    elif not fused:
                        gA = _grad_quantize(sl.grad_quantizer_A, gA)
                        if clamp_w:
                            gA = _lib.ste_backward(gA, _lib.LOG)
                if need_B:
                    # dB[n,j] = scaling * sum_m dY[m,n] t[m,j],  t[m,j] = t16[m,j] / tau
                    gqs, fused = _grad_quantizer_scale(sl.grad_quantizer_B, N)
                    sink = _grad_sink(sl.lora_B, (N, r)) if fused else None
                    gB = sink if sink is not None else torch.empty((N, r), dtype=torch.float32, device=dev)
                    _lib.gemm_tn(g16, t2, gB, alpha=lo['scaling'], alpha_dev=gmax1, j_scale=lo['inv_tmul_vec'],
                                 clamp_abs=clamp_w if fused else 0.0, gq_scale_i=gqs,
                                 gq_bits=sl.grad_quantizer_B.num_bits if gqs is not None else 8, accumulate=sink is not None)
                    if sink is not None:
                        gB = None
                    elif not fused:
                        gB = _grad_quantize(sl.grad_quantizer_B, gB)
                        if clamp_w:
                            gB = _lib.ste_backward(gB, _lib.LOG)
        if need_x:
            # both terms flow through q_in(x) here, so the STE clamp applies to their sum
            gx = torch.empty((M, K), dtype=torch.float32, device=dev)
            if dt16 is not None:
                _lib.qgemm(g16, bw['WT_op'], M, K, N, gx, A2=dt16, B2=bw['lora']['A_kr_op'], K2=lo['rank'],
                           row_scale=eg, col_scale=bw['pk'], clamp_abs=clamp_in)
            else:
                _lib.qgemm(g16, bw['WT_op'], M, K, N, gx, row_scale=eg, col_scale=bw['pk'], clamp_abs=clamp_in)
            gx = gx.view(ctx.x_shape)
        if need_w:
            gG = _to_f16_operand(g2d.float(), row_mul=(1.0 / gmax1).expand(M).contiguous())
            gw = torch.empty((N, K), dtype=torch.float32, device=dev)
            _lib.gemm_tn(gG, a_q, gw, alpha=1.0, alpha_dev=gmax1, j_scale=act['absorb'],
                         clamp_abs=10.0 if ctx.weight_qtype == 'log' else 0.0)
        if ctx.has_bias and need_b:
            gb = g2d.float().sum(dim=0)
        return gx, gw, gb, gA, gB, None, None, None


class CPTLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bit_widths: list = [4, 6, 8],
                 quantizer_per_bit: dict = None, gradient_bits: int = 8, bias: bool = True,
                 shared_lora_rank: int = 16, shared_lora_alpha: int = 32):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.bit_widths = bit_widths
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.shared_lora = LoRAAdapter(in_features, out_features, rank=shared_lora_rank, alpha=shared_lora_alpha,
                                       num_bits=8, quantizer_type='log', gradient_bits=gradient_bits)
        if quantizer_per_bit is None:
            quantizer_per_bit = {bits: 'log' for bits in bit_widths}
        self.lora_weight_quantizers = nn.ModuleDict({
            f'{bits}bit': LearnableFakeQuantize(num_bits=bits, quantizer_type=quantizer_per_bit.get(bits, 'log'),
                                                channel_dim=1, per_channel=True)
            for bits in bit_widths})
        max_bits = max([b for b in bit_widths if b < 32]) if any(b < 32 for b in bit_widths) else 8
        max_quant_type = quantizer_per_bit.get(max_bits, 'log')
        self.quantizer_weight = LearnableFakeQuantize(num_bits=max_bits, quantizer_type=max_quant_type,
                                                      channel_dim=0, per_channel=True)
        self.quantizer_input = LearnableFakeQuantize(num_bits=max_bits, quantizer_type=max_quant_type,
                                                     channel_dim=-1, per_channel=True, is_input=True)
        self.current_bits = max(bit_widths)
        self.calibration_mode = False
        self._op_cache = {}
        self._fp_cache = _FpWeightCache()

    def set_precision(self, num_bits: int):
        if num_bits not in self.bit_widths:
            raise ValueError(f"Precision {num_bits} not in widths {self.bit_widths}")
        self.current_bits = num_bits
        if num_bits < 32:
            self.quantizer_weight.set_num_bits(num_bits)
            self.quantizer_input.set_num_bits(num_bits)

    # ---------------------------------------------------------------- operand caches (see lora.py)
    def _operands_for(self, bits, want_lora):
        """Three cache levels, as in lora.py: the base level (weight + both calibrations of this width) survives the
        per-step precision cycling and the optimizer steps; the shared-LoRA level is rebuilt whenever the adapter or
        the width's LoRA quantiser changed -- 2 dequantise launches, ONE scale kernel (spq_cpt_lora_scales) and 4
        operand builds, forward and backward operands together (the first version spent ~85 eager launches on it)."""
        qi, qw = self.quantizer_input, self.quantizer_weight
        W = self.linear.weight
        ent = self._op_cache.setdefault(bits, {'base': None, 'lora': None, 'bwd': None})
        base_key = (W.data_ptr(), W._version, qw.generation, qi.generation, bits)
        base = ent['base']
        if base is None or base['key'] != base_key:
            with torch.no_grad():
                act = _act_config(qi, self.in_features)
                wq = _dequant(qw, W)
                pw = _norm_pow2((wq.abs() * act['absorb']).amax(dim=1))
                B_op = _quantized_operand(qw, W, row_mul=1.0 / pw, col_mul=act['absorb'])
                xb = qi.abs_bound().detach().float().reshape(-1).expand(self.in_features).contiguous()
                pk = _norm_pow2(wq.abs().amax(dim=0))
            base = ent['base'] = dict(key=base_key, act=act, wq=wq, pw=pw, inv_pw=(1.0 / pw).contiguous(), B_op=B_op, xb=xb,
                                      pk=pk, inv_pk=(1.0 / pk).contiguous(), WT_op=None)
            ent['lora'] = ent['bwd'] = None
        if not want_lora:
            return base, None
        sl = self.shared_lora
        lq = self.lora_weight_quantizers[f'{bits}bit']
        lkey = (base_key, sl.lora_A.data_ptr(), sl.lora_A._version, sl.lora_B.data_ptr(), sl.lora_B._version, lq.generation)
        lora = ent['lora']
        if lora is None or lora['key'] != lkey:
            with torch.no_grad():
                K, r = sl.lora_A.shape
                N = self.out_features
                act = base['act']
                aq = _dequant(lq, sl.lora_A)                                   # [K, r]
                bq = _dequant(lq, sl.lora_B)                                   # [N, r]
                if 1024 % r == 0:
                    vec = torch.empty(8 * r, dtype=torch.float32, device=aq.device)
                    _lib.cpt_lora_scales(aq, bq, act['absorb'], base['xb'], float(sl.scaling), vec)
                    pa, inv_pa, tmul_vec, inv_tmul_vec = vec[:r], vec[r:2 * r], vec[2 * r:3 * r], vec[3 * r:4 * r]
                    pa_tmul, bl_colmul, pb, inv_pb = vec[4 * r:5 * r], vec[5 * r:6 * r], vec[6 * r:7 * r], vec[7 * r:]
                else:                                                          # ranks that do not divide 1024: torch ops
                    pa = _norm_pow2((aq.abs() * act['absorb'][:, None]).amax(dim=0), 0)
                    tmax = (base['xb'][:, None] * aq.abs()).sum(dim=0).max()
                    tmul = torch.where(tmax > 0, (2.0 ** 14) / pow2_ceil(tmax), torch.ones_like(tmax))
                    tmul_vec = tmul.expand(r).contiguous()
                    inv_pa, inv_tmul_vec = (1.0 / pa).contiguous(), (1.0 / tmul_vec).contiguous()
                    pa_tmul, bl_colmul = (pa * tmul_vec).contiguous(), (sl.scaling / tmul_vec).contiguous()
                    pb = _norm_pow2(bq.abs().amax(dim=0) * abs(sl.scaling), 0)
                    inv_pb = (1.0 / pb).contiguous()
                dt_mul = 2.0 ** -max(0, math.ceil(math.log2(max(N, 2))) - 7)
                lora = ent['lora'] = dict(
                    key=lkey, rank=r, pa=pa, pa_tmul=pa_tmul, tmul_vec=tmul_vec, inv_tmul_vec=inv_tmul_vec,
                    scaling=float(sl.scaling), qtype=lq.quantizer_type,
                    A_op=_to_f16_operand(aq, row_mul=act['absorb'], col_mul=inv_pa, transposed=True),        # [r, K]
                    Bl_op=_to_f16_operand(bq, row_mul=base['inv_pw'], col_mul=bl_colmul),                     # [N, r]
                    bwd=dict(pb=pb, dt_mul=dt_mul,
                             B_rn_op=_to_f16_operand(bq, col_mul=inv_pb, mul=float(sl.scaling), transposed=True),   # [r, N]
                             A_kr_op=_to_f16_operand(aq, row_mul=base['inv_pk'], mul=1.0 / dt_mul)))              # [K, r]
            ent['bwd'] = None
        return base, lora

    def _backward_operands_for(self, bits, want_lora):
        base, lora = self._operands_for(bits, want_lora)
        if base['WT_op'] is None:
            with torch.no_grad():
                base['WT_op'] = _quantized_operand(self.quantizer_weight, self.linear.weight, col_mul=base['inv_pk'], transposed=True)
        return dict(pk=base['pk'], WT_op=base['WT_op'], lora=None if lora is None else lora['bwd'])

    def forward_with_cross_entropy(self, x: torch.Tensor, targets: torch.Tensor):
        """(F.cross_entropy(self(x).view(-1, N), targets.view(-1), ignore_index=-100), self(x)) -- CPTModel's LM head with
        labels (p2/cpt_model.py: lm_head + the shifted next-token loss).  At a calibrated quantised width with frozen base
        weight / bias the loss kernel feeds this layer's backward directly (see _CPTLinearFn.forward); otherwise the two
        steps are simply composed."""
        qi, qw = self.quantizer_input, self.quantizer_weight
        lora_on = not self.calibration_mode and self.shared_lora.rank > 0
        lq = self.lora_weight_quantizers[f'{self.current_bits}bit'] if (lora_on and self.current_bits < 32) else None
        frozen = not (torch.is_grad_enabled() and (self.linear.weight.requires_grad or
                                                   (self.linear.bias is not None and self.linear.bias.requires_grad)))
        if (self.current_bits < 32 and x.is_cuda and frozen and qi.ready() and qw.ready() and (lq is None or lq.ready())
                and os.environ.get('SPQ_CPT_FUSED_CE', '1') != '0'):
            tg = targets.reshape(-1).to(torch.int64).contiguous()
            return _CPTLinearFn.apply(x, self.linear.weight, self.linear.bias,
                                      self.shared_lora.lora_A if lora_on else None,
                                      self.shared_lora.lora_B if lora_on else None, self, self.current_bits, tg)
        logits = self.forward(x)
        return F.cross_entropy(logits.reshape(-1, logits.size(-1)), targets.reshape(-1), ignore_index=-100), logits

    # ---------------------------------------------------------------- forward (p2/cpt_model.py:92-114)
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.current_bits == 32:
            return linear_fp(x, self.linear.weight, self.linear.bias, self._fp_cache)
        qi, qw = self.quantizer_input, self.quantizer_weight
        lora_on = not self.calibration_mode and self.shared_lora.rank > 0
        lq = self.lora_weight_quantizers[f'{self.current_bits}bit'] if lora_on else None
        if qi.ready() and qw.ready() and (lq is None or lq.ready()):
            return _CPTLinearFn.apply(x, self.linear.weight, self.linear.bias,
                                      self.shared_lora.lora_A if lora_on else None,
                                      self.shared_lora.lora_B if lora_on else None, self, self.current_bits)
        # calibration pass / uncalibrated width: the reference's composition, module by module
        x_quant = qi(x)
        weight_quant = qw(self.linear.weight)
        out = linear_fp(x_quant, weight_quant, self.linear.bias)
        if self.calibration_mode:
            return out
        lq = self.lora_weight_quantizers[f'{self.current_bits}bit']
        a_q = GradientQuantizer.apply(lq(self.shared_lora.lora_A), self.shared_lora.grad_quantizer_A)
        b_q = GradientQuantizer.apply(lq(self.shared_lora.lora_B), self.shared_lora.grad_quantizer_B)
        lora_output = linear_fp(linear_fp(x_quant, a_q.t()), b_q)
        return out + lora_output * self.shared_lora.scaling


class CPTSelfAttention(nn.Module):
    def __init__(self, config, bit_widths: list, quantizer_per_bit: dict = None, gradient_bits: int = 8,
                 shared_lora_rank: int = 16, shared_lora_alpha: int = 32):
        super().__init__()
        self.n_head = config.n_head
        self.n_embd = config.n_embd
        self.head_dim = self.n_embd // self.n_head
        kw = dict(shared_lora_rank=shared_lora_rank, shared_lora_alpha=shared_lora_alpha)
        self.c_attn = CPTLinear(self.n_embd, 3 * self.n_embd, bit_widths, quantizer_per_bit, gradient_bits, **kw)
        self.c_proj = CPTLinear(self.n_embd, self.n_embd, bit_widths, quantizer_per_bit, gradient_bits, **kw)
        self.attn_dropout = nn.Dropout(config.embd_pdrop)
        self.resid_dropout = nn.Dropout(config.embd_pdrop)
        # 'fp32': float32 q/k/v through torch SDPA (upstream computes the attention in float32); 'fp16': fused flash
        # attention on fp16 q/k/v (a library call either way: between two hot-path linears)
        self.attention_dtype = getattr(config, 'attention_dtype', 'fp32')

    def set_precision(self, num_bits: int):
        self.c_attn.set_precision(num_bits)
        self.c_proj.set_precision(num_bits)

    def forward(self, hidden_states, attention_mask=None, past_key_value=None):
        B, T, _ = hidden_states.shape
        q, k, v = self.c_attn(hidden_states).split(self.n_embd, dim=-1)
        q = q.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        k = k.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        v = v.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        if past_key_value is not None:
            k = torch.cat([past_key_value[0], k], dim=2)
            v = torch.cat([past_key_value[1], v], dim=2)
        present = (k, v)
        S = k.size(2)
        p_drop = self.attn_dropout.p if self.training else 0.0
        half = self.attention_dtype == 'fp16' and q.dtype == torch.float32
        if half:
            q, k, v = q.half(), k.half(), v.half()
        if attention_mask is None and S == T:
            o = F.scaled_dot_product_attention(q, k, v, is_causal=True, dropout_p=p_drop)
        else:
            causal = torch.tril(torch.ones(S, S, device=q.device, dtype=torch.bool))[-T:, :]
            bias = torch.zeros(T, S, device=q.device, dtype=q.dtype).masked_fill(~causal, float('-inf'))
            if attention_mask is not None:
                bias = bias + attention_mask
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=p_drop)
        if half:
            o = o.float()
        o = o.transpose(1, 2).contiguous().view(B, T, self.n_embd)
        return self.resid_dropout(self.c_proj(o)), present


CPTMLP = nn.ModuleDict          # the reference keeps the MLP as a ModuleDict {'fc_in', 'fc_out'} (:177-180)


class CPTLayerNorm(nn.LayerNorm):
    """nn.LayerNorm (upstream's class: same parameters, same state_dict keys) whose forward / backward run on this repo's
    row-resident LayerNorm kernels (csrc/spq_layernorm.cu) when the input is a CUDA tensor."""

    def forward(self, x):
        if not x.is_cuda or self.weight is None or self.bias is None:
            return super().forward(x)
        from ..switchable_batchnorm import _LayerNormFn
        return _LayerNormFn.apply(x, self.weight, self.bias, self.eps, int(self.normalized_shape[-1]))


class CPTBlock(nn.Module):
    def __init__(self, config, bit_widths: list, quantizer_per_bit: dict = None, gradient_bits: int = 8,
                 shared_lora_rank: int = 16, shared_lora_alpha: int = 32):
        super().__init__()
        self.ln_1 = CPTLayerNorm(config.n_embd, eps=config.layer_norm_epsilon)
        self.ln_2 = CPTLayerNorm(config.n_embd, eps=config.layer_norm_epsilon)
        self.bit_widths = bit_widths
        kw = dict(shared_lora_rank=shared_lora_rank, shared_lora_alpha=shared_lora_alpha)
        self.attn = CPTSelfAttention(config, bit_widths, quantizer_per_bit, gradient_bits, shared_lora_rank, shared_lora_alpha)
        self.mlp = nn.ModuleDict({
            'fc_in': CPTLinear(config.n_embd, 4 * config.n_embd, bit_widths, quantizer_per_bit, gradient_bits, **kw),
            'fc_out': CPTLinear(4 * config.n_embd, config.n_embd, bit_widths, quantizer_per_bit, gradient_bits, **kw)})
        self.mlp_dropout = nn.Dropout(config.embd_pdrop)

    def set_precision(self, num_bits: int):
        self.attn.set_precision(num_bits)
        self.mlp['fc_in'].set_precision(num_bits)
        self.mlp['fc_out'].set_precision(num_bits)

    def forward(self, hidden_states, attention_mask=None, past_key_value=None):
        a, present = self.attn(self.ln_1(hidden_states), attention_mask, past_key_value)
        hidden_states = hidden_states + a
        m = self.mlp['fc_out'](F.gelu(self.mlp['fc_in'](self.ln_2(hidden_states))))
        return hidden_states + self.mlp_dropout(m), present


class CPTModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        mc = config['model']
        self.wte = nn.Embedding(mc.vocab_size, mc.n_embd)
        self.wpe = nn.Embedding(mc.n_positions, mc.n_embd)
        self.drop = nn.Dropout(mc.embd_pdrop)
        self.h = nn.ModuleList([
            CPTBlock(mc, mc.bit_widths, mc.quantizer_per_bit, mc.gradient_bits, mc.shared_lora_rank, mc.shared_lora_alpha)
            for _ in range(mc.n_layer)])
        self.ln_f = CPTLayerNorm(mc.n_embd, eps=mc.layer_norm_epsilon)
        self.lm_head = CPTLinear(mc.n_embd, mc.vocab_size, mc.bit_widths, mc.quantizer_per_bit, mc.gradient_bits,
                                 bias=False, shared_lora_rank=mc.shared_lora_rank, shared_lora_alpha=mc.shared_lora_alpha)
        self.apply(self._init_weights)
        self.current_precision = config['training'].target_bits

    def _init_weights(self, module):
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=0.02)

    def set_precision(self, num_bits: int):
        self.current_precision = num_bits
        for block in self.h:
            block.set_precision(num_bits)
        self.lm_head.set_precision(num_bits)

    def forward(self, input_ids, attention_mask=None, past_key_values=None, labels=None, use_cache=False):
        from transformers.modeling_outputs import CausalLMOutputWithPast
        B, T = input_ids.shape
        pos = torch.arange(T, dtype=torch.long, device=input_ids.device).unsqueeze(0).expand(B, -1)
        hidden_states = self.drop(self.wte(input_ids) + self.wpe(pos))
        if attention_mask is not None:
            attention_mask = (1.0 - attention_mask.unsqueeze(1).unsqueeze(2)) * -10000.0
        presents = [] if use_cache else None
        for i, block in enumerate(self.h):
            pkv = past_key_values[i] if past_key_values is not None else None
            hidden_states, present = block(hidden_states, attention_mask, pkv)
            if use_cache:
                presents.append(present)
        hidden_states = self.ln_f(hidden_states)
        loss = None
        if labels is not None:
            targets = torch.full_like(labels, -100)
            targets[..., :-1] = labels[..., 1:]
            loss, logits = self.lm_head.forward_with_cross_entropy(hidden_states, targets)
        else:
            logits = self.lm_head(hidden_states)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=presents,
                                      hidden_states=hidden_states, attentions=None)

    def disable_lora_for_calibration(self):
        for module in self.modules():
            if isinstance(module, (CPTLinear, LoRAAdapter)):
                module.calibration_mode = True

    def enable_lora_after_calibration(self):
        for module in self.modules():
            if isinstance(module, (CPTLinear, LoRAAdapter)):
                module.calibration_mode = False

    def generate(self, input_ids, max_length=100, temperature=1.0, do_sample=True, **kwargs):
        self.eval()
        with torch.no_grad():
            for _ in range(max_length - input_ids.size(1)):
                nxt = self.forward(input_ids, use_cache=False).logits[:, -1, :] / temperature
                token = torch.multinomial(F.softmax(nxt, dim=-1), num_samples=1) if do_sample \
                    else torch.argmax(nxt, dim=-1, keepdim=True)
                input_ids = torch.cat([input_ids, token], dim=1)
                if kwargs.get('eos_token_id') is not None and (token == kwargs['eos_token_id']).all():
                    break
        return input_ids
