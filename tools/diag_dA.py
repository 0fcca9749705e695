"""Diagnostic: step-by-step check of the dA path of _SPLinearFn.backward on synthetic data (N sweep)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import SPLinearWithLoRA, _lib
from llm_qat_on_gpt2_b200.lora import _rowscaled_f16, _to_f16_operand

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())

torch.manual_seed(0)
for (K, N, M, xkind) in [(768, 768, 192, "randn"), (768, 2304, 192, "randn"), (768, 3072, 192, "randn"), (768, 2304, 192, "ln"), (768, 2304, 4096, "ln"), (768, 768, 192, "ln")]:
    bits, r = 16, 64
    m = SPLinearWithLoRA(K, N, bit_widths=[bits, 32], lora_rank_per_bit={bits: r, 32: 0}, lora_alpha_per_bit={bits: r, 32: 0},
                         quantizer_per_bit={bits: "log", 32: None}).cuda()
    key = f"{bits}bit"
    lo = m.lora_adapters[key]
    with torch.no_grad():
        m.linear.weight.normal_(0, 0.02); lo.lora_B.normal_(0, 0.02)
    m.set_precision(bits)
    qw = m.quantizers_weight[key]; qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
    for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
        qq.start_calibration(); qq(w.data); qq.finish_calibration()
    if xkind == "ln":
        h = torch.randn(M, K, device="cuda") * 0.03
        x = torch.nn.functional.layer_norm(h, (K,)) * (1 + 0.1 * torch.randn(K, device="cuda")) + 0.05 * torch.randn(K, device="cuda")
    else:
        x = torch.randn(M, K, device="cuda")
    qi = m.quantizers_input[key]; qi.start_calibration(); m.calibration_mode = True
    with torch.no_grad():
        m(x[None])
    m.calibration_mode = False; qi.finish_calibration()
    m.linear.weight.requires_grad_(False); m.linear.bias.requires_grad_(False)
    gy = torch.randn(M, N, device="cuda") * 1e-3 * torch.exp(torch.randn(M, 1, device="cuda"))
    xg = x.clone().requires_grad_(True)
    y = m(xg[None]); y.backward(gy[None])
    with torch.no_grad():
        Aq = lo.quantize_A(lo.lora_A); Bq = lo.quantize_B(lo.lora_B)
        dt = (gy @ Bq.t()) * lo.scaling
        dA = x.t() @ dt
        t = x @ Aq
        dB = (t.t() @ gy) * lo.scaling
        # step by step with the module's own operands
        base, lora = m._operands_for(bits, True)
        bw = m._backward_operands_for(bits, True)
        lb = bw['lora']; act = base['act']
        g16, eg = _rowscaled_f16(gy)
        dtn = torch.empty((M, r), device="cuda")
        _lib.qgemm(g16, lb['B_rn_op'], M, r, N, dtn, col_scale=lb['pb'])
        e_dtn = rel(dtn * eg[:, None], dt)
        gmax = eg.max(); egn = eg / gmax
        dt2 = _to_f16_operand(dtn, row_mul=egn, mul=lb['dt_mul'])
        e_dt2 = rel(dt2.float() / lb['dt_mul'] * gmax, dt)
        a_q = torch.empty((M, K), dtype=torch.float16, device="cuda"); a_raw = torch.empty_like(a_q)
        _lib.quantize_act(x, act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'], act['symmetric'], act['kind'], act['col_mul'], act['mul'], a_q, a_raw, act['raw_mul'])
        e_araw = rel(a_raw.float() * act['inv_raw_mul'][None, :], x)
        gA_emul = (a_raw.float() * act['inv_raw_mul'][None, :]).t() @ (dt2.float() / lb['dt_mul'] * gmax)
        e_emul = rel(gA_emul, dA)
        gA = torch.empty((K, r), device="cuda")
        _lib.gemm_tn(a_raw, dt2, gA, alpha=1.0 / lb['dt_mul'], alpha_dev=gmax.reshape(1), i_scale=act['inv_raw_mul'])
        e_k = rel(gA, dA)
    print(f"K{K} N{N} M{M} {xkind}: module dA {rel(lo.lora_A.grad, dA):.2e} dB {rel(lo.lora_B.grad, dB):.2e} | dtn {e_dtn:.2e} dt2 {e_dt2:.2e} a_raw {e_araw:.2e} emul(fp32 of operands) {e_emul:.2e} kernel {e_k:.2e} | raw_mul range {float(act['raw_mul'].min())}-{float(act['raw_mul'].max())} dt_mul {lb['dt_mul']} a_raw absmax {float(a_raw.abs().max())} dt2 absmax {float(dt2.abs().max()):.3g} rms {float(dt2.float().pow(2).mean().sqrt()):.3g}")
