"""Diagnostic: teacher-forced backward of every linear of a 2-layer model (upstream's x and dY fed to our layer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import upstream as up
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import test_gpu_upstream as T

torch.backends.cuda.matmul.allow_tf32 = False
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ref, ours, cfg = T._make_pair(2, (4, 8, 16), seed=3)
ids = torch.randint(0, cfg.vocab_size, (2, 96), generator=torch.Generator().manual_seed(5)).cuda()
cap = {}
hooks = []
for name, mod in ref.named_modules():
    if mod.__class__.__name__ == "SPLinearWithLoRA":
        hooks.append(mod.register_forward_hook(lambda m, i, o, _n=name: cap.setdefault(_n, {}).update(x=i[0].detach().clone(), y=o.detach().clone())))
        hooks.append(mod.register_full_backward_hook(lambda m, gi, go, _n=name: cap[_n].update(gy=go[0].detach().clone(), gx=None if gi[0] is None else gi[0].detach().clone())))
with up.quiet():
    ref.set_precision(bits); ours.set_precision(bits)
for n, p in ref.named_parameters():
    p.requires_grad_(f"lora_adapters.{bits}bit.lora_" in n or n.endswith(f"weights.{bits}") or n.endswith(f"biases.{bits}"))
emb = (ref.transformer.wte(ids) + ref.transformer.wpe(torch.arange(ids.shape[1], device="cuda")[None])).detach().requires_grad_(True)
with up.quiet():
    out = ref(inputs_embeds=emb, labels=ids)
out["loss"].backward()
for h in hooks:
    h.remove()
mine = dict(ours.named_modules())
refm = dict(ref.named_modules())
key = f"{bits}bit"
for name, c in cap.items():
    m = mine[name]
    lo = m.lora_adapters[key]
    for p in (lo.lora_A, lo.lora_B):
        p.requires_grad_(True); p.grad = None
    x = c["x"].clone().requires_grad_(True)
    y = m(x)
    y.backward(c["gy"])
    rl = refm[name].lora_adapters[key]
    e = dict(y=T.rel(y, c["y"]), dA=T.rel(lo.lora_A.grad, rl.lora_A.grad), dB=T.rel(lo.lora_B.grad, rl.lora_B.grad))
    if c["gx"] is not None:
        e["dx"] = T.rel(x.grad, c["gx"])
    # fp32 torch recomputation of dA from upstream's own tensors, to see conditioning
    with torch.no_grad(), up.quiet():
        rq = refm[name]
        Bq = rl.quantize_B(rl.lora_B)
        dt = (c["gy"].reshape(-1, c["gy"].shape[-1]) @ Bq.t()) * rl.scaling
        dA = c["x"].reshape(-1, c["x"].shape[-1]).t() @ dt
        e["dA_fp32_recompute"] = T.rel(dA, rl.lora_A.grad)
        gyn = c["gy"].reshape(-1, c["gy"].shape[-1]).abs().amax(dim=1)
        e["tok_grad_max/min"] = float(gyn.max() / gyn[gyn > 0].min())
        x2 = c["x"].reshape(-1, c["x"].shape[-1]); gy2 = c["gy"].reshape(-1, c["gy"].shape[-1])
        # plain fp16 rounding of x and dt (what any fp16-operand implementation, AMP included, does at best)
        dA16 = x2.half().float().t() @ (dt * 1024).half().float() / 1024
        e["dA_fp16_operands"] = T.rel(dA16, rl.lora_A.grad)
        # conditioning of the token reduction: |sum| vs sum of |terms|
        e["cancel"] = float((x2.abs().t() @ dt.abs()).norm() / dA.norm())
        # our pipeline step by step
        from llm_qat_on_gpt2_b200 import _lib
        from llm_qat_on_gpt2_b200.lora import _rowscaled_f16, _to_f16_operand
        base, lora = m._operands_for(bits, True); bw = m._backward_operands_for(bits, True); lb = bw['lora']; act = base['act']
        M, r, N = x2.shape[0], lora['rank'], gy2.shape[1]
        g16, eg = _rowscaled_f16(gy2.contiguous())
        dtn = torch.empty((M, r), device="cuda")
        _lib.qgemm(g16, lb['B_rn_op'], M, r, N, dtn, col_scale=lb['pb'])
        e["dtn"] = T.rel(dtn * eg[:, None], dt)
        gmax = eg.max(); egn = eg / gmax
        dt2 = _to_f16_operand(dtn, row_mul=egn, mul=lb['dt_mul'])
        e["dt2"] = T.rel(dt2.float() / lb['dt_mul'] * gmax, dt)
        a_q = torch.empty((M, x2.shape[1]), dtype=torch.float16, device="cuda"); a_raw = torch.empty_like(a_q)
        _lib.quantize_act(x2.contiguous(), act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'], act['symmetric'], act['kind'], act['col_mul'], act['mul'], a_q, a_raw, act['raw_mul'])
        e["a_raw"] = T.rel(a_raw.float() * act['inv_raw_mul'][None, :], x2)
        e["a_raw_absmax"] = float(a_raw.float().abs().max())
        e["emul"] = T.rel((a_raw.float() * act['inv_raw_mul'][None, :]).t() @ (dt2.float() / lb['dt_mul'] * gmax), dA)
        e["dt2_absmax"] = float(dt2.float().abs().max()); e["dt2_rms"] = float(dt2.float().pow(2).mean().sqrt())
        e["egn_min"] = float(egn.min())
        e["dt_absmax/rms"] = float(dt.abs().max() / dt.pow(2).mean().sqrt())
    print(name[-28:], {k: f"{v:.2e}" for k, v in e.items()})
