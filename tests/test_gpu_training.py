"""GPU tests of the switchable-precision training step (training.SPTrainer / FlatTrainState, spq_adamw_flat,
spq_grad_sumsq) -- BASELINE.json configs[2], p1/train_sp.py:341-397.  pytest -m gpu."""
import random

import numpy as np
import pytest
import torch

from oracle import upstream as up

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_adamw_flat_and_sumsq_match_torch():
    from llm_qat_on_gpt2_b200 import _lib
    torch.manual_seed(0)
    n = 1 << 20
    p0 = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ss = torch.zeros(1, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * (5.0 if step == 2 else 1e-4)       # one step clips, the others do not
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        _lib.grad_sumsq(g, ss, scale=0.5)
        assert abs(float(ss) - float((g.double() * 0.5).pow(2).sum())) <= 1e-5 * float((g.double() * 0.5).pow(2).sum())
        _lib.grad_sumsq(g * 2.0, ss, scale=0.5)                                 # world = 2: sum of two equal shards
        _lib.adamw_flat(p, g * 2.0, m, v, 3e-4, (0.9, 0.999), 1e-8, 0.01, step, grad_scale=0.5, total_sumsq=ss, max_norm=1.0)
        assert rel(p, ref.data) <= 2e-6, (step, rel(p, ref.data))
    st = opt.state[ref]
    assert rel(m, st["exp_avg"]) <= 1e-5 and rel(v, st["exp_avg_sq"]) <= 5e-5     # fma vs addcmul rounding order
    # deterministic reduction
    a, b = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    _lib.grad_sumsq(g, a); _lib.grad_sumsq(g, b)
    assert torch.equal(a, b)


def _tiny_model(bit_widths, n_layer=2, seed=0, lora_b_std=0.02):
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    cfg = up.gpt2_config(n_layer=n_layer, bit_widths=bit_widths, embd_pdrop=0.0)
    cfg.attention_dtype = "fp32"
    torch.manual_seed(seed)
    with up.quiet():
        m = SPLMHeadModel(cfg).cuda()
    with torch.no_grad():
        m.transformer.wte.weight.normal_(0, 0.02)
        m.transformer.wpe.weight.normal_(0, 0.01)
        for n, p in m.named_parameters():
            if n.endswith("lora_B"):
                p.normal_(0, lora_b_std)
    return m, cfg


def _calibrate(model, bits_list, batches):
    key_mods = [m for m in model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
    with up.quiet(), torch.no_grad():
        for bits in bits_list:
            key = f"{bits}bit"
            model.set_precision(bits)
            for m in key_mods:
                qw = m.quantizers_weight[key]
                qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
                lo = m.lora_adapters[key]
                for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                    qq.start_calibration(); qq(w.data); qq.finish_calibration()
            for m in key_mods:
                m.quantizers_input[key].start_calibration()
            model.disable_lora_for_calibration()
            for b in batches:
                model(b)
            model.enable_lora_after_calibration()
            for m in key_mods:
                m.quantizers_input[key].finish_calibration()


def test_lm_head_loss_matches_torch():
    """training.lm_head_loss (LM-head GEMM + softmax-loss kernel + backward GEMM fed by the fp16 gradient operand)
    against F.cross_entropy / kl_div on float64 logits: value and d loss / d hidden."""
    import torch.nn.functional as F
    from llm_qat_on_gpt2_b200.training import lm_head_loss
    model, cfg = _tiny_model((8, 32))
    for p in model.parameters():
        p.requires_grad_(False)
    B, T, C, V = 2, 48, 768, 50257
    torch.manual_seed(1)
    hidden = torch.randn(B, T, C, device="cuda")
    labels = torch.randint(0, V, (B, T), device="cuda")
    W = model.lm_head.weight.double()
    # a teacher well away from the student: d = softmax(s) - softmax(t) is then not a difference of nearly equal
    # numbers, and the comparison measures the kernels rather than the conditioning of the loss
    t_logits = (0.5 * hidden.double() + torch.randn(B, T, C, device="cuda").double()) @ W.t()
    for kind in ("ce", "kl"):
        h = hidden.clone().requires_grad_(True)
        hd = hidden.double().clone().requires_grad_(True)
        logits = hd @ W.t()
        if kind == "ce":
            loss = lm_head_loss(model, h, labels=labels)
            ref = F.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1))
        else:
            loss = lm_head_loss(model, h, teacher_logits=t_logits.float(), temperature=3.0)
            ref = F.kl_div(F.log_softmax(logits[:, :-1] / 3.0, -1).reshape(-1, V), F.log_softmax(t_logits[:, :-1] / 3.0, -1).reshape(-1, V),
                           reduction="batchmean", log_target=True) * 9.0
        (loss * 0.125).backward()
        (ref * 0.125).backward()
        assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref)), (kind, float(loss), float(ref))
        assert rel(h.grad, hd.grad) <= 1e-3, (kind, rel(h.grad, hd.grad))


def test_flat_state_views_and_segments():
    from llm_qat_on_gpt2_b200.training import FlatTrainState
    model, cfg = _tiny_model((4, 8, 32))
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    st = FlatTrainState(model, 32, [4, 8])
    # 5 LayerNorms x (w, b) per width; 8 linears x (A, B) per student width
    assert len(st.slots) == 3 * 10 + 2 * 16
    assert st.segments[32][1] - st.segments[32][0] == 10 * 768
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), before[n])
        if n in st.slots:
            _, off, cnt = st.slots[n]
            assert p.requires_grad and p.data_ptr() == st.flat_param.data_ptr() + 4 * off
            assert p.grad.data_ptr() == st.flat_grad.data_ptr() + 4 * off
        else:
            assert not p.requires_grad
    # autograd accumulates IN the flat buffer
    model.train(); model.set_precision(32)
    ids = torch.randint(0, 50257, (1, 16), device="cuda")
    model(ids, labels=ids)["loss"].backward()
    a, b = st.segments[32]
    assert st.flat_grad[a:b].abs().sum() > 0 and st.flat_grad[b:].abs().sum() == 0
    v0 = st.params[0]._version
    st.bump_versions()
    assert st.params[0]._version == v0 + 1


def test_trainer_graph_replay_matches_eager():
    """The CUDA-graph micro-steps are the eager micro-steps: same losses, same parameters after two optimizer steps
    (up to the fp32 atomics of the weight-gradient GEMM)."""
    from llm_qat_on_gpt2_b200.training import SPTrainer
    g = torch.Generator().manual_seed(0)
    calib = [torch.randint(0, 50257, (2, 64), generator=g).cuda() for _ in range(2)]
    batches = [torch.randint(0, 50257, (2, 64), generator=g).cuda() for _ in range(2)]
    results = []
    for use_graphs in (False, True):
        model, cfg = _tiny_model((4, 8, 32))
        _calibrate(model, (4, 8), calib)
        model.train()
        tr = SPTrainer(model, [4, 8, 32], grad_accum=4, lr=1e-3, rng=random.Random(5), use_graphs=use_graphs,
                       total_lr_steps=40)
        outs = [tr.train_step(b) for b in batches]
        results.append((outs, tr.state.flat_param.clone(), tr))
    (o_e, p_e, _), (o_g, p_g, tg) = results
    assert [o["precisions"] for o in o_e] == [o["precisions"] for o in o_g]
    for a, b in zip(o_e, o_g):
        assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(a["loss"]), (a, b)
    # Adam's first steps move every element by ~lr regardless of the gradient's size, so compare the movement
    assert rel(p_g, p_e) <= 1e-4, rel(p_g, p_e)
    assert len(tg.graphs) >= 2
    from llm_qat_on_gpt2_b200 import _lib
    assert _lib.debug_status() == 0


@pytest.mark.skipif(not up.available(), reason="baseline/_ref not installed")
def test_trainer_accumulated_gradients_vs_upstream_components():
    """Schedule fidelity: one optimizer step (teacher CE fwd+bwd + cache forward, then student micro-steps at
    random widths with LoRA recalibration, KL(T=3) + 1e-7 MSE, / G) against the same step assembled from the
    unmodified upstream components in float32 (oracle/upstream_train.py).  16-bit log students: a quantiser level is
    as fine as the fp16 operand rounding, so code flips cost nothing and the accumulated gradients must agree to
    rel 2e-3 (median 1e-3); precisions drawn and the reported loss must match."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from llm_qat_on_gpt2_b200.training import SPTrainer
    from oracle.upstream_train import UpstreamTrainStep, make_config
    torch.backends.cuda.matmul.allow_tf32 = False
    bw = (12, 16, 32)
    cfg = up.gpt2_config(n_layer=2, bit_widths=bw, embd_pdrop=0.0)
    torch.manual_seed(0)
    with up.quiet():
        ref = up.p1("models_sp").SPLMHeadModel(cfg).cuda()
    with torch.no_grad():
        ref.transformer.wte.weight.normal_(0, 0.02); ref.transformer.wpe.weight.normal_(0, 0.01)
        for n, p in ref.named_parameters():
            if n.endswith("lora_B"):
                p.normal_(0, 0.02)
    g = torch.Generator().manual_seed(1)
    calib = [torch.randint(0, 50257, (2, 64), generator=g).cuda() for _ in range(2)]
    ids = torch.randint(0, 50257, (2, 64), generator=g).cuda()
    CalibrationManager = up.p1_bare("train_sp").CalibrationManager
    with up.quiet():
        mgr = CalibrationManager(ref, [{"input_ids": b} for b in calib], torch.device("cuda"))
        mgr.calibrate_all_precisions([12, 16], num_batches=2)
    cfg2 = up.gpt2_config(n_layer=2, bit_widths=bw, embd_pdrop=0.0)
    cfg2.attention_dtype = "fp32"
    with up.quiet():
        ours = SPLMHeadModel(cfg2).cuda()
        ours.load_state_dict(ref.state_dict(), strict=True)
    for n, p in ref.named_parameters():
        p.requires_grad_("lora_A" in n or "lora_B" in n or ".weights." in n or ".biases." in n)
    tcfg = make_config(grad_accum=4)
    init_params = {n: p.detach().clone() for n, p in ref.named_parameters()}
    random.seed(11)
    total_ref, used_ref, g_ref = UpstreamTrainStep(ref, bw, tcfg, lr=1e-4, total_lr_steps=40).step(ids)
    ours.train()
    rng = random.Random(11)
    tr = SPTrainer(ours, list(bw), grad_accum=4, lr=1e-4, rng=rng, use_graphs=True, total_lr_steps=40)
    # capture the accumulated gradient before the optimizer consumes it
    import llm_qat_on_gpt2_b200._lib as L
    grabbed = {}
    orig = L.grad_sumsq

    def grab(flat, out, scale=1.0):
        grabbed["g"] = flat.clone()
        return orig(flat, out, scale=scale)
    L.grad_sumsq = grab
    try:
        out = tr.train_step(ids)
    finally:
        L.grad_sumsq = orig
    assert out["precisions"] == used_ref
    assert abs(out["loss"] - total_ref) <= 2e-3 * abs(total_ref), (out["loss"], total_ref)
    errs = {}
    for n, (p, off, cnt) in tr.state.slots.items():
        if n in g_ref:
            errs[n] = rel(grabbed["g"][off:off + cnt].view_as(p), g_ref[n])
        else:
            assert float(grabbed["g"][off:off + cnt].abs().max()) == 0.0, n     # untouched width: no gradient
    assert len(errs) == len(g_ref)
    # model-level, through 2 layers + LM head and G micro-steps: the per-GEMM 3-4e-4 adds in quadrature
    # (measured worst 1.07e-3, median 6e-4; upstream under autocast deviates 2.3e-3 at this depth)
    bad = {k: v for k, v in errs.items() if not v <= 2e-3}
    assert not bad, bad
    assert float(np.median(list(errs.values()))) <= 1e-3
    # the parameters moved like upstream's (AdamW, clip 1.0, cosine LR after G micro-steps).  Adam's first step moves
    # every element by ~lr * sign(g), so elements whose gradient is noise-level may go the other way: compare the
    # whole vector, and the direction where the gradient is significant
    refp = dict(ref.named_parameters())
    names = [n for n in tr.state.slots if n in g_ref]
    ours_vec = torch.cat([tr.state.slots[n][0].detach().reshape(-1) for n in names])
    ref_vec = torch.cat([refp[n].detach().reshape(-1) for n in names])
    assert rel(ours_vec, ref_vec) <= 1e-4, rel(ours_vec, ref_vec)
    before = torch.cat([init_params[n].reshape(-1) for n in names])
    gcat = torch.cat([g_ref[n].reshape(-1) for n in names])
    big = gcat.abs() > 1e-2 * gcat.abs().max()
    agree = (torch.sign(ours_vec - before)[big] == torch.sign(ref_vec - before)[big]).float().mean()
    assert float(agree) >= 0.999 and int(big.sum()) > 100, (float(agree), int(big.sum()))
