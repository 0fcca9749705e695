"""GPU parity tests of the CPT (part2) variant against fixtures generated from the unmodified
reference (tests/golden/make_golden_cpt.py).  pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=torch.float32)


def rel_fro(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def calibrate(m, bits, x_batches):
    m.set_precision(bits)
    qw = m.quantizer_weight
    qw.set_num_bits(bits); qw.start_calibration()
    with torch.no_grad():
        qw(m.linear.weight.data)
    qw.finish_calibration()
    qi = m.quantizer_input
    qi.set_num_bits(bits); qi.start_calibration()
    m.calibration_mode = True
    with torch.no_grad():
        for xb in x_batches:
            m(xb)
    m.calibration_mode = False
    qi.finish_calibration()
    lq = m.lora_weight_quantizers[f"{bits}bit"]
    lq.set_num_bits(bits); lq.start_calibration()
    with torch.no_grad():
        lq(m.shared_lora.lora_A); lq(m.shared_lora.lora_B)
    lq.finish_calibration()


def test_cpt_linear_against_golden():
    from llm_qat_on_gpt2_b200.cpt import CPTLinear
    g = np.load(os.path.join(GOLDEN, "cpt_linear.npz"))
    K, N, r = [int(v) for v in g["meta"]]
    m = CPTLinear(K, N, bit_widths=[4, 8, 32], quantizer_per_bit={4: "minmax", 8: "log", 32: None},
                  gradient_bits=8, shared_lora_rank=r, shared_lora_alpha=16).cuda()
    with torch.no_grad():
        m.linear.weight.copy_(dev(g["weight"])); m.linear.bias.copy_(dev(g["bias"]))
        m.shared_lora.lora_A.copy_(dev(g["lora_A"])); m.shared_lora.lora_B.copy_(dev(g["lora_B"]))
    m.train()
    xc = [dev(x) for x in g["x_calib"]]
    x, gy = dev(g["x"]), dev(g["grad_y"])
    for bits in (8, 4):
        calibrate(m, bits, xc)
        exact = bits == 4                                   # min-max parameters are bit-exact, log within 1 ulp
        for nm, q in (("qw", m.quantizer_weight), ("qi", m.quantizer_input), ("lq", m.lora_weight_quantizers[f"{bits}bit"])):
            for suffix, table in (("scale", q.scales), ("zp", q.zero_points)):
                got, want = table[bits].cpu().numpy(), g[f"{nm}{bits}_{suffix}"]
                assert got.shape == want.shape, (nm, bits, suffix)
                if exact:
                    assert np.array_equal(got, want), (nm, bits, suffix)
                else:
                    d = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
                    assert d.max() <= 4, (nm, bits, suffix, d.max())
        m.zero_grad()
        xg = x.clone().requires_grad_(True)
        y = m(xg)
        y.backward(gy)
        assert rel_fro(y.detach().cpu().numpy(), g[f"y{bits}"]) <= TOL
        for got, key in ((xg.grad, "gx"), (m.linear.weight.grad, "gw"), (m.linear.bias.grad, "gb"),
                         (m.shared_lora.lora_A.grad, "gA"), (m.shared_lora.lora_B.grad, "gB")):
            assert rel_fro(got.cpu().numpy(), g[f"{key}{bits}"]) <= TOL, (key, bits, rel_fro(got.cpu().numpy(), g[f"{key}{bits}"]))
    # cycling back needs no recalibration
    m.set_precision(8)
    with torch.no_grad():
        assert rel_fro(m(x).cpu().numpy(), g["y8_again"]) <= TOL
    # gradient quantisers: collect on one backward, quantise the next
    for gq in (m.shared_lora.grad_quantizer_A, m.shared_lora.grad_quantizer_B):
        gq.start_calibration()
    m.zero_grad(); m(x.clone()).backward(gy)
    for gq in (m.shared_lora.grad_quantizer_A, m.shared_lora.grad_quantizer_B):
        gq.finish_calibration()
    assert tuple(m.shared_lora.grad_quantizer_A.scales[8].shape) == g["gqA_scale"].shape
    m.zero_grad(); m(x.clone()).backward(gy)
    # 8-bit fake quantisation of a gradient that itself carries 3e-4 of GEMM rounding: a few codes
    # may differ by one step (1/127 of the channel range), so the bar is 2 steps of 254 on the max
    for got, key in ((m.shared_lora.lora_A.grad, "gA8_gq"), (m.shared_lora.lora_B.grad, "gB8_gq")):
        gn, want = got.cpu().numpy(), g[key]
        assert rel_fro(gn, want) <= 1e-2, (key, rel_fro(gn, want))
    keys = sorted(m.state_dict().keys())
    assert keys == sorted(str(k) for k in g["state_keys"])
    with torch.no_grad():
        m.set_precision(32)
        assert rel_fro(m(x).cpu().numpy(), g["y32"]) <= TOL
    with pytest.raises(ValueError):
        m.set_precision(5)


def test_cpt_quantizer_state_roundtrip_and_eval_passthrough():
    from llm_qat_on_gpt2_b200.cpt import LearnableFakeQuantize
    q = LearnableFakeQuantize(8, channel_dim=-1, quantizer_type="log", is_input=True).cuda()
    x = torch.randn(4, 10, 32, device="cuda")
    q.eval()
    assert q(x) is x                                        # uncalibrated width outside training: pass-through
    q.train()
    with pytest.raises(RuntimeError, match="not calibrated"):
        q(x)
    for bits in (8, 4):
        q.set_num_bits(bits); q.start_calibration(); q(x); q.finish_calibration()
    assert q.calibrated_bits == {4, 8}
    sd = q.state_dict()
    assert {"_scales_4", "_scales_8", "_zero_points_4", "_zero_points_8", "_calibrated_bits", "running_min", "running_max"} <= set(sd)
    q2 = LearnableFakeQuantize(8, channel_dim=-1, quantizer_type="log", is_input=True).cuda()
    q2.load_state_dict(sd)
    assert q2.calibrated_bits == {4, 8} and torch.equal(q2.scales[4], q.scales[4])
    q2.set_num_bits(4); q.set_num_bits(4)
    assert torch.equal(q2(x), q(x))


def test_cpt_model_forward_backward_smoke():
    """Tiny CPTModel: calibrate two widths, cycle precision per step, loss finite, LoRA grads flow
    through the quantised LM head (vocab 211 -> odd leading dimensions on every GEMM path)."""
    from types import SimpleNamespace
    from llm_qat_on_gpt2_b200.cpt import CPTLinear, CPTModel
    mc = SimpleNamespace(vocab_size=211, n_positions=32, n_embd=64, n_layer=2, n_head=4, embd_pdrop=0.0,
                         layer_norm_epsilon=1e-5, bit_widths=[4, 8, 32], quantizer_per_bit={4: "minmax", 8: "log", 32: None},
                         gradient_bits=8, shared_lora_rank=8, shared_lora_alpha=16)
    model = CPTModel({"model": mc, "training": SimpleNamespace(target_bits=8)}).cuda()
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, CPTLinear):
                mod.shared_lora.lora_B.normal_(0, 0.02)
    ids = torch.randint(0, 211, (2, 32), device="cuda")
    lins = [m for m in model.modules() if isinstance(m, CPTLinear)]
    for bits in (8, 4):
        model.set_precision(bits)
        for m in lins:
            m.quantizer_weight.set_num_bits(bits); m.quantizer_weight.start_calibration()
            with torch.no_grad():
                m.quantizer_weight(m.linear.weight.data)
            m.quantizer_weight.finish_calibration()
            m.quantizer_input.set_num_bits(bits); m.quantizer_input.start_calibration()
        model.disable_lora_for_calibration()
        with torch.no_grad():
            model(ids)
        model.enable_lora_after_calibration()
        for m in lins:
            m.quantizer_input.finish_calibration()
            lq = m.lora_weight_quantizers[f"{bits}bit"]
            lq.set_num_bits(bits); lq.start_calibration()
            with torch.no_grad():
                lq(m.shared_lora.lora_A); lq(m.shared_lora.lora_B)
            lq.finish_calibration()
    model.train()
    for p in model.parameters():
        p.requires_grad_(False)
    for m in lins:
        m.shared_lora.lora_A.requires_grad_(True); m.shared_lora.lora_B.requires_grad_(True)
    ref32 = None
    for bits in (8, 4, 8, 32):
        model.set_precision(bits)
        model.zero_grad()
        out = model(ids, labels=ids)
        assert torch.isfinite(out.loss), bits
        if bits < 32:
            out.loss.backward()
            gA = model.lm_head.shared_lora.lora_A.grad
            assert gA is not None and torch.isfinite(gA).all() and gA.abs().sum() > 0
        else:
            ref32 = out.logits
    assert ref32.shape == (2, 32, 211)


def _tiny_cpt(widths=(4, 6, 8), vocab=211, n_embd=64, n_layer=2):
    from types import SimpleNamespace
    from llm_qat_on_gpt2_b200.cpt import CPTModel
    mc = SimpleNamespace(vocab_size=vocab, n_positions=64, n_embd=n_embd, n_layer=n_layer, n_head=4, embd_pdrop=0.0,
                         layer_norm_epsilon=1e-5, bit_widths=list(widths) + [32],
                         quantizer_per_bit={**{b: "log" for b in widths}, 32: None}, gradient_bits=8, shared_lora_rank=8,
                         shared_lora_alpha=16)
    torch.manual_seed(0)
    model = CPTModel({"model": mc, "training": SimpleNamespace(target_bits=widths[-1])}).cuda()
    with torch.no_grad():
        for mod in model.modules():
            if mod.__class__.__name__ == "LoRAAdapter" and mod.lora_B is not None:
                mod.lora_B.normal_(0, 0.02)
    return model


def test_cpt_calibration_manager_and_fused_gradient_quantizer():
    """cpt.CalibrationManager follows upstream's order; with the gradient quantisers calibrated (on real LoRA gradients
    at a quantised width) the GradientQuantizer + log STE clamp fused into the gradient GEMM's fold pass gives the same
    LoRA gradients as the unfused tail (separate fake-quant and clamp launches)."""
    from llm_qat_on_gpt2_b200.cpt import CalibrationManager, cpt_model
    model = _tiny_cpt()
    g = torch.Generator().manual_seed(0)
    loader = [{"input_ids": torch.randint(0, 211, (2, 32), generator=g)} for _ in range(2)]
    mgr = CalibrationManager(model, loader, torch.device("cuda"))
    mgr.calibrate_gradient_quantizers()                      # upstream: at 32 bits -> nothing collected, identity
    gqs = [q for m in model.modules() if m.__class__.__name__ == "LoRAAdapter" for q in (m.grad_quantizer_A, m.grad_quantizer_B)]
    assert all(8 not in q.calibrated_bits for q in gqs)
    for b in (4, 6, 8):
        mgr.ensure_calibrated(b, num_batches=2)
    assert mgr.calibrated_bits == {4, 6, 8} == mgr.lora_calibrated_bits
    mgr.calibrate_gradient_quantizers(precision=8)           # extension: statistics from real LoRA gradients
    assert all(8 in q.calibrated_bits for q in gqs)
    model.train()
    ids = loader[0]["input_ids"].cuda()

    def lora_grads(fused):
        orig = cpt_model._grad_quantizer_scale
        if not fused:
            cpt_model._grad_quantizer_scale = lambda q, rows: (None, q is None or not (q.collecting_stats or q.num_bits in q.calibrated_bits))
        try:
            model.zero_grad(set_to_none=True)
            model.set_precision(6)
            model(ids, labels=ids).loss.backward()
        finally:
            cpt_model._grad_quantizer_scale = orig
        return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None and "shared_lora" in n}
    a, b = lora_grads(True), lora_grads(False)
    assert a.keys() == b.keys() and len(a) == 2 * 9
    for n in a:
        assert torch.equal(a[n], b[n]), n
        # the quantised gradient sits on the 8-bit grid of its row scale
    sl = model.h[0].attn.c_attn.shared_lora
    gA = a["h.0.attn.c_attn.shared_lora.lora_A"]
    sc = sl.grad_quantizer_A.scales[8].reshape(-1, 1)
    codes = gA / sc
    assert torch.allclose(codes, codes.round(), atol=1e-3) and float(codes.abs().max()) <= 127.001


def test_cpt_lm_head_fused_cross_entropy(monkeypatch):
    """CPTModel with labels: the quantised LM head hands its logits to the softmax-loss kernel, whose fp16 gradient
    operand feeds the head's backward (CPTLinear.forward_with_cross_entropy).  Against the composition it replaces
    (F.cross_entropy on the logits, SPQ_CPT_FUSED_CE=0): same loss, same logits, LoRA gradients of every layer and the
    gradient reaching the embeddings within the fp16-operand tolerance; eval mode and the 32-bit width fall through."""
    from llm_qat_on_gpt2_b200.cpt import CalibrationManager
    g = torch.Generator().manual_seed(3)
    loader = [{"input_ids": torch.randint(0, 211, (2, 32), generator=g)} for _ in range(2)]
    ids = loader[0]["input_ids"].cuda()
    labels = ids.clone(); labels[0, 5:9] = -100
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("SPQ_CPT_FUSED_CE", fused)
        model = _tiny_cpt()
        mgr = CalibrationManager(model, loader, torch.device("cuda"))
        for b in (4, 6, 8):
            mgr.ensure_calibrated(b, num_batches=2)
        model.train()
        for p in model.parameters():
            p.requires_grad_(False)
        for m in model.modules():
            if m.__class__.__name__ == "LoRAAdapter" and m.lora_A is not None:
                m.lora_A.requires_grad_(True); m.lora_B.requires_grad_(True)
        model.wte.weight.requires_grad_(True)
        model.set_precision(6)
        out = model(ids, labels=labels)
        out.loss.backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        with torch.no_grad():
            ev = model(ids, labels=labels).loss.item()
            model.set_precision(32)
            l32 = model(ids, labels=labels).loss.item()
        res[fused] = (out.loss.item(), out.logits.detach().float().clone(), grads, ev, l32)
    a, b = res["1"], res["0"]
    assert abs(a[0] - b[0]) <= 1e-5 * abs(b[0]) and abs(a[3] - b[3]) <= 1e-5 * abs(b[3]) and a[4] == b[4]
    assert torch.equal(a[1], b[1]) and a[1].shape == (2, 32, 211)
    assert a[2].keys() == b[2].keys() and len(a[2]) >= 2 * 9 + 1
    for n in a[2]:
        ga, gb = a[2][n].double(), b[2][n].double()
        assert float((ga - gb).norm() / gb.norm().clamp_min(1e-30)) <= 2e-3, n


def test_cpt_trainer_graphs_match_eager_and_cycle():
    """CPTTrainer: per-width CUDA graphs (LoRA-level operand rebuild + forward + loss + backward) reproduce the eager
    steps while the width cycles per step along CyclicPrecisionScheduler; parameters move; losses agree."""
    from llm_qat_on_gpt2_b200.cpt import CalibrationManager, CPTTrainer, CyclicPrecisionScheduler
    g = torch.Generator().manual_seed(1)
    loader = [{"input_ids": torch.randint(0, 211, (2, 32), generator=g)} for _ in range(2)]
    batches = [torch.randint(0, 211, (2, 32), generator=g).cuda() for _ in range(6)]
    sched = CyclicPrecisionScheduler(bit_widths=[4, 6, 8], schedule_type="cosine", total_epochs=12, total_cycles=3)
    seq = [sched.get_precision_for_epoch(i) for i in range(6)]
    assert set(seq) == {4, 6, 8}
    res = []
    for use_graphs in (False, True):
        model = _tiny_cpt()
        mgr = CalibrationManager(model, loader, torch.device("cuda"))
        for b in (4, 6, 8):
            mgr.ensure_calibrated(b, num_batches=2)
        model.train()
        tr = CPTTrainer(model, lr=1e-3, total_lr_steps=100, use_graphs=use_graphs)
        p0 = tr.flat_param.clone()
        losses = [tr.train_step(x, b)["loss"] for x, b in zip(batches, seq)]
        res.append((losses, tr.flat_param.clone(), p0, tr))
    (l_e, p_e, p0, _), (l_g, p_g, _, tg) = res
    assert sorted(tg.graphs) == [4, 6, 8]
    for a, b in zip(l_e, l_g):
        assert abs(a - b) <= 1e-4 * abs(a), (l_e, l_g)
    assert float((p_e - p0).abs().max()) > 1e-4
    assert float((p_g.double() - p_e.double()).norm() / (p_e.double() - p0.double()).norm()) <= 2e-2
