"""GPU parity tests of the drop-in modules (SPLinearWithLoRA, LoRALayer, SPLMHeadModel) against
the reference-generated golden fixtures and the numpy oracle.  pytest -m gpu.

Tolerance for everything that goes through the fp16-operand / fp32-accumulate GEMM: rel-Frobenius
1e-3 against float32 (BASELINE.json); observed ~3e-4.
"""
import os

import numpy as np
import pytest
import torch

from oracle import QuantizerState, sp_linear_backward, sp_linear_forward, collect_statistics, finish_calibration
from oracle.model_oracle import SPModelOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=torch.float32)


def rel_fro(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _calibrate_linear(m, bits, x_batches):
    """Reference calibration order (p1/train_sp.py:47-123): weight -> LoRA -> inputs with LoRA off."""
    key = f"{bits}bit"
    m.set_precision(bits)
    qw = m.quantizers_weight[key]
    qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
    lo = m.lora_adapters[key]
    for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
        qq.start_calibration(); qq(w.data); qq.finish_calibration()
    qi = m.quantizers_input[key]
    qi.start_calibration()
    m.calibration_mode = True
    with torch.no_grad():
        for xb in x_batches:
            m(xb)
    m.calibration_mode = False
    qi.finish_calibration()


def _build_linear(g):
    from llm_qat_on_gpt2_b200 import SPLinearWithLoRA
    K, N, r, bits, qt, pc = [int(v) for v in g["meta"]]
    qtype = "minmax" if qt == 0 else "log"
    m = SPLinearWithLoRA(K, N, bit_widths=[bits, 32], lora_rank_per_bit={bits: r, 32: 0},
                         lora_alpha_per_bit={bits: 2 * r, 32: 0}, quantizer_per_bit={bits: qtype, 32: None},
                         per_channel=bool(pc)).cuda()
    key = f"{bits}bit"
    with torch.no_grad():
        m.linear.weight.copy_(dev(g["weight"])); m.linear.bias.copy_(dev(g["bias"]))
        m.lora_adapters[key].lora_A.copy_(dev(g["lora_A"])); m.lora_adapters[key].lora_B.copy_(dev(g["lora_B"]))
    return m, bits, qtype, key


@pytest.mark.parametrize("name", ["minmax4", "log8", "minmax4_pertensor"])
def test_sp_linear_against_golden(name):
    g = np.load(os.path.join(GOLDEN, f"linear_{name}.npz"))
    m, bits, qtype, key = _build_linear(g)
    _calibrate_linear(m, bits, [dev(xb) for xb in g["x_calib"]])
    # calibrated parameters: bit-exact for min-max; log within 1 ulp of the torch-CPU SLEEF values
    for qn, qq in (("qw", m.quantizers_weight[key]), ("qin", m.quantizers_input[key]),
                   ("qA", m.lora_adapters[key].quantize_A), ("qB", m.lora_adapters[key].quantize_B)):
        for suffix, attr in (("scale", "scale"), ("zp", "zero_point"), ("rmin", "running_min"), ("rmax", "running_max")):
            got, want = getattr(qq, attr).cpu().numpy(), g[f"{qn}_{suffix}"]
            assert got.shape == want.shape, (qn, attr)
            if qtype == "minmax" and qn != "qin":
                assert np.array_equal(got, want), (qn, attr)
            elif qtype == "minmax":
                # the input statistics are taken on this repo's GEMM-free path (x itself) -> exact too
                assert np.array_equal(got, want), (qn, attr)
            else:
                d = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
                assert d.max() <= 4, (qn, attr, d.max())
    for p in m.parameters():
        p.requires_grad_(True)
    x = dev(g["x"]).requires_grad_(True)
    y = m(x)
    assert rel_fro(y.detach().cpu().numpy(), g["y"]) <= TOL
    y.backward(dev(g["grad_y"]))
    lo = m.lora_adapters[key]
    for got, want in ((x.grad, "grad_x"), (m.linear.weight.grad, "grad_weight"), (m.linear.bias.grad, "grad_bias"),
                      (lo.lora_A.grad, "grad_lora_A"), (lo.lora_B.grad, "grad_lora_B")):
        assert got is not None, want
        assert rel_fro(got.cpu().numpy(), g[want]) <= TOL, (want, rel_fro(got.cpu().numpy(), g[want]))
    with torch.no_grad():
        m.calibration_mode = True
        assert rel_fro(m(x).cpu().numpy(), g["y_base"]) <= TOL
        m.calibration_mode = False
        m.set_precision(32)
        assert rel_fro(m(x).cpu().numpy(), g["y32"]) <= TOL
    # 32-bit path gradients (teacher): plain linear
    m.zero_grad(); x.grad = None
    m(x).backward(dev(g["grad_y"]))
    gy = g["grad_y"].reshape(-1, g["grad_y"].shape[-1]); x2 = g["x"].reshape(-1, g["x"].shape[-1])
    assert rel_fro(x.grad.cpu().numpy().reshape(x2.shape), gy @ g["weight"]) <= TOL
    assert rel_fro(m.linear.weight.grad.cpu().numpy(), gy.T @ x2) <= TOL
    assert rel_fro(m.linear.bias.grad.cpu().numpy(), gy.sum(0)) <= 1e-5


@pytest.mark.parametrize("K,N,r,bits,qtype,M", [(768, 2304, 64, 4, "minmax", 1024), (3072, 768, 64, 8, "log", 515),
                                                (1600, 4800, 64, 8, "log", 640), (768, 768, 16, 6, "log", 300),
                                                (1024, 1024, 64, 3, "minmax", 256)])
def test_sp_linear_against_oracle_gpt2_shapes(K, N, r, bits, qtype, M):
    """GPT-2 small / XL layer shapes, forward + STE backward, against the numpy oracle fed with the
    GPU-calibrated parameters (so both sides quantise with identical scales)."""
    from llm_qat_on_gpt2_b200 import SPLinearWithLoRA
    torch.manual_seed(K + N + bits)
    m = SPLinearWithLoRA(K, N, bit_widths=[bits, 32], lora_rank_per_bit={bits: r, 32: 0},
                         lora_alpha_per_bit={bits: r, 32: 0}, quantizer_per_bit={bits: qtype, 32: None}).cuda()
    key = f"{bits}bit"
    lo = m.lora_adapters[key]
    with torch.no_grad():
        m.linear.weight.normal_(0, 0.02); m.linear.bias.normal_(0, 0.02); lo.lora_B.normal_(0, 0.02)
    gen = torch.Generator(device="cuda").manual_seed(3)
    def batch():
        x = torch.randn(M, K, device="cuda", generator=gen)
        x[:, 5] *= 20.0; x[:, 11] *= 20.0           # outlier channels
        return x.view(1, M, K)
    _calibrate_linear(m, bits, [batch(), batch()])
    for p in (lo.lora_A, lo.lora_B, m.linear.weight, m.linear.bias):
        p.requires_grad_(True)
    x = batch().requires_grad_(True)
    gy = torch.randn(1, M, N, device="cuda", generator=gen) * 0.01
    y = m(x); y.backward(gy)

    def st(q, cd, is_input=False):
        s = QuantizerState(bits, channel_dim=cd, quantizer_type=qtype, is_input=is_input)
        s.scale, s.zero_point = q.scale.cpu().numpy(), q.zero_point.cpu().numpy()
        s.calibrated = True
        return s
    lora = {"A": lo.lora_A.detach().cpu().numpy(), "B": lo.lora_B.detach().cpu().numpy(),
            "q_A": st(lo.quantize_A, 1), "q_B": st(lo.quantize_B, 1), "scaling": lo.scaling}
    qin, qw = st(m.quantizers_input[key], -1, True), st(m.quantizers_weight[key], 0)
    W, b = m.linear.weight.detach().cpu().numpy(), m.linear.bias.detach().cpu().numpy()
    xn, gn = x.detach().cpu().numpy(), gy.cpu().numpy()
    y_ref = sp_linear_forward(xn, W, b, bits, qin, qw, lora)
    gr = sp_linear_backward(gn, xn, W, bits, qin, qw, lora)
    assert rel_fro(y.detach().cpu().numpy(), y_ref) <= TOL
    assert rel_fro(x.grad.cpu().numpy(), gr["x"]) <= TOL
    assert rel_fro(lo.lora_A.grad.cpu().numpy(), gr["lora_A"]) <= TOL
    assert rel_fro(lo.lora_B.grad.cpu().numpy(), gr["lora_B"]) <= TOL
    assert rel_fro(m.linear.weight.grad.cpu().numpy(), gr["weight"]) <= TOL
    assert rel_fro(m.linear.bias.grad.cpu().numpy(), gr["bias"]) <= 1e-5
    from llm_qat_on_gpt2_b200 import _lib
    assert _lib.debug_status() == 0


def test_lora_layer_standalone_and_errors():
    from llm_qat_on_gpt2_b200 import LoRALayer, SPLinearWithLoRA
    lo = LoRALayer(64, 96, 8, 16, 4, "minmax").cuda()
    with torch.no_grad():
        lo.lora_B.normal_(0, 0.05)
    x = torch.randn(2, 10, 64, device="cuda")
    with pytest.raises(RuntimeError):
        lo(x)                                         # A/B quantisers uncalibrated
    for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
        qq.start_calibration(); qq(w.data); qq.finish_calibration()
    y = lo(x)
    a_q, b_q = lo.quantize_A(lo.lora_A), lo.quantize_B(lo.lora_B)
    ref = (x @ a_q @ b_q) * lo.scaling
    assert rel_fro(y.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TOL
    off = LoRALayer(64, 96, 0, 0, 32, None).cuda()
    assert not off.enabled and torch.count_nonzero(off(x)) == 0 and off(x).shape == (2, 10, 96)
    m = SPLinearWithLoRA(64, 96, bit_widths=[4, 32], lora_rank_per_bit={4: 8, 32: 0}, lora_alpha_per_bit={4: 16, 32: 0},
                         quantizer_per_bit={4: "minmax", 32: None}).cuda()
    assert m.current_bits == 4
    with pytest.raises(RuntimeError):
        m(x)                                          # uncalibrated, as the reference
    m.current_bits = 6
    with pytest.raises(KeyError):
        m(x)
    keys = set(m.state_dict().keys())
    for k in ("weight_quantized", "linear.weight", "linear.bias", "quantizers_weight.4bit.scale",
              "quantizers_input.4bit.running_max", "lora_adapters.4bit.lora_A", "lora_adapters.4bit.lora_B_quantized",
              "lora_adapters.4bit.quantize_A.zero_point"):
        assert k in keys, k


def _tiny_config():
    from transformers import GPT2Config
    cfg = GPT2Config(vocab_size=211, n_positions=32, n_embd=64, n_layer=2, n_head=4, layer_norm_epsilon=1e-5, embd_pdrop=0.0)
    cfg.bit_widths = [4, 8, 32]
    cfg.lora_rank_per_bit = {4: 8, 8: 8, 32: 0}
    cfg.lora_alpha_per_bit = {4: 16, 8: 16, 32: 0}
    cfg.quantizer_per_bit = {4: "minmax", 8: "log", 32: None}
    cfg.per_channel_quantization = True
    return cfg


def _calibrate_model(model, bits, calib):
    key = f"{bits}bit"
    model.set_precision(bits)
    mods = [m for m in model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
    with torch.no_grad():
        for m in mods:
            qw = m.quantizers_weight[key]
            qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
            lo = m.lora_adapters[key]
            for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                qq.start_calibration(); qq(w.data); qq.finish_calibration()
        for m in mods:
            m.quantizers_input[key].start_calibration()
        model.disable_lora_for_calibration()
        for c in calib:
            model(c)
        model.enable_lora_after_calibration()
        for m in mods:
            m.quantizers_input[key].finish_calibration()


def test_tiny_model_against_golden_and_oracle():
    """Whole-model check.  The state_dict of the reference loads with strict=True; logits at 32 bits
    agree to 1e-3.  At 4/8 bits quantisation is discontinuous, so a rounding difference upstream
    can flip a code downstream: the whole-model bar is statistical (2 %), the exact bars are the
    per-kernel and per-layer tests above."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    g = np.load(os.path.join(GOLDEN, "tiny_model.npz"))
    model = SPLMHeadModel(_tiny_config()).cuda().eval()
    sd = model.state_dict()
    loaded = 0
    with torch.no_grad():
        for k in g.files:
            if k.startswith("sd::"):
                name = k[4:]
                assert name in sd, name
                assert tuple(sd[name].shape) == g[k].shape, name
                sd[name].copy_(dev(g[k])); loaded += 1
    assert loaded > 50
    ids = torch.as_tensor(g["ids"]).cuda()
    calib = [torch.as_tensor(c).cuda() for c in g["calib_ids"]]
    with torch.no_grad():
        model.set_precision(32)
        assert rel_fro(model(ids).cpu().numpy(), g["logits32"]) <= TOL
    for b in (4, 8):
        _calibrate_model(model, b, calib)
        ok, details = model.verify_precision_consistency()
        assert ok, details
        with torch.no_grad():
            out = model(ids, output_hidden_states=True)
        # Each layer is within 1e-3 of float32 on identical inputs (tests above), but that 3e-4 of fp16
        # operand rounding is enough to flip a handful of 4-bit codes in the NEXT layer of this 64-wide toy
        # model, and one flip moves an activation by 1/7 of its range (two float32 implementations -- the
        # oracle and the reference -- agree to 3e-7 here because they flip nothing).  So the whole-model
        # bar is statistical: relative error of the first block's output and direction of the logits.
        h1, want_h1 = out["hidden_states"][1].cpu().numpy(), g[f"hidden{b}_1"]
        lg, want_lg = out["logits"].cpu().numpy().astype(np.float64), g[f"logits{b}"].astype(np.float64)
        cos = float((lg * want_lg).sum() / np.linalg.norm(lg) / np.linalg.norm(want_lg))
        assert rel_fro(h1, want_h1) <= 6e-2, (b, rel_fro(h1, want_h1))
        assert cos >= 0.98, (b, cos)
    with pytest.raises(ValueError):
        model.set_precision(5)
    # training step smoke: CE loss, LoRA + LN gradients flow, base weights frozen
    model.train()
    for n, p in model.named_parameters():
        p.requires_grad_(("lora_" in n) or ("weights." in n) or ("biases." in n))
    model.set_precision(4)
    out = model(ids, labels=ids)
    out["loss"].backward()
    assert torch.isfinite(out["loss"])
    got = [n for n, p in model.named_parameters() if p.grad is not None]
    assert any("lora_adapters.4bit.lora_A" in n for n in got) and any("ln_1.weights.4" in n for n in got)
    assert not any("lora_adapters.8bit" in n for n in got)
    assert model.transformer.h[0].attn.c_attn.linear.weight.grad is None


def test_eval_fast_paths_match_autograd_paths():
    """no_grad fast paths (GELU in the GEMM epilogue, fused cross-entropy on stride-padded logits)
    give the same numbers as the differentiable composition."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    torch.manual_seed(0)
    model = SPLMHeadModel(_tiny_config()).cuda().eval()
    ids = torch.randint(0, 211, (2, 32), device="cuda")
    _calibrate_model(model, 8, [ids])
    for bits in (8, 32):
        model.set_precision(bits)
        with torch.no_grad():
            fast = model(ids, labels=ids)
        for p in model.parameters():
            p.requires_grad_(False)
        model.transformer.h[0].ln_1.weights[str(bits)].requires_grad_(True)
        slow = model(ids, labels=ids)
        assert slow["logits"].requires_grad
        assert tuple(fast["logits"].shape) == (2, 32, 211)
        # the epilogue GELU uses a 1.5e-7-accurate erf: a handful of 8-bit codes flip downstream
        assert rel_fro(fast["logits"].cpu().numpy(), slow["logits"].detach().cpu().numpy()) <= TOL
        assert abs(fast["loss"].item() - slow["loss"].item()) <= TOL * abs(slow["loss"].item())
        slow["loss"].backward()                       # LM-head backward with the odd vocab width
        assert torch.isfinite(model.transformer.h[0].ln_1.weights[str(bits)].grad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("qtype", ["minmax", "log"])
def test_sp_linear_half_input_equals_float_input(qtype):
    """A float16 layer input (fp16 attention output -> c_proj) gives the bits of the same values in float32,
    through calibration and through the quantised forward."""
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    torch.manual_seed(3)
    dev = torch.device("cuda")
    outs = []
    xh = (torch.randn(4, 96, 256, device=dev) * 2).half()
    for x in (xh.float(), xh):
        torch.manual_seed(5)
        m = SPLinearWithLoRA(256, 384, [4, 8, 32], {4: 8, 8: 8, 32: 0}, {4: 16, 8: 16, 32: 0},
                             {4: qtype, 8: qtype, 32: None}).to(dev)
        with torch.no_grad():
            m.lora_adapters['8bit'].lora_B.normal_(0, 0.02)
        m.set_precision(8)
        with torch.no_grad():
            m.quantizers_weight['8bit'].start_calibration(); m.quantizers_weight['8bit'](m.linear.weight)
            m.quantizers_weight['8bit'].finish_calibration()
            m.calibration_mode = True
            m.quantizers_input['8bit'].start_calibration()
            y_cal = m(x)
            m.quantizers_input['8bit'].finish_calibration()
            m.calibration_mode = False
            for q in (m.lora_adapters['8bit'].quantize_A, m.lora_adapters['8bit'].quantize_B):
                q.start_calibration()
            q = m.lora_adapters['8bit']
            q.quantize_A(q.lora_A); q.quantize_B(q.lora_B)
            q.quantize_A.finish_calibration(); q.quantize_B.finish_calibration()
            y = m(x)
            y32 = None
            m.set_precision(32)
            y32 = m(x)
        outs.append((y_cal.float(), y.float(), y32.float(), m.quantizers_input['8bit'].scale.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_sp_linear_residual_epilogue_is_the_separate_add():
    """forward(x, residual=r) under no_grad (residual added in the GEMM epilogue) == r + forward(x), bit for bit,
    at a quantised precision with LoRA, at 32 bits, and during the calibration pass; with autograd on it is the
    plain add and gradients reach the residual."""
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    torch.manual_seed(11)
    dev = torch.device("cuda")
    m = SPLinearWithLoRA(256, 384, [4, 8, 32], {4: 8, 8: 8, 32: 0}, {4: 16, 8: 16, 32: 0},
                         {4: "minmax", 8: "log", 32: None}).to(dev)
    x = torch.randn(3, 50, 256, device=dev)
    r = torch.randn(3, 50, 384, device=dev)
    with torch.no_grad():
        for bits in (4, 8):
            m.lora_adapters[f'{bits}bit'].lora_B.normal_(0, 0.02)
            m.set_precision(bits)
            wq = m.quantizers_weight[f'{bits}bit']
            wq.start_calibration(); wq(m.linear.weight); wq.finish_calibration()
            m.calibration_mode = True
            iq = m.quantizers_input[f'{bits}bit']
            iq.start_calibration()
            y_cal = m(x, residual=r)
            iq2_min = iq.temp_min.clone()
            assert torch.equal(y_cal, r + m(x))          # second collection of the same batch: statistics unchanged
            assert torch.equal(iq2_min, iq.temp_min)
            iq.finish_calibration()
            m.calibration_mode = False
            lo = m.lora_adapters[f'{bits}bit']
            for q, t in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                q.start_calibration(); q(t); q.finish_calibration()
            assert torch.equal(m(x, residual=r), r + m(x))
        m.set_precision(32)
        assert torch.equal(m(x, residual=r), r + m(x))
    m.set_precision(8)
    r2 = r.clone().requires_grad_(True)
    y = m(x, residual=r2)
    y.sum().backward()
    assert torch.equal(r2.grad, torch.ones_like(r2))
    with torch.no_grad():
        # the no-grad forward stores the LoRA intermediate as fp16 straight from the down-projection epilogue;
        # its scales are powers of two, so it is the value the training path rounds in two steps
        assert torch.equal(y, r + m(x))


@pytest.mark.gpu
def test_calibrate_many_equals_per_quantiser_calibration():
    """One launch for many small tensors == start_calibration / forward / finish_calibration per quantiser, bit for
    bit: per-column, per-row and per-tensor layouts, min-max (symmetric and not) and log, NaN columns, and the
    log quantiser whose tensor is all zeros (the reference's default-shape case, taken through the slow path)."""
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, calibrate_many
    torch.manual_seed(21)
    dev = torch.device("cuda")
    specs = []
    for qtype in ("minmax", "log"):
        for sym in (True, False):
            specs += [dict(bits=8, cd=1, pc=True, qtype=qtype, sym=sym, shape=(768, 64)),
                      dict(bits=4, cd=0, pc=True, qtype=qtype, sym=sym, shape=(64, 2304)),
                      dict(bits=8, cd=1, pc=True, qtype=qtype, sym=sym, shape=(64, 2304)),
                      dict(bits=6, cd=0, pc=False, qtype=qtype, sym=sym, shape=(333, 77)),
                      dict(bits=8, cd=-1, pc=True, qtype=qtype, sym=sym, shape=(3072, 16))]
    tensors = [torch.randn(*sp["shape"], device=dev) * (0.02 + 0.3 * i) for i, sp in enumerate(specs)]
    tensors[2][5, 7] = float("nan")
    tensors[1][:, 3] = 0.0
    specs.append(dict(bits=8, cd=1, pc=True, qtype="log", sym=True, shape=(64, 2304)))
    tensors.append(torch.zeros(64, 2304, device=dev))           # fresh lora_B: nothing above eps
    def make(sp):
        return LearnableFakeQuantize(sp["bits"], channel_dim=sp["cd"], quantizer_type=sp["qtype"], symmetric=sp["sym"],
                                     per_channel=sp["pc"]).to(dev)
    ref = [make(sp) for sp in specs]
    for q, w in zip(ref, tensors):
        q.start_calibration(); q(w); q.finish_calibration()
    got = [make(sp) for sp in specs]
    for rep in range(2):                                          # second call: cached job table, new values
        if rep == 1:
            for w in tensors[:-1]:
                w.mul_(1.7)
            for q, w in zip(ref, tensors):
                q.start_calibration(); q(w); q.finish_calibration()
        calibrate_many(got, tensors)
        for a, b, sp in zip(ref, got, specs):
            assert b.calibrated and not b.collecting_stats
            for name in ("running_min", "running_max", "scale", "zero_point"):
                x, y = getattr(a, name), getattr(b, name)
                assert x.shape == y.shape, (sp, name, x.shape, y.shape)
                assert torch.equal(x.view(torch.int32), y.view(torch.int32)), (sp, name)
    assert torch.equal(ref[0](tensors[0]), got[0](tensors[0]))          # and the quantisers quantise alike


@pytest.mark.gpu
@pytest.mark.parametrize("qtype", ["log", "minmax"])
def test_lora_refresher_graph_equals_eager(qtype):
    """training.LoRARefresher (CUDA-graph replay of LoRA recalibration + operand rebuild) leaves every linear in
    the state of the eager path: forward outputs and all gradients identical over several optimizer-like updates,
    including the first step with an all-zero lora_B (log quantiser without data -> host fallback)."""
    import copy
    from llm_qat_on_gpt2_b200 import calibrate_many
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    from llm_qat_on_gpt2_b200.training import LoRARefresher
    torch.manual_seed(31)
    dev = torch.device("cuda")
    shapes = [(256, 384), (384, 256), (256, 1024)]
    eager = [SPLinearWithLoRA(i, o, [8, 32], {8: 16, 32: 0}, {8: 32, 32: 0}, {8: qtype, 32: None}).to(dev) for i, o in shapes]
    xs = [torch.randn(4, 33, i, device=dev) for i, _ in shapes]
    for m, x in zip(eager, xs):
        m.set_precision(8)
        with torch.no_grad():
            wq = m.quantizers_weight['8bit']; wq.start_calibration(); wq(m.linear.weight); wq.finish_calibration()
            m.calibration_mode = True
            iq = m.quantizers_input['8bit']; iq.start_calibration(); m(x); iq.finish_calibration()
            m.calibration_mode = False
    graphed = copy.deepcopy(eager)
    ref = LoRARefresher(graphed, 8)
    for step in range(4):
        if step > 0:
            with torch.no_grad():
                for a, b in zip(eager, graphed):
                    for name in ("lora_A", "lora_B"):
                        d = torch.randn_like(getattr(a.lora_adapters['8bit'], name)) * 0.01
                        getattr(a.lora_adapters['8bit'], name).add_(d)
                        getattr(b.lora_adapters['8bit'], name).add_(d)
        los = [m.lora_adapters['8bit'] for m in eager]
        with torch.no_grad():
            calibrate_many([q for lo in los for q in (lo.quantize_A, lo.quantize_B)],
                           [w.data for lo in los for w in (lo.lora_A, lo.lora_B)])
        ref.refresh()
        for a, b, x in zip(eager, graphed, xs):
            for m in (a, b):
                for p in m.parameters():
                    p.grad = None
            xa = x.clone().requires_grad_(True); xb = x.clone().requires_grad_(True)
            ya = a(xa); yb = b(xb)
            assert torch.equal(ya, yb), step
            g = torch.randn_like(ya)
            ya.backward(g); yb.backward(g)
            assert torch.equal(xa.grad, xb.grad)
            la, lb = a.lora_adapters['8bit'], b.lora_adapters['8bit']
            # dA / dB come from the split-K kernel (fp32 atomics: summation order varies run to run)
            for ga, gb in ((la.lora_A.grad, lb.lora_A.grad), (la.lora_B.grad, lb.lora_B.grad)):
                assert float((ga - gb).abs().max()) <= 1e-5 * max(float(ga.abs().max()), 1e-30)
            for qa, qb in ((la.quantize_A, lb.quantize_A), (la.quantize_B, lb.quantize_B)):
                assert torch.equal(qa.scale, qb.scale) and torch.equal(qa.zero_point, qb.zero_point)
    assert ref.graph is not None


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,padded", [(3, 17, 211, True), (2, 64, 50257, True), (4, 9, 1000, False)])
def test_distillation_kl_loss_matches_torch(B, T, V, padded):
    """training.distillation_kl_loss (one kernel: value + gradient) == the reference's log_softmax / kl_div
    composition (p1/distillation_manager.py:64-80) evaluated in float64."""
    import torch.nn.functional as F
    from llm_qat_on_gpt2_b200.training import distillation_kl_loss
    torch.manual_seed(V + T)
    dev = torch.device("cuda")
    Tmp = 3.0
    ld = (V + 31) // 32 * 32 if padded else V
    sbuf = torch.randn(B, T, ld, device=dev) * 4; tbuf = torch.randn(B, T, ld, device=dev) * 4
    s = sbuf[..., :V].detach().requires_grad_(True)
    t = tbuf[..., :V]
    loss = distillation_kl_loss(s, t, Tmp)
    loss.backward()
    s64 = sbuf[..., :V].double().detach().requires_grad_(True)
    ref = F.kl_div(F.log_softmax(s64[:, :-1] / Tmp, dim=-1), F.log_softmax(t.double()[:, :-1] / Tmp, dim=-1),
                   reduction="sum", log_target=True) * (Tmp * Tmp / (B * (T - 1)))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    gerr = (s.grad.double() - s64.grad).abs().max() / s64.grad.abs().max()
    assert float(gerr) <= 1e-4, float(gerr)
    assert float(s.grad[:, -1].abs().max()) == 0.0                  # the last position takes no part


@pytest.mark.gpu
def test_distillation_loss_against_reference_fixture():
    """training.distillation_loss vs DistillationManager.compute_distillation_loss of the unmodified reference
    (tests/golden/make_golden_distill.py): loss within 1e-5 relative, gradient w.r.t. the student logits within
    1e-4 of its largest entry (float32 softmax on both sides)."""
    import os
    import numpy as np
    from llm_qat_on_gpt2_b200.training import distillation_loss
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "distill_kl.npz"))
    dev = torch.device("cuda")
    s = torch.tensor(g["s_logits"], device=dev).requires_grad_(True)
    t = torch.tensor(g["t_logits"], device=dev)
    hs = [torch.tensor(h, device=dev) for h in g["s_hidden"]]
    ht = [torch.tensor(h, device=dev) for h in g["t_hidden"]]
    loss = distillation_loss({'logits': s, 'hidden_states': hs}, {'logits': t, 'hidden_states': ht},
                             float(g["temperature"]), float(g["alpha_kl"]), float(g["alpha_feature"]), accumulative=True)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    ref_grad = torch.tensor(g["grad"], device=dev)
    assert float((s.grad - ref_grad).abs().max()) <= 1e-4 * float(ref_grad.abs().max())


@pytest.mark.gpu
def test_export_integer_weights_roundtrip():
    """deploy.export_integer_weights: codes and scales of the calibrated weight quantisers reproduce the
    fake-quantised weights exactly (min-max), int8 storage when the codes fit."""
    from llm_qat_on_gpt2_b200.deploy import export_integer_weights
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    torch.manual_seed(5)
    dev = torch.device("cuda")
    net = torch.nn.ModuleDict({"a": SPLinearWithLoRA(64, 96, [4, 8, 32], {4: 4, 8: 4, 32: 0}, {4: 8, 8: 8, 32: 0},
                                                     {4: "minmax", 8: "minmax", 32: None}),
                               "b": SPLinearWithLoRA(96, 64, [4, 8, 32], {4: 4, 8: 4, 32: 0}, {4: 8, 8: 8, 32: 0},
                                                     {4: "minmax", 8: "minmax", 32: None})}).to(dev)
    for m in net.values():
        m.set_precision(8)
        q = m.quantizers_weight["8bit"]
        q.start_calibration(); q(m.linear.weight.data); q.finish_calibration()
    out = export_integer_weights(net, 8)
    for name, m in net.items():
        codes, scale = out[f"{name}.codes"], out[f"{name}.scale"]
        assert codes.dtype == torch.int8 and codes.shape == m.linear.weight.shape
        dq = m.quantizers_weight["8bit"](m.linear.weight.data).cpu()
        assert torch.equal(codes.float() * scale, dq)
    with pytest.raises(RuntimeError):
        export_integer_weights(net, 4)                      # 4-bit quantisers were never calibrated


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [4, 8])
def test_gpt2_small_against_oracle(bits):
    """BASELINE.json configs[0] at the real model size (GPT-2 small, 124M; 4-bit min-max per-channel, and the 8-bit
    log configuration of the headline metric): calibrate + forward of the CUDA path against the numpy oracle on
    the same weights and tokens.  Calibrated first-layer statistics exact; 32-bit logits within 1e-3; at 4/8 bits
    the per-layer outputs are within 1e-3 but code flips compound over 12 layers, so the whole-model bar is the
    statistical one of the tiny-model test (first block, logit direction, loss)."""
    from transformers import GPT2Config
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from oracle.model_oracle import SPModelOracle, cross_entropy_shifted, random_state_dict
    V, P, B, T = 50257, 1024, 2, 64
    widths = [4, 8, 32]
    qt = {4: "minmax", 8: "log", 32: None}
    ocfg = dict(n_layer=12, n_head=12, n_embd=768, layer_norm_epsilon=1e-5, bit_widths=widths, quantizer_per_bit=qt,
                lora_rank_per_bit={4: 16, 8: 16, 32: 0}, lora_alpha_per_bit={4: 32, 8: 32, 32: 0}, per_channel=True)
    sd = random_state_dict(ocfg, V, P, seed=3)
    oracle = SPModelOracle(ocfg, sd)
    cfg = GPT2Config(vocab_size=V, n_positions=P, n_embd=768, n_layer=12, n_head=12, layer_norm_epsilon=1e-5, embd_pdrop=0.0)
    cfg.bit_widths = widths
    cfg.lora_rank_per_bit = ocfg["lora_rank_per_bit"]
    cfg.lora_alpha_per_bit = ocfg["lora_alpha_per_bit"]
    cfg.quantizer_per_bit = qt
    cfg.per_channel_quantization = True
    model = SPLMHeadModel(cfg).cuda().eval()
    msd = model.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            assert k in msd and tuple(msd[k].shape) == v.shape, k
            msd[k].copy_(dev(v))
    rng = np.random.default_rng(11)
    ids_np = rng.integers(0, V, (B, T))
    ids = torch.as_tensor(ids_np).cuda()
    with torch.no_grad():
        model.set_precision(32)
        l32 = model(ids).cpu().numpy()
    oracle.set_precision(32)
    assert rel_fro(l32, oracle.forward(ids_np)) <= TOL
    _calibrate_model(model, bits, [ids])
    oracle.set_precision(bits)
    oracle.calibrate_weights(bits); oracle.calibrate_lora(bits); oracle.calibrate_inputs(bits, [ids_np])
    grabbed = {}
    hook = model.transformer.h[0].attn.c_attn.register_forward_hook(lambda m, i, o: grabbed.setdefault("qkv", o.detach()))
    with torch.no_grad():
        out = model(ids, labels=ids, output_hidden_states=True)
    hook.remove()
    want_logits, want_hidden = oracle.forward(ids_np, return_hidden=True)
    # the first quantised linear sees identical inputs on both sides: the per-layer bar (1e-3) applies
    want_qkv = oracle.linears[0]["c_attn"].forward(oracle._ln(want_hidden[0], "transformer.h.0.ln_1"))
    assert rel_fro(grabbed["qkv"].float().cpu().numpy(), want_qkv) <= TOL, rel_fro(grabbed["qkv"].float().cpu().numpy(), want_qkv)
    lg = out["logits"].cpu().numpy().astype(np.float64); wl = want_logits.astype(np.float64)
    cos = float((lg * wl).sum() / np.linalg.norm(lg) / np.linalg.norm(wl))
    h1 = out["hidden_states"][1].cpu().numpy()
    print(f"gpt2-small {bits}-bit: block-0 output rel {rel_fro(h1, want_hidden[1]):.3e}, logits cosine {cos:.5f}")
    # 4-bit: one flipped code moves an activation by 1/7 of its channel range, and the first block already
    # contains four quantised linears fed by each other
    assert rel_fro(h1, want_hidden[1]) <= (0.15 if bits == 4 else 6e-2), rel_fro(h1, want_hidden[1])
    want_loss = cross_entropy_shifted(want_logits, ids_np)
    print(f"gpt2-small {bits}-bit: loss {out['loss'].item():.5f} vs oracle {want_loss:.5f}")
    assert cos >= (0.95 if bits == 4 else 0.98), cos
    assert abs(out["loss"].item() - want_loss) <= 2e-2 * want_loss, (out["loss"].item(), want_loss)


@pytest.mark.gpu
def test_graphed_no_grad_forward_equals_eager():
    """training.GraphedNoGradForward: the CUDA-graph replay of a 32-bit (teacher) forward returns the eager result
    bit for bit, for new token ids on every call."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from llm_qat_on_gpt2_b200.training import GraphedNoGradForward
    torch.manual_seed(2)
    model = SPLMHeadModel(_tiny_config()).cuda().eval()
    model.set_precision(32)
    teacher = GraphedNoGradForward(model, output_hidden_states=True, return_dict=True)
    for step in range(3):
        ids = torch.randint(0, 211, (2, 32), device="cuda")
        got = teacher(ids)
        with torch.no_grad():
            want = model(ids, output_hidden_states=True, return_dict=True)
        assert torch.equal(got["logits"], want["logits"]), step
        assert all(torch.equal(a, b) for a, b in zip(got["hidden_states"], want["hidden_states"]))
    assert teacher.graph is not None


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [8, 4])
def test_graphed_calibrated_forward_equals_eager(bits):
    """training.GraphedCalibratedForward (the headline step of bench.py: calibration pass of the input quantisers on the
    batch + quantised forward + CE as two CUDA-graph replays) against the eager module calls of upstream's
    CalibrationManager sequence (p1/train_sp.py:47-83) on the same ids: calibrated statistics, scale / zero-point,
    logits and loss bit for bit, for new token ids on every call; eager calls may be interleaved."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from llm_qat_on_gpt2_b200.training import GraphedCalibratedForward
    torch.manual_seed(5)
    model = SPLMHeadModel(_tiny_config()).cuda().eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("lora_B"):
                p.normal_(0, 0.02)
    _calibrate_model(model, bits, [torch.randint(0, 211, (2, 32), device="cuda")])
    key = f"{bits}bit"
    mods = [m for m in model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
    qs = [m.quantizers_input[key] for m in mods]
    step = GraphedCalibratedForward(model)

    def eager(ids):
        with torch.no_grad():
            for q in qs:
                q.start_calibration()
            model.disable_lora_for_calibration()
            model.transformer(ids)
            model.enable_lora_after_calibration()
            for q in qs:
                q.finish_calibration()
            out = model(ids, labels=ids)
        return out, [(q.running_min.clone(), q.running_max.clone(), q.scale.clone(), q.zero_point.clone()) for q in qs]

    for it in range(4):
        ids = torch.randint(0, 211, (2, 32), device="cuda")
        got = step(ids)
        got_logits, got_loss = got["logits"].clone(), got["loss"].clone()
        got_stats = [(q.running_min.clone(), q.running_max.clone(), q.scale.clone(), q.zero_point.clone()) for q in qs]
        if it % 2 == 0:                     # an eager call in between must not disturb the next replay
            want, want_stats = eager(ids)
            for a, b in zip(got_stats, want_stats):
                assert all(torch.equal(x, y) for x, y in zip(a, b)), it
            assert torch.equal(got_logits, want["logits"]), it
            assert torch.equal(got_loss.reshape(()), want["loss"].reshape(())), it
    assert step.g1 is not None and step.kernels_per_replay > 0
    assert step.nodata_count() == 0


@pytest.mark.gpu
def test_mlp_activation_fp16_option():
    """config.mlp_activation_dtype = 'fp16': under no_grad the c_fc epilogue stores gelu(.) as float16 and c_proj's
    statistics / quantise kernels read it directly.  Against the float32 default on the same weights: 32-bit logits
    within 1e-3 (one fp16 rounding of a 4C-wide intermediate per block), 8-bit logits in the same direction (code
    flips), calibrated statistics of the c_proj inputs within one fp16 ulp; with autograd on the option is ignored."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    outs = {}
    for mode in ("fp32", "fp16"):
        torch.manual_seed(6)
        cfg = _tiny_config()
        cfg.mlp_activation_dtype = mode
        model = SPLMHeadModel(cfg).cuda().eval()
        ids = torch.randint(0, 211, (2, 32), device="cuda")
        with torch.no_grad():
            model.set_precision(32)
            l32 = model(ids).float()
        _calibrate_model(model, 8, [ids])
        with torch.no_grad():
            l8 = model(ids).float()
        q = model.transformer.h[0].mlp.c_proj.quantizers_input["8bit"]
        model.set_precision(32)
        xg = model.transformer.wte(ids).detach().requires_grad_(True)
        out = model(inputs_embeds=xg)
        (out["logits"] if isinstance(out, dict) else out).sum().backward()
        outs[mode] = (l32, l8, q.running_max.clone(), xg.grad.clone())
    a, b = outs["fp32"], outs["fp16"]
    assert float((a[0] - b[0]).norm() / a[0].norm()) <= 1e-3
    cos = float((a[1] * b[1]).sum() / a[1].norm() / b[1].norm())
    assert cos >= 0.995, cos
    assert float((a[2] - b[2]).abs().max()) <= 2e-3          # log2-domain statistics: one fp16 ulp of the value is 7e-4
    assert torch.equal(a[3], b[3])                           # training path untouched


@pytest.mark.gpu
@pytest.mark.parametrize("bits,per_channel", [(32, True), (8, True), (4, True), (4, False)])
def test_layernorm_fused_into_consumer_matches_unfused(bits, per_channel, monkeypatch):
    """Under no_grad SPBlock hands ln_1 / ln_2 to c_attn / c_fc (`pre_norm`), which normalise inside their activation-side
    kernel (spq_ln_quantize_act / spq_ln_rowscale_stats).  Against the unfused path (SPQ_FUSE_LN=0 semantics) on the same
    weights: calibrated statistics of the LayerNorm-fed quantisers and the logits agree to fp32 rounding of the LayerNorm
    (the two kernels reduce a row in a different order) -- 32-bit logits within 1e-5, quantised logits within the
    code-flip bar of the tiny-model test; with autograd on nothing is fused."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    import llm_qat_on_gpt2_b200.lora as lora_mod
    outs = {}
    for fused in (True, False):
        monkeypatch.setattr(lora_mod, "_FUSE_LN", fused)
        torch.manual_seed(8)
        cfg = _tiny_config()
        # per_channel=False at 4 bits is upstream's evaluation configuration (p1/deploy.py:210,238): per-tensor scales,
        # e4m3 integer-code operands -- the fused kernel then writes one byte per code
        cfg.per_channel_quantization = per_channel
        model = SPLMHeadModel(cfg).cuda().eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("lora_B"):
                    p.normal_(0, 0.02)
                if ".weights." in n or ".biases." in n:
                    p.add_(0.1 * torch.randn_like(p))
        ids = torch.randint(0, 211, (2, 32), device="cuda")
        if bits < 32:
            _calibrate_model(model, bits, [ids])
            q = model.transformer.h[1].attn.c_attn.quantizers_input[f"{bits}bit"]
            stats = (q.running_min.clone(), q.running_max.clone())
        else:
            model.set_precision(32)
            stats = None
        n0 = lora_mod._lib.launch_count()
        with torch.no_grad():
            logits = model(ids).float()
        outs[fused] = (logits, stats, lora_mod._lib.launch_count() - n0)
    a, b = outs[True], outs[False]
    assert a[2] < b[2]                                       # fewer launches: the LayerNorm kernels are gone
    rel = float((a[0] - b[0]).norm() / b[0].norm())
    if bits == 32:
        assert rel <= 1e-5, rel
    else:
        assert float((a[1][0] - b[1][0]).abs().max()) <= 1e-4 and float((a[1][1] - b[1][1]).abs().max()) <= 1e-4
        cos = float((a[0] * b[0]).sum() / a[0].norm() / b[0].norm())
        assert cos >= 0.999, (rel, cos)


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [8, 4])
def test_gelu_backward_fused_into_rowscale(bits, monkeypatch):
    """Under autograd with frozen base weights, c_fc's autograd Function returns gelu(y) and its backward folds gelu'(y)
    into the row-scaling pass of the incoming gradient (spq_rowscale_dgelu_f16_max).  Against torch's GELU after the call
    (SPQ_FUSE_DGELU=0 semantics) on the same model: identical logits, LoRA / LayerNorm / input gradients within the
    fp16-operand tolerance; with a trainable base weight the fusion stays off."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    import llm_qat_on_gpt2_b200.lora as lora_mod
    res = {}
    for fused in (True, False):
        monkeypatch.setattr(lora_mod, "_FUSE_DGELU", fused)
        torch.manual_seed(12)
        model = SPLMHeadModel(_tiny_config()).cuda().train()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("lora_B"):
                    p.normal_(0, 0.02)
        ids = torch.randint(0, 211, (2, 32), device="cuda")
        _calibrate_model(model, bits, [ids])
        for n, p in model.named_parameters():
            p.requires_grad_(("lora_" in n) or (".weights." in n) or (".biases." in n))
        xe = model.transformer.wte(ids).detach().requires_grad_(True)
        n0 = lora_mod._lib.launch_count()
        out = model(inputs_embeds=xe, labels=ids)
        out["loss"].backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        grads["inputs_embeds"] = xe.grad.clone()
        res[fused] = (out["logits"].detach().float().clone(), grads, lora_mod._lib.launch_count() - n0)
    a, b = res[True], res[False]
    assert torch.equal(a[0], b[0])
    assert a[1].keys() == b[1].keys() and len(a[1]) > 10
    for n in a[1]:
        ga, gb = a[1][n].double(), b[1][n].double()
        assert float((ga - gb).norm() / gb.norm().clamp_min(1e-30)) <= 2e-3, n
    # trainable base weight: the Function needs the plain gradient for dW -> torch's GELU stays outside
    monkeypatch.setattr(lora_mod, "_FUSE_DGELU", True)
    model.transformer.h[0].mlp.c_fc.linear.weight.requires_grad_(True)
    model.zero_grad()
    model(ids, labels=ids)["loss"].backward()
    assert model.transformer.h[0].mlp.c_fc.linear.weight.grad is not None


def test_sp_linear_fp8_path_per_tensor_4bit(monkeypatch):
    """The evaluation configuration (per_channel=False, 4-bit min-max: p1/deploy.py:210,238): SPLinearWithLoRA takes the
    e4m3 integer-code GEMM.  With the LoRA branch off the output is the exact product of the codes times s_x s_w (fp32
    rounding only); with it on the result agrees with float64 to the fp16 LoRA-operand rounding; SPQ_FP8=0 gives the
    fp16-operand path, which agrees to 1e-3."""
    from llm_qat_on_gpt2_b200 import SPLinearWithLoRA, quantize_codes
    K, N, r, bits, M = 768, 2304, 64, 4, 1000

    def build():
        torch.manual_seed(4)
        m = SPLinearWithLoRA(K, N, bit_widths=[bits, 32], lora_rank_per_bit={bits: r, 32: 0}, lora_alpha_per_bit={bits: r, 32: 0},
                             quantizer_per_bit={bits: "minmax", 32: None}, per_channel=False).cuda()
        with torch.no_grad():
            m.linear.weight.normal_(0, 0.02); m.lora_adapters[f"{bits}bit"].lora_B.normal_(0, 0.02)
        x = torch.randn(M, K, device="cuda") * 2
        _calibrate_linear(m, bits, [x[None]])
        return m, x
    m, x = build()
    key = f"{bits}bit"
    qi, qw, lo = m.quantizers_input[key], m.quantizers_weight[key], m.lora_adapters[key]
    assert qi.scale.numel() == 1 and qw.scale.numel() == 1
    with torch.no_grad():
        y = m(x[None])[0]
        base, lora = m._operands_for(bits, True)
        assert base["f8"] is not None and base["f8"]["B8"].dtype == torch.uint8 and lora["Bl_op8"] is not None
        m.calibration_mode = True
        y_base = m(x[None])[0]
        m.calibration_mode = False
        _, cx, _ = quantize_codes(x, qi.scale, qi.zero_point, bits, True, "minmax")
        _, cw, _ = quantize_codes(m.linear.weight.data, qw.scale, qw.zero_point, bits, True, "minmax")
        exact = (cx.double() @ cw.double().t()) * float(qi.scale) * float(qw.scale) + m.linear.bias.double()
        assert ((y_base.double() - exact).norm() / exact.norm()) <= 3e-7
        aq = lo.quantize_A(lo.lora_A).double(); bq = lo.quantize_B(lo.lora_B).double()
        full = exact + (x.double() @ aq @ bq) * lo.scaling
        assert ((y.double() - full).norm() / full.norm()) <= 5e-4
    monkeypatch.setenv("SPQ_FP8", "0")
    m2, x2 = build()
    with torch.no_grad():
        y2 = m2(x2[None])[0]
        assert m2._operands_for(bits, True)[0]["f8"] is None
    assert ((y2.double() - full).norm() / full.norm()) <= 1e-3
    # the backward is the fp16-operand STE path either way
    monkeypatch.setenv("SPQ_FP8", "1")
    m.linear.weight.requires_grad_(False); m.linear.bias.requires_grad_(False)
    xg = x.clone().requires_grad_(True)
    out = m(xg[None]); out.backward(torch.randn_like(out) * 1e-2)
    assert torch.isfinite(xg.grad).all() and lo.lora_A.grad is not None and lo.lora_B.grad is not None
