"""GPU parity against the UNMODIFIED upstream code (baseline/_ref, installed by tools/install_ref.py) executed on
the same box, on torch-CPU and on torch-CUDA (fp32, TF32 off), on identical inputs.  pytest -m gpu.

 (a) end-to-end calibrate -> quantise: calibrated statistics / scale / zero-point and integer codes of this repo's
     kernels vs upstream, as mismatch RATES, at the shapes of all four GPT-2-small linears (activations and weights),
     4-bit min-max and 8-bit log.
 (b) all 48 linears of GPT-2 small teacher-forced: every SPLinearWithLoRA is fed the input the upstream model fed
     its twin and must reproduce the twin's output to rel 1e-3 -- through the plain call, and through the internal
     fast paths the model wrapper uses (fp16 output, fused exact GELU, residual epilogue, fp16 input).
 (c) model-level gradients (LoRA A/B, LayerNorm pairs, inputs_embeds) of a 2-layer model vs upstream autograd.

Upstream has no golden vectors for this path; these tests pin parity by running upstream itself.
"""
import contextlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import upstream as up

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not up.available(), reason="baseline/_ref not installed")]
TOL = 1e-3
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _report(name, payload):
    """Measured rates are kept for profiles/ (the assertions below bound them)."""
    try:
        os.makedirs(OUT_DIR, exist_ok=True)
        with open(os.path.join(OUT_DIR, "upstream_parity.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **payload}) + "\n")
    except OSError:
        pass


@contextlib.contextmanager
def record_round(store):
    orig = torch.round

    def rec(t, *a, **k):
        r = orig(t, *a, **k)
        store.append(r.detach().clone())
        return r
    torch.round = rec
    try:
        yield
    finally:
        torch.round = orig


def heavy(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * torch.exp(1.2 * torch.randn(*shape, generator=g)) * scale
    if shape[-1] >= 8:
        x[..., 3] *= 20.0
        x[..., shape[-1] - 2] *= 0.01
    flat = x.view(-1)
    flat[::97] = 0.0
    flat[5::211] = 3e-6
    return x.contiguous()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ulps(a, b):
    a = a.detach().cpu().contiguous().view(torch.int32).long()
    b = b.detach().cpu().contiguous().view(torch.int32).long()
    return (a - b).abs()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


# ------------------------------------------------------------------------------------------------- (a)
LINEAR_SHAPES = [(768, 2304), (768, 768), (768, 3072), (3072, 768)]


def _upstream_quantise(tensor_batches, test_tensor, bits, qtype, channel_dim, is_input, device):
    UQ = up.p1("quantization").LearnableFakeQuantize
    q = UQ(bits, channel_dim=channel_dim, quantizer_type=qtype, is_input=is_input).to(device)
    with up.quiet(), torch.no_grad():
        q.start_calibration()
        for t in tensor_batches:
            q(t.to(device))
        q.finish_calibration()
        store = []
        with record_round(store):
            out = q(test_tensor.to(device))
    n = 2 ** (bits - 1) - 1
    codes = torch.clamp(store[-1], -n, n).to(torch.int32)
    return q, out, codes


@pytest.mark.parametrize("bits,qtype", [(4, "minmax"), (8, "log")])
@pytest.mark.parametrize("K,N", LINEAR_SHAPES)
def test_calibrate_then_quantise_vs_upstream(K, N, bits, qtype):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, quantize_codes
    cases = {"input": ([heavy((4, 256, K), 11), heavy((4, 256, K), 12)], heavy((4, 256, K), 13), -1, True),
             "weight": ([torch.randn(N, K, generator=torch.Generator().manual_seed(5)) * 0.02], None, 0, False)}
    for what, (batches, test, cd, is_in) in cases.items():
        test = batches[0] if test is None else test
        ours = LearnableFakeQuantize(bits, channel_dim=cd, quantizer_type=qtype, is_input=is_in).cuda()
        with torch.no_grad():
            ours.start_calibration()
            for t in batches:
                ours(t.cuda())
            ours.finish_calibration()
            _, codes, sign = quantize_codes(test.cuda(), ours.scale, ours.zero_point, bits, True, qtype)
        for ref_dev in ("cpu", "cuda"):
            q, out, ref_codes = _upstream_quantise(batches, test, bits, qtype, cd, is_in, ref_dev)
            stat = {}
            for name in ("running_min", "running_max", "scale", "zero_point"):
                a, b = getattr(ours, name), getattr(q, name)
                assert tuple(a.shape) == tuple(b.shape), (what, name)
                d = ulps(a, b)
                stat[name] = {"max_ulp": int(d.max()), "frac_differ": float((d > 0).float().mean())}
            mism = (codes.cpu() != ref_codes.cpu())
            rate = float(mism.float().mean())
            off_by = int((codes.cpu() - ref_codes.cpu()).abs().max())
            _report("calibrate_then_quantise", {"K": K, "N": N, "bits": bits, "qtype": qtype, "tensor": what,
                                                "upstream_device": ref_dev, "stats": stat, "code_mismatch_rate": rate,
                                                "max_code_distance": off_by, "elements": int(mism.numel())})
            if qtype == "minmax":
                # min / max are order independent: bit-exact against upstream on either device
                assert stat["running_min"]["max_ulp"] == 0 and stat["running_max"]["max_ulp"] == 0, (what, ref_dev, stat)
                if ref_dev == "cpu":
                    # true IEEE division, as torch-CPU: scale and every code identical
                    assert stat["scale"]["max_ulp"] == 0 and rate == 0.0, (what, stat, rate)
                else:
                    # torch-CUDA computes tensor / python_scalar as tensor * (1 / scalar): <= 1 ulp on the scale,
                    # which moves codes only at rounding ties
                    assert stat["scale"]["max_ulp"] <= 1 and rate <= 1e-5 and off_by <= 1, (what, stat, rate)
            else:
                # log2 is not correctly rounded on either upstream device (SLEEF / libdevice); this repo uses the
                # correctly rounded value: statistics within 1 ulp (scale = difference of two: 4), codes differ
                # only where the pre-rounding level sits within an ulp of a tie
                assert stat["running_min"]["max_ulp"] <= 1 and stat["running_max"]["max_ulp"] <= 1, (what, ref_dev, stat)
                assert stat["scale"]["max_ulp"] <= 4 and stat["zero_point"]["max_ulp"] <= 1, (what, ref_dev, stat)
                assert rate <= (0.0 if ref_dev == "cpu" else 1e-5) and off_by <= 1, (what, ref_dev, rate, off_by)   # measured: 0 / <= 4.3e-7


# ------------------------------------------------------------------------------------------------- (b), (c)
def _calibrate_like_upstream(model, bits, batches, device):
    """p1/train_sp.py:47-163 through upstream's own CalibrationManager (weights -> inputs with LoRA off), then the
    LoRA quantisers (calibrate_lora_only)."""
    CalibrationManager = up.p1_bare("train_sp").CalibrationManager
    loader = [{"input_ids": b} for b in batches]
    mgr = CalibrationManager(model, loader, device)
    with up.quiet():
        model.set_precision(bits)
        mgr._calibrate_precision(bits, num_batches=len(loader))
        mgr.calibrate_lora_only(bits)
    return mgr


def _make_pair(n_layer, bits_list, seed=0, lora_b_std=0.02):
    """(upstream model on cuda, this repo's model on cuda) with identical parameters and calibration."""
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    cfg = up.gpt2_config(n_layer=n_layer, bit_widths=tuple(bits_list) + (32,))
    torch.manual_seed(seed)
    with up.quiet():
        ref = up.p1("models_sp").SPLMHeadModel(cfg).cuda().eval()
    with torch.no_grad():
        ref.transformer.wte.weight.normal_(0, 0.02)
        ref.transformer.wpe.weight.normal_(0, 0.01)
        for n, p in ref.named_parameters():
            if n.endswith("lora_B"):
                p.normal_(0, lora_b_std)
            if "weights." in n:
                p.add_(0.1 * torch.randn_like(p))
            if "biases." in n:
                p.add_(0.05 * torch.randn_like(p))
    g = torch.Generator().manual_seed(seed + 1)
    calib = [torch.randint(0, cfg.vocab_size, (2, 128), generator=g).cuda() for _ in range(2)]
    for b in bits_list:
        _calibrate_like_upstream(ref, b, calib, torch.device("cuda"))
    cfg2 = up.gpt2_config(n_layer=n_layer, bit_widths=tuple(bits_list) + (32,))
    cfg2.attention_dtype = "fp32"
    with up.quiet():
        ours = SPLMHeadModel(cfg2).cuda().eval()
        missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)     # upstream checkpoint, strict
    assert not missing and not unexpected
    return ref, ours, cfg


def test_all_48_linears_teacher_forced():
    """GPT-2 small, 4-bit min-max and 8-bit log.  Every linear of this repo's model gets the input upstream's twin
    saw (captured with forward hooks on the unmodified upstream model running on torch-CUDA fp32) and must return
    the twin's output to rel <= 1e-3: (i) plain forward; (ii) the fast paths SPBlock uses under no_grad -- fp16
    output (c_attn), fused exact GELU (c_fc), residual epilogue (both c_proj), fp16 input (attn.c_proj)."""
    import torch.nn.functional as F
    ref, ours, cfg = _make_pair(12, (4, 8))
    ids = torch.randint(0, cfg.vocab_size, (2, 192), generator=torch.Generator().manual_seed(77)).cuda()
    worst = {}
    for bits in (4, 8):
        captured = {}
        hooks = []
        for name, mod in ref.named_modules():
            if mod.__class__.__name__ == "SPLinearWithLoRA":
                hooks.append(mod.register_forward_hook(
                    lambda m, inp, out, _n=name: captured.__setitem__(_n, (inp[0].detach().clone(), out.detach().clone()))))
        with up.quiet(), torch.no_grad():
            ref.set_precision(bits)
            ours.set_precision(bits)
            ref(ids)
        for h in hooks:
            h.remove()
        assert len(captured) == 48
        mine = dict(ours.named_modules())
        errs = {}
        for name, (x, y) in captured.items():
            m = mine[name]
            with torch.no_grad():
                e = {"plain": rel(m(x), y)}
                res = torch.randn_like(y)
                if name.endswith("attn.c_attn"):
                    e["fp16_out"] = rel(m(x, out_half=True).float(), y)
                if name.endswith("mlp.c_fc"):
                    e["fused_gelu"] = rel(m(x, fuse_gelu=True), F.gelu(y))
                if name.endswith("c_proj"):
                    e["residual"] = rel(m(x, residual=res), res + y)
                if name.endswith("attn.c_proj"):
                    # the fp16 attention output as input: upstream is given the same (fp16-representable) values
                    xh = x.half()
                    with up.quiet():
                        yh = dict(ref.named_modules())[name](xh.float())
                    e["fp16_in"] = rel(m(xh, residual=res), res + yh)
            # with autograd on (training): the unfused compositions
            xg = x.clone().requires_grad_(True)
            e["grad_mode"] = rel(m(xg), y)
            errs[name] = e
        flat = {f"{n}:{k}": v for n, e in errs.items() for k, v in e.items()}
        w = max(flat, key=flat.get)
        worst[bits] = (w, flat[w])
        _report("teacher_forced_48", {"bits": bits, "worst": w, "worst_rel": flat[w],
                                      "median_rel": float(np.median(list(flat.values()))), "checks": len(flat)})
        bad = {k: v for k, v in flat.items() if not v <= TOL}
        assert not bad, (bits, bad)
    print("teacher-forced worst:", worst)


def _loss_and_grads(model, bits, ids, autocast=False):
    with up.quiet():
        model.set_precision(bits)
    model.zero_grad(set_to_none=True)
    for n, p in model.named_parameters():
        p.requires_grad_(f"lora_adapters.{bits}bit.lora_" in n or n.endswith(f"weights.{bits}") or n.endswith(f"biases.{bits}"))
    emb = (model.transformer.wte(ids) + model.transformer.wpe(torch.arange(ids.shape[1], device="cuda")[None])).detach()
    emb.requires_grad_(True)
    with up.quiet(), torch.amp.autocast("cuda", enabled=autocast):
        out = model(inputs_embeds=emb, labels=ids)
    out["loss"].backward()
    t = {"loss": out["loss"].detach().float().reshape(1), "logits": out["logits"].detach().float(),
         "inputs_embeds.grad": emb.grad.detach().float()}
    t.update({n: p.grad.detach().float() for n, p in model.named_parameters() if p.grad is not None})
    return t


def test_model_gradients_vs_upstream_autograd():
    """2-layer model, CE loss: logits and the gradients of the active LoRA A/B, of the active LayerNorm pairs and of
    inputs_embeds against upstream autograd on torch-CUDA fp32.

    The per-GEMM bar (rel 1e-3, teacher-forced test above; measured 3-4e-4) does not compose to 1e-3 through a stack
    of layers for ANY implementation that is not bit-identical to torch fp32 eager: the fp16 operand rounding adds in
    quadrature per layer, and below 16 bits the downstream quantisers turn it into code flips (a level is 4 % of the
    value at 8-bit log, 14 % of the range at 4-bit min-max).  Upstream's own training numerics -- the same model
    under torch.amp.autocast, p1/train_sp.py:319 -- are the yardstick: no tensor may deviate from upstream-fp32 by
    more than max(1e-3, 1.5 x the deviation of upstream-under-autocast on the same inputs), and the median deviation
    must not exceed upstream-autocast's.  16-bit log (levels as fine as the operand rounding: flips cost nothing)
    isolates the backward chain itself (STE, LoRA / LayerNorm gradients, LM head, CE)."""
    ref, ours, cfg = _make_pair(2, (4, 8, 16), seed=3)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, cfg.vocab_size, (2, 96), generator=g).cuda()
    for bits in (16, 8, 4):
        t_ref = _loss_and_grads(ref, bits, ids)
        t_amp = _loss_and_grads(ref, bits, ids, autocast=True)
        t_our = _loss_and_grads(ours, bits, ids)
        assert set(t_ref) == set(t_our) and len(t_ref) == 3 + 2 * 4 * 2 + 5 * 2
        e_our = {n: rel(t_our[n], t_ref[n]) for n in t_ref}
        e_amp = {n: rel(t_amp[n], t_ref[n]) for n in t_ref}
        w = max(e_our, key=e_our.get)
        _report("model_gradients", {"bits": bits, "worst": w, "worst_rel": e_our[w], "worst_rel_upstream_amp": max(e_amp.values()),
                                    "median_rel": float(np.median(list(e_our.values()))),
                                    "median_rel_upstream_amp": float(np.median(list(e_amp.values()))), "tensors": len(e_our)})
        _report("model_gradients_detail", {"bits": bits, "ours": e_our, "upstream_amp": e_amp})
        bad = {k: (v, e_amp[k]) for k, v in e_our.items() if not v <= max(TOL, 1.5 * e_amp[k])}
        assert float(np.median(list(e_our.values()))) <= max(TOL, float(np.median(list(e_amp.values())))), (bits, e_our, e_amp)
        assert not bad, (bits, bad)


def test_all_linears_teacher_forced_backward():
    """The backward twin of the 48-linear test, on a 2-layer model at 16 / 8 / 4 bits: every SPLinearWithLoRA gets
    the (x, dY) pair upstream's twin saw during a real CE backward (captured with hooks on the unmodified upstream
    model, torch-CUDA fp32) and must return upstream's dX, dA (lora_A.grad) and dB (lora_B.grad) to rel 1e-3.
    Real in-model gradients have all-zero rows (the last position of every sequence carries no loss), widely
    different magnitudes per token and per q/k/v block -- which synthetic dY never had (regression: zero rows used
    to report row scale 1 and pushed the token-reduction operand of dA into the fp16 subnormal range, 5e-3..1e-2)."""
    ref, ours, cfg = _make_pair(2, (4, 8, 16), seed=3)
    ids = torch.randint(0, cfg.vocab_size, (2, 96), generator=torch.Generator().manual_seed(5)).cuda()
    mine, refm = dict(ours.named_modules()), dict(ref.named_modules())
    for bits in (16, 8, 4):
        key = f"{bits}bit"
        cap, hooks = {}, []
        for name, mod in ref.named_modules():
            if mod.__class__.__name__ == "SPLinearWithLoRA":
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, _n=name: cap.setdefault(_n, {}).update(x=i[0].detach().clone(), y=o.detach().clone())))
                hooks.append(mod.register_full_backward_hook(
                    lambda m, gi, go, _n=name: cap[_n].update(gy=go[0].detach().clone(),
                                                              gx=None if gi[0] is None else gi[0].detach().clone())))
        t_ref = _loss_and_grads(ref, bits, ids)
        for h in hooks:
            h.remove()
        with up.quiet():
            ours.set_precision(bits)
        assert len(cap) == 8
        errs = {}
        for name, c in cap.items():
            m = mine[name]
            lo, rl = m.lora_adapters[key], refm[name].lora_adapters[key]
            for p in (lo.lora_A, lo.lora_B):
                p.requires_grad_(True)
                p.grad = None
            assert float(c["gy"].reshape(-1, c["gy"].shape[-1]).abs().amax(dim=1).min()) == 0.0     # zero rows are there
            x = c["x"].clone().requires_grad_(True)
            y = m(x)
            y.backward(c["gy"])
            errs[f"{name}:y"] = rel(y, c["y"])
            errs[f"{name}:dA"] = rel(lo.lora_A.grad, rl.lora_A.grad)
            errs[f"{name}:dB"] = rel(lo.lora_B.grad, rl.lora_B.grad)
            if c["gx"] is not None:
                errs[f"{name}:dx"] = rel(x.grad, c["gx"])
        w = max(errs, key=errs.get)
        _report("teacher_forced_backward", {"bits": bits, "worst": w, "worst_rel": errs[w],
                                            "median_rel": float(np.median(list(errs.values()))), "checks": len(errs)})
        bad = {k: v for k, v in errs.items() if not v <= TOL}
        assert not bad, (bits, bad)


def test_exported_integer_weights_equal_upstream_quantised_weights():
    """deploy.export_integer_weights against the UNMODIFIED upstream weight quantiser on torch-CPU (not against this
    repo's own dequant): 4-bit min-max codes * scale reproduce upstream's q_w(W) bit for bit, packed int4 included;
    8-bit log levels + signs reproduce upstream's level indices."""
    from llm_qat_on_gpt2_b200.deploy import export_integer_weights, unpack_int4
    ref, ours, cfg = _make_pair(2, (4, 8), seed=9)
    refm = dict(ref.named_modules())
    for bits in (4, 8):
        # calibrate the weight quantisers with this repo's kernels (true IEEE division = torch-CPU semantics; the pair
        # was calibrated by upstream on torch-CUDA, whose scale is 1 ulp off its own CPU result)
        with torch.no_grad():
            for m in ours.modules():
                if m.__class__.__name__ == "SPLinearWithLoRA":
                    qw = m.quantizers_weight[f"{bits}bit"]
                    qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
        exp = export_integer_weights(ours, bits, packed=True)
        names = [n for n, m in ours.named_modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
        assert len(names) == 8
        for n in names:
            rm = refm[n]
            W = rm.linear.weight.detach().cpu()
            UQ = up.p1("quantization").LearnableFakeQuantize
            q = UQ(bits, channel_dim=0, quantizer_type=cfg.quantizer_per_bit[bits])
            with up.quiet(), torch.no_grad():
                q.start_calibration(); q(W); q.finish_calibration()
                store = []
                with record_round(store):
                    wq = q(W)
            nmax = 2 ** (bits - 1) - 1
            ref_codes = torch.clamp(store[-1], -nmax, nmax).to(torch.int32)
            if bits == 4:
                codes = unpack_int4(exp[f"{n}.codes_packed"], W.shape[1]).to(torch.int32)
                assert exp[f"{n}.codes_shape"] == tuple(W.shape)
                assert torch.equal(codes, ref_codes)
                assert torch.equal(codes.float() * exp[f"{n}.scale"], wq)            # upstream's fake-quantised weight, exactly
            else:
                assert torch.equal(exp[f"{n}.codes"].to(torch.int32), ref_codes)
                assert torch.equal(exp[f"{n}.sign"].float(), torch.sign(wq) * (W.abs() >= 1e-5))
