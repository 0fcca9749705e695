#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference and does
not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every array written here is an input fed to, or an output produced by, the
reference's own classes (part1_switchable_precision.{quantization,
quantization_methods, lora, switchable_batchnorm, models_sp}) on CPU, float32,
torch.manual_seed-seeded.  Nothing is computed by this repo's code.  The level
index of the log quantiser is captured by recording what ``torch.round``
returned inside the reference's forward (p1/quantization_methods.py:54/:59) and
applying the clamp the reference applies on the next line.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "part1_switchable_precision"))
sys.dont_write_bytecode = True

from part1_switchable_precision.quantization import LearnableFakeQuantize  # noqa: E402
from part1_switchable_precision.lora import SPLinearWithLoRA  # noqa: E402
from part1_switchable_precision.switchable_batchnorm import SwitchableLayerNorm  # noqa: E402
from part1_switchable_precision.models_sp import SPLMHeadModel  # noqa: E402
from transformers import GPT2Config  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


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


def make_input(shape, seed, outlier_cols=True, zeros=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g)
    # heavy-tailed magnitudes so log quantisation sees many binades
    x = x * torch.exp(1.5 * torch.randn(*shape, generator=g))
    if outlier_cols and shape[-1] >= 8:
        x[..., 3] *= 20.0
        x[..., shape[-1] - 2] *= 0.01
    if zeros:
        flat = x.view(-1)
        idx = torch.randperm(flat.numel(), generator=g)[: max(1, flat.numel() // 50)]
        flat[idx[: len(idx) // 2]] = 0.0
        flat[idx[len(idx) // 2:]] = 3e-6        # below the 1e-5 zero threshold
    return x.contiguous()


def quantizer_case(name, shape, channel_dim, qtype, bits, symmetric=True, per_channel=True,
                   is_input=False, nbatch=2, seed=0):
    q = LearnableFakeQuantize(bits, channel_dim=channel_dim, quantizer_type=qtype,
                              symmetric=symmetric, per_channel=per_channel, is_input=is_input)
    xs = [make_input(shape, seed + i) for i in range(nbatch)]
    q.start_calibration()
    for x in xs:
        y = q(x)
        assert y is x                           # collecting mode passes x through
    temp_min = q.temp_min.clone()
    temp_max = q.temp_max.clone()
    with quiet():
        q.finish_calibration()
    xt = make_input(shape, seed + 100)          # a batch that partly exceeds the calibrated range
    xt.view(-1)[:7] = torch.tensor([0.0, 1e-5, -1e-5, 9.9e-6, 1e30, -1e30, 1e-38])
    rounds = []
    with record_round(rounds):
        out = q(xt)
    d = {
        "x_calib": torch.stack(xs).numpy(), "x_test": xt.numpy(), "out": out.numpy(),
        "temp_min": temp_min.numpy(), "temp_max": temp_max.numpy(),
        "running_min": q.running_min.numpy(), "running_max": q.running_max.numpy(),
        "scale": q.scale.numpy(), "zero_point": q.zero_point.numpy(),
        "meta": np.array([bits, -99 if channel_dim is None else channel_dim, int(symmetric),
                          int(per_channel), int(is_input), 0 if qtype == "minmax" else 1]),
    }
    r = rounds[-1]
    if qtype == "minmax":
        lo, hi = (-(2 ** (bits - 1) - 1), 2 ** (bits - 1) - 1) if symmetric else (0, 2 ** bits - 1)
    else:
        lo, hi = (-(2 ** (bits - 1) - 1), 2 ** (bits - 1) - 1) if symmetric else (0, 2 ** bits - 1)
    d["codes"] = torch.clamp(r, lo, hi).to(torch.int32).numpy()
    np.savez_compressed(os.path.join(OUT, f"quant_{name}.npz"), **d)
    print(f"quant_{name}: stats {tuple(q.scale.shape)} codes [{d['codes'].min()}, {d['codes'].max()}]")


def zero_tensor_case():
    """Fresh lora_B (all zeros) through a log quantiser: the reference's default-shape
    quirk (p1/quantization.py:164-172, 194-197)."""
    q = LearnableFakeQuantize(8, channel_dim=1, quantizer_type="log")
    x = torch.zeros(8, 24)
    q.start_calibration(); q(x)
    with quiet():
        q.finish_calibration()
    out = q(x)
    np.savez_compressed(os.path.join(OUT, "quant_log_allzero.npz"), x=x.numpy(), out=out.numpy(),
                        running_min=q.running_min.numpy(), running_max=q.running_max.numpy(),
                        scale=q.scale.numpy(), zero_point=q.zero_point.numpy())
    print("quant_log_allzero: stats shape", tuple(q.scale.shape))


def _calibrate_linear(m, bits, x_batches):
    key = f"{bits}bit"
    with quiet():
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


def linear_case(name, K, N, r, bits, qtype, B=2, T=24, per_channel=True, seed=0):
    torch.manual_seed(seed)
    m = SPLinearWithLoRA(K, N, bit_widths=[bits, 32], lora_rank_per_bit={bits: r, 32: 0},
                         lora_alpha_per_bit={bits: 2 * r, 32: 0}, quantizer_per_bit={bits: qtype, 32: None},
                         per_channel=per_channel)
    key = f"{bits}bit"
    with torch.no_grad():
        m.lora_adapters[key].lora_B.normal_(0, 0.05)      # make the LoRA branch non-trivial
        m.linear.bias.normal_(0, 0.1)
    xc = [make_input((B, T, K), seed + 10 + i, zeros=False) * 0.5 for i in range(2)]
    _calibrate_linear(m, bits, xc)
    for p in m.parameters():
        p.requires_grad_(True)
    x = (make_input((B, T, K), seed + 50, zeros=True) * 0.5).requires_grad_(True)
    y = m(x)
    g = torch.Generator().manual_seed(seed + 7)
    gy = torch.randn(y.shape, generator=g) * 0.1
    gy.view(-1)[:3] = torch.tensor([25.0, -40.0, 11.0])    # exercises the +-10 clamp of the log STE
    y.backward(gy)
    lo = m.lora_adapters[key]
    m.set_precision(32)
    with torch.no_grad():
        y32 = m(x)
        m.set_precision(bits)
        m.calibration_mode = True
        ybase = m(x)
        m.calibration_mode = False
    d = {
        "x_calib": torch.stack(xc).numpy(), "x": x.detach().numpy(), "weight": m.linear.weight.detach().numpy(),
        "bias": m.linear.bias.detach().numpy(), "lora_A": lo.lora_A.detach().numpy(),
        "lora_B": lo.lora_B.detach().numpy(), "scaling": np.float32(lo.scaling),
        "y": y.detach().numpy(), "y32": y32.numpy(), "y_base": ybase.numpy(), "grad_y": gy.numpy(),
        "grad_x": x.grad.numpy(), "grad_weight": m.linear.weight.grad.numpy(),
        "grad_bias": m.linear.bias.grad.numpy(), "grad_lora_A": lo.lora_A.grad.numpy(),
        "grad_lora_B": lo.lora_B.grad.numpy(),
        "meta": np.array([K, N, r, bits, 0 if qtype == "minmax" else 1, int(per_channel)]),
    }
    for qn, qq in (("qw", m.quantizers_weight[key]), ("qin", m.quantizers_input[key]),
                   ("qA", lo.quantize_A), ("qB", lo.quantize_B)):
        d[f"{qn}_scale"] = qq.scale.numpy(); d[f"{qn}_zp"] = qq.zero_point.numpy()
        d[f"{qn}_rmin"] = qq.running_min.numpy(); d[f"{qn}_rmax"] = qq.running_max.numpy()
    np.savez_compressed(os.path.join(OUT, f"linear_{name}.npz"), **d)
    print(f"linear_{name}: y {tuple(y.shape)} |y| {y.abs().mean():.4f}")


def layernorm_case():
    torch.manual_seed(3)
    ln = SwitchableLayerNorm(96, precision_levels=[4, 8, 32], eps=1e-5)
    with torch.no_grad():
        for k in ln.weights:
            ln.weights[k].normal_(1.0, 0.2); ln.biases[k].normal_(0, 0.2)
    x = make_input((3, 17, 96), 11, zeros=False).requires_grad_(True)
    d = {"x": x.detach().numpy()}
    gy = torch.randn(3, 17, 96, generator=torch.Generator().manual_seed(5))
    d["grad_y"] = gy.numpy()
    for p in (4, 8, 32):
        ln.set_precision(p)
        x.grad = None; ln.zero_grad()
        y = ln(x); y.backward(gy)
        d[f"w{p}"] = ln.weights[str(p)].detach().numpy(); d[f"b{p}"] = ln.biases[str(p)].detach().numpy()
        d[f"y{p}"] = y.detach().numpy(); d[f"gx{p}"] = x.grad.numpy().copy()
        d[f"gw{p}"] = ln.weights[str(p)].grad.numpy().copy(); d[f"gb{p}"] = ln.biases[str(p)].grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "layernorm.npz"), **d)
    print("layernorm: ok")


def tiny_model_case():
    torch.manual_seed(0)
    bits = [4, 8, 32]
    cfg = GPT2Config(vocab_size=211, n_positions=32, n_embd=64, n_layer=2, n_head=4,
                     layer_norm_epsilon=1e-5, embd_pdrop=0.0)
    cfg.bit_widths = bits
    cfg.lora_rank_per_bit = {4: 8, 8: 8, 32: 0}
    cfg.lora_alpha_per_bit = {4: 16, 8: 16, 32: 0}
    cfg.quantizer_per_bit = {4: "minmax", 8: "log", 32: None}
    cfg.per_channel_quantization = True
    model = SPLMHeadModel(cfg).eval()
    with torch.no_grad():
        model.transformer.wte.weight.mul_(0.1)
        model.transformer.wpe.weight.mul_(0.1)
        for name, p in model.named_parameters():
            if name.endswith("lora_B"):
                p.normal_(0, 0.02)
    g = torch.Generator().manual_seed(42)
    calib = [torch.randint(0, 211, (2, 32), generator=g) for _ in range(2)]
    ids = torch.randint(0, 211, (2, 32), generator=g)
    d = {"calib_ids": torch.stack(calib).numpy(), "ids": ids.numpy()}
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for b in (4, 8):
        key = f"{b}bit"
        with quiet(), torch.no_grad():
            model.set_precision(b)
            mods = [m for m in model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
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
            out = model(ids, output_hidden_states=True)
        d[f"logits{b}"] = out["logits"].numpy()
        d[f"hidden{b}_1"] = out["hidden_states"][1].numpy()
        blk = model.transformer.h[0]
        d[f"qin{b}_c_attn_scale"] = blk.attn.c_attn.quantizers_input[key].scale.numpy()
        d[f"qin{b}_c_attn_zp"] = blk.attn.c_attn.quantizers_input[key].zero_point.numpy()
        d[f"qin{b}_mlp_proj_rmax"] = blk.mlp.c_proj.quantizers_input[key].running_max.numpy()
    with quiet(), torch.no_grad():
        model.set_precision(32)
        d["logits32"] = model(ids).numpy()
    for k, v in sd0.items():
        if v.numel() > 1 and not k.endswith("_quantized") and ".attn.bias" not in k:
            d["sd::" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_model.npz"), **d)
    print("tiny_model: logits", d["logits4"].shape, "keys", len(d))


if __name__ == "__main__":
    quantizer_case("minmax4_weight", (48, 40), 0, "minmax", 4)
    quantizer_case("minmax8_input", (2, 24, 40), -1, "minmax", 8, is_input=True)
    quantizer_case("minmax4_loraA", (40, 8), 1, "minmax", 4, nbatch=1)
    quantizer_case("minmax3_pertensor", (2, 24, 40), -1, "minmax", 3, per_channel=False, is_input=True)
    quantizer_case("minmax8_asym", (48, 40), 0, "minmax", 8, symmetric=False)
    quantizer_case("log8_weight", (48, 40), 0, "log", 8)
    quantizer_case("log8_input", (2, 24, 40), -1, "log", 8, is_input=True)
    quantizer_case("log6_loraB", (8, 40), 1, "log", 6, nbatch=1)
    quantizer_case("log16_input", (2, 24, 40), -1, "log", 16, is_input=True)
    quantizer_case("log5_pertensor", (48, 40), 0, "log", 5, per_channel=False)
    quantizer_case("log8_asym", (48, 40), 0, "log", 8, symmetric=False)
    quantizer_case("log8_big", (4, 128, 96), -1, "log", 8, is_input=True, seed=21)
    zero_tensor_case()
    linear_case("minmax4", 64, 96, 8, 4, "minmax")
    linear_case("log8", 64, 96, 8, 8, "log")
    linear_case("minmax4_pertensor", 64, 96, 8, 4, "minmax", per_channel=False)
    layernorm_case()
    tiny_model_case()
