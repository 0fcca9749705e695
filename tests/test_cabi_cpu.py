"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/spq_b200.h declares; the host modules keep the reference's state machine, keys and
exceptions; nothing computes without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from llm_qat_on_gpt2_b200 import build, _lib
    build.build()
    return _lib.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "spq_b200.h")).read()
    return sorted(set(re.findall(r"SPQ_API\s+[\w\s\*]+?\b(spq_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from llm_qat_on_gpt2_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in spq_b200.h but not exported by libspq_b200.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.spq_abi_version() == _lib.ABI_VERSION == 6
    assert lib.spq_launch_count() == 0


def test_sass_is_blackwell_native():
    """The GEMM must be tcgen05 + TMA (UTCHMMA / UTMALDG / LDTM in SASS), not a legacy HMMA path."""
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    from llm_qat_on_gpt2_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    assert "sm_100a" in sass


def test_no_cpu_fallback(lib):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, SPLinearWithLoRA, SwitchableLayerNorm, apply_minmax_quantization
    x = torch.randn(2, 8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        apply_minmax_quantization(x, torch.ones(1), torch.zeros(1), 4)
    q = LearnableFakeQuantize(4)
    q.start_calibration()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SwitchableLayerNorm(16, precision_levels=[4, 32])(x)
    m = SPLinearWithLoRA(16, 24, [4, 32], {4: 4, 32: 0}, {4: 8, 32: 0}, {4: "minmax", 32: None})
    m.set_precision(32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)


def test_quantizer_state_machine_matches_reference():
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize
    q = LearnableFakeQuantize(40, channel_dim=1, per_channel=False)
    assert q.num_bits == 32 and q.channel_dim is None
    x = torch.randn(3, 3)
    assert q(x) is x                                   # >= 32 bits: identity (reference :212-213)
    q = LearnableFakeQuantize(8, quantizer_type="log")
    assert (q.quant_min, q.quant_max) == (-128, 127)
    assert tuple(q.scale.shape) == (1,) and not q.calibrated
    with pytest.raises(RuntimeError, match="not calibrated"):
        q(x)
    q.calibrated = True
    q.set_num_bits(8)
    assert q.calibrated                                # unchanged bits: calibration kept (:83-85)
    q.set_num_bits(6)
    assert not q.calibrated and q.quant_max == 31
    q.finish_calibration()                             # nothing collected: stays uncalibrated (:136-139)
    assert not q.calibrated and not q.collecting_stats
    q2 = LearnableFakeQuantize(4, symmetric=False)
    assert (q2.quant_min, q2.quant_max) == (0, 15)
    # checkpoint loading resizes buffers and marks calibrated (:40-75); legacy [B,T,C] input stats fold over T
    src = LearnableFakeQuantize(8, channel_dim=-1, is_input=True)
    sd = {"scale": torch.rand(1, 5, 7) + 1, "zero_point": torch.zeros(1, 5, 7),
          "running_min": -torch.rand(1, 5, 7), "running_max": torch.rand(1, 5, 7)}
    want_min, want_max = sd["running_min"].min(1, keepdim=True)[0], sd["running_max"].max(1, keepdim=True)[0]
    src.load_state_dict(sd)
    assert src.calibrated and tuple(src.scale.shape) == (1, 1, 7)
    assert torch.equal(src.running_min, want_min) and torch.equal(src.running_max, want_max)
    q.quantizer_type = "bogus"; q.calibrated = True
    with pytest.raises(ValueError, match="Unknown quantizer type"):
        q(x)


def _cfg():
    from transformers import GPT2Config
    cfg = GPT2Config(vocab_size=97, n_positions=16, n_embd=32, n_layer=2, n_head=4, layer_norm_epsilon=1e-5, embd_pdrop=0.0)
    cfg.bit_widths = [4, 8, 32]
    cfg.lora_rank_per_bit = {4: 4, 8: 4, 32: 0}
    cfg.lora_alpha_per_bit = {4: 8, 8: 8, 32: 0}
    cfg.quantizer_per_bit = {4: "minmax", 8: "log", 32: None}
    return cfg


def test_model_host_api_matches_reference():
    from llm_qat_on_gpt2_b200 import SPLMHeadModel, SPLinearWithLoRA, SwitchableLayerNorm
    model = SPLMHeadModel(_cfg())
    assert model.lm_head.weight is model.transformer.wte.weight
    assert model.get_current_precision() == 32
    lin = model.transformer.h[0].attn.c_attn
    assert isinstance(lin, SPLinearWithLoRA) and lin.current_bits == 8      # second-largest width
    assert model.set_precision(4) == 4
    ok, details = model.verify_precision_consistency()
    assert ok and details["expected"] == 4
    assert all(m.current_bits == 4 for m in model.modules() if isinstance(m, SPLinearWithLoRA))
    assert all(m.current_precision == 4 for m in model.modules() if isinstance(m, SwitchableLayerNorm))
    with pytest.raises(ValueError):
        model.set_precision(5)
    model.disable_lora_for_calibration()
    assert all(m.calibration_mode for m in model.modules() if isinstance(m, SPLinearWithLoRA))
    model.enable_lora_after_calibration()
    assert not any(m.calibration_mode for m in model.modules() if isinstance(m, SPLinearWithLoRA))
    keys = set(model.state_dict().keys())
    p = "transformer.h.1.mlp.c_proj."
    for k in (p + "weight_quantized", p + "linear.weight", p + "quantizers_input.8bit.zero_point",
              p + "lora_adapters.4bit.quantize_B.running_min", p + "lora_adapters.8bit.lora_A_quantized",
              "transformer.h.0.ln_2.weights.32", "transformer.ln_f.biases.4", "transformer.h.0.attn.bias",
              "transformer.wte.weight", "lm_head.weight"):
        assert k in keys, k
    ln = model.transformer.ln_f
    ln.ln_layers["8"].weight.data = torch.full((32,), 2.0)
    ln.ln_layers["8"].bias.requires_grad = False
    assert torch.equal(ln.weights["8"].data, torch.full((32,), 2.0)) and not ln.biases["8"].requires_grad
    for prm in model.parameters():
        prm.requires_grad = False
    model.transformer.unfreeze_weights(32)
    assert lin.linear.weight.requires_grad and ln.weights["4"].requires_grad
    assert not lin.lora_adapters["4bit"].lora_A.requires_grad
    with pytest.raises(ValueError):
        model.transformer(None)


def test_disabled_lora_and_32bit_only_linear():
    from llm_qat_on_gpt2_b200 import LoRALayer
    off = LoRALayer(8, 12, 0, 0, 32, None)
    assert not off.enabled and off.scaling == 0 and tuple(off.lora_A.shape) == (1, 1)
    assert set(off.state_dict().keys()) == {"lora_A", "lora_B"}
    y = off(torch.randn(2, 3, 8))
    assert y.shape == (2, 3, 12) and torch.count_nonzero(y) == 0
    on = LoRALayer(8, 12, 4, 8, 6, "log")
    assert on.enabled and on.scaling == 2.0 and torch.count_nonzero(on.lora_B) == 0
    assert on.quantize_A.channel_dim == 1 and on.quantize_B.quantizer_type == "log"


def test_side_stream_gradients_are_opt_in():
    """lora.side_stream_grads (the LoRA weight-gradient GEMMs on a side stream) is a training-driver feature: off by
    default, scoped to the `with` block, restored on exit -- plain module users keep every kernel on their own stream."""
    from llm_qat_on_gpt2_b200 import lora
    assert lora._GradSide.stream is None
    with lora.side_stream_grads(None):
        assert lora._GradSide.stream is None
