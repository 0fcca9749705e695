"""Deploy / checkpoint formats on CPU: the integer export equals the unmodified reference's
convert_to_int8 on the same weights (fixture: tests/golden/make_golden_deploy.py), and the per-precision
checkpoint files round-trip through the evaluation loader with strict state_dict loading."""
import os
import types

import numpy as np
import torch
from transformers import GPT2Config

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _tiny_config():
    cfg = GPT2Config(vocab_size=211, n_positions=32, n_embd=64, n_layer=2, n_head=4, layer_norm_epsilon=1e-5, embd_pdrop=0.0)
    cfg.bit_widths = [4, 8, 32]
    cfg.lora_rank_per_bit = {4: 8, 8: 8, 32: 0}
    cfg.lora_alpha_per_bit = {4: 16, 8: 16, 32: 0}
    cfg.quantizer_per_bit = {4: "minmax", 8: "log", 32: None}
    cfg.per_channel_quantization = True
    return cfg


def test_convert_to_int8_matches_reference_fixture():
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from llm_qat_on_gpt2_b200.deploy import convert_to_int8
    g = np.load(os.path.join(GOLDEN, "deploy_int8.npz"))
    model = SPLMHeadModel(_tiny_config()).eval()
    sd = {k[4:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    for bits in (8, 4):
        model.set_precision(bits)
        got = convert_to_int8(model)
        want = {k.split("::", 1)[1]: g[k] for k in g.files if k.startswith(f"int8@{bits}::")}
        assert set(got.keys()) == set(want.keys())
        for k, v in want.items():
            a = got[k].numpy()
            assert a.dtype == v.dtype and a.shape == v.shape, k
            assert np.array_equal(a, v), k


def test_sp_checkpoints_roundtrip(tmp_path):
    from llm_qat_on_gpt2_b200 import SPLMHeadModel
    from llm_qat_on_gpt2_b200.deploy import load_model_for_evaluation, save_int8_checkpoint, save_sp_checkpoints
    cfg = _tiny_config()
    cfg.per_channel_quantization = False              # the evaluation loader rebuilds per-tensor quantisers
    torch.manual_seed(3)
    model = SPLMHeadModel(cfg).eval()
    mc = types.SimpleNamespace(vocab_size=211, n_positions=32, n_embd=64, n_layer=2, n_head=4, layer_norm_epsilon=1e-5,
                               bit_widths=[4, 8, 32], lora_rank_per_bit=cfg.lora_rank_per_bit,
                               lora_alpha_per_bit=cfg.lora_alpha_per_bit, quantizer_per_bit=cfg.quantizer_per_bit)
    saved = save_sp_checkpoints(model, str(tmp_path / "sp"), mc)
    assert sorted(saved) == [4, 8]
    for bits, path in saved.items():
        ck = torch.load(path, map_location="cpu", weights_only=False)
        assert ck["bit_width"] == bits and set(ck) >= {"model_state_dict", "model_config", "training_config", "timestamp"}
        back = load_model_for_evaluation(path, device="cpu")        # strict load_state_dict inside
        assert back.get_current_precision() == bits
        a, b = model.state_dict(), back.state_dict()
        assert a.keys() == b.keys()
        # `*_quantized` scratch buffers are uninitialised memory upstream too (torch.empty): may hold NaN patterns
        assert all(torch.equal(a[k], b[k]) for k in a if not k.endswith("_quantized"))
    ck = save_int8_checkpoint(model, str(tmp_path / "m_int8.pth"), model_config=mc, target_bits=8)
    assert ck["model_info"]["target_bits"] == 8 and ck["model_info"]["compression_ratio"] > 1.0
    assert any(k.endswith("weight_int8") for k in ck["int8_state_dict"])


def test_pack_unpack_int4_roundtrip():
    import torch
    from llm_qat_on_gpt2_b200.deploy import pack_int4, unpack_int4
    g = torch.Generator().manual_seed(0)
    for K in (8, 7, 1):
        c = torch.randint(-7, 8, (5, K), generator=g)
        p = pack_int4(c)
        assert p.dtype == torch.uint8 and p.shape == (5, (K + 1) // 2)
        assert torch.equal(unpack_int4(p, K).to(torch.int64), c)
        u = torch.randint(0, 16, (3, K), generator=g)
        assert torch.equal(unpack_int4(pack_int4(u), K, signed=False).to(torch.int64), u)
    assert int(pack_int4(torch.tensor([[-1, 7]]))[0, 0]) == 0x7F
