"""numpy restatement of the GPT-2 wrapper around the path (TEST INFRASTRUCTURE ONLY).

Follows p1/models_sp.py:18-171 (SPAttention / SPMLP / SPBlock), :173-336
(SPModel), :390-458 (SPLMHeadModel) and the calibration procedure of
p1/train_sp.py:47-163 (weights -> LoRA -> inputs with LoRA disabled).
Parameters come from a ``state_dict`` with the reference's key names, as numpy
arrays, so the same weights can drive the reference, the oracle and the CUDA
modules.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import numpy as np
from scipy.special import erf as _erf

from .quant_oracle import F32, QuantizerState, collect_statistics, fake_quantize, finish_calibration
from .layers_oracle import _mm, sp_linear_forward, switchable_layernorm_forward


class _LinearOracle:
    """State of one SPLinearWithLoRA (p1/lora.py:58-103)."""

    def __init__(self, sd, prefix, cfg):
        self.weight = sd[prefix + "linear.weight"].astype(F32)
        self.bias = sd[prefix + "linear.bias"].astype(F32)
        self.q_w: Dict[int, QuantizerState] = {}
        self.q_in: Dict[int, QuantizerState] = {}
        self.lora: Dict[int, Dict] = {}
        pc = cfg.get("per_channel", True)
        for b in cfg["bit_widths"]:
            if b >= 32:
                continue
            qt = cfg["quantizer_per_bit"][b]
            self.q_w[b] = QuantizerState(b, channel_dim=0, quantizer_type=qt, per_channel=pc)
            self.q_in[b] = QuantizerState(b, channel_dim=-1, quantizer_type=qt, per_channel=pc, is_input=True)
            r = cfg["lora_rank_per_bit"][b]
            a = cfg["lora_alpha_per_bit"][b]
            if r > 0:
                self.lora[b] = {
                    "A": sd[f"{prefix}lora_adapters.{b}bit.lora_A"].astype(F32),
                    "B": sd[f"{prefix}lora_adapters.{b}bit.lora_B"].astype(F32),
                    "q_A": QuantizerState(b, channel_dim=1, quantizer_type=qt, per_channel=pc),
                    "q_B": QuantizerState(b, channel_dim=1, quantizer_type=qt, per_channel=pc),
                    "scaling": a / r,
                }
        # p1/lora.py:71
        self.current_bits = sorted(cfg["bit_widths"], reverse=True)[1] if len(cfg["bit_widths"]) > 1 else 32
        self.calibration_mode = False

    def forward(self, x):
        b = self.current_bits
        return sp_linear_forward(x, self.weight, self.bias, b, self.q_in.get(b), self.q_w.get(b),
                                 self.lora.get(b), self.calibration_mode)


class SPModelOracle:
    """SPLMHeadModel restated (p1/models_sp.py:390-458)."""

    def __init__(self, cfg: dict, sd: Dict[str, np.ndarray]):
        self.cfg = cfg
        self.n_layer, self.n_head, self.n_embd = cfg["n_layer"], cfg["n_head"], cfg["n_embd"]
        self.eps = cfg.get("layer_norm_epsilon", 1e-5)
        self.sd = sd
        self.wte = sd["transformer.wte.weight"].astype(F32)
        self.wpe = sd["transformer.wpe.weight"].astype(F32)
        self.bits = max(cfg["bit_widths"])
        self.linears: List[Dict[str, _LinearOracle]] = []
        for i in range(self.n_layer):
            p = f"transformer.h.{i}."
            self.linears.append({
                "c_attn": _LinearOracle(sd, p + "attn.c_attn.", cfg),
                "attn_proj": _LinearOracle(sd, p + "attn.c_proj.", cfg),
                "c_fc": _LinearOracle(sd, p + "mlp.c_fc.", cfg),
                "mlp_proj": _LinearOracle(sd, p + "mlp.c_proj.", cfg),
            })

    # -- state fan-out (p1/models_sp.py:224-246) --------------------------------
    def all_linears(self) -> Iterable[_LinearOracle]:
        for blk in self.linears:
            yield from blk.values()

    def set_precision(self, bits: int) -> int:
        if bits not in self.cfg["bit_widths"]:
            raise ValueError(f"Bit width {bits} not in configured widths {self.cfg['bit_widths']}")
        self.bits = bits
        for lin in self.all_linears():
            lin.current_bits = 32 if bits >= 32 else bits
        return bits

    def _ln(self, x, name):
        w = self.sd[f"{name}.weights.{self.bits}"].astype(F32)
        b = self.sd[f"{name}.biases.{self.bits}"].astype(F32)
        return switchable_layernorm_forward(x, w, b, self.eps)[0]

    # -- forward (p1/models_sp.py:58-76, 124-128, 160-171, 300-336, 421-439) -----
    def forward(self, input_ids: np.ndarray, return_hidden: bool = False, lm_head: bool = True):
        B, T = input_ids.shape
        C, H = self.n_embd, self.n_head
        hd = C // H
        h = (self.wte[input_ids] + self.wpe[np.arange(T)][None]).astype(F32)   # eval: dropout is identity
        hidden = []
        causal = np.tril(np.ones((T, T), dtype=bool))
        for i, blk in enumerate(self.linears):
            if return_hidden:
                hidden.append(h.copy())
            p = f"transformer.h.{i}"
            a = self._ln(h, p + ".ln_1")
            qkv = blk["c_attn"].forward(a)
            q, k, v = np.split(qkv, 3, axis=2)
            q = q.reshape(B, T, H, hd).transpose(0, 2, 1, 3)
            k = k.reshape(B, T, H, hd).transpose(0, 2, 1, 3)
            v = v.reshape(B, T, H, hd).transpose(0, 2, 1, 3)
            att = _mm(q, k.transpose(0, 1, 3, 2)) / F32(np.sqrt(hd))
            att = np.where(causal[None, None], att, F32(-np.inf))
            att = att - att.max(axis=-1, keepdims=True)
            att = np.exp(att)
            att = (att / att.sum(axis=-1, keepdims=True)).astype(F32)
            o = _mm(att, v).transpose(0, 2, 1, 3).reshape(B, T, C)
            h = h + blk["attn_proj"].forward(o)
            m = self._ln(h, p + ".ln_2")
            m = blk["c_fc"].forward(m)
            m = (F32(0.5) * m * (F32(1) + _erf(m * F32(0.7071067811865476)))).astype(F32)  # nn.GELU() erf form
            h = (h + blk["mlp_proj"].forward(m)).astype(F32)
        h = self._ln(h, "transformer.ln_f")
        if return_hidden:
            hidden.append(h.copy())
        if not lm_head:
            return h
        logits = _mm(h, self.wte.T)          # tied, unquantised LM head (p1/models_sp.py:396-398)
        return (logits, hidden) if return_hidden else logits

    # -- calibration (p1/train_sp.py:47-163) ------------------------------------
    def calibrate_weights(self, bits: int):
        for lin in self.all_linears():
            q = lin.q_w[bits]
            q.start_calibration()
            fake_quantize(q, lin.weight)
            finish_calibration(q)

    def calibrate_lora(self, bits: int):
        for lin in self.all_linears():
            lo = lin.lora.get(bits)
            if lo is None:
                continue
            for qk, wk in (("q_A", "A"), ("q_B", "B")):
                lo[qk].start_calibration()
                fake_quantize(lo[qk], lo[wk])
                finish_calibration(lo[qk])

    def calibrate_inputs(self, bits: int, batches: Iterable[np.ndarray]):
        for lin in self.all_linears():
            lin.q_in[bits].start_calibration()
            lin.calibration_mode = True
        for ids in batches:
            self.forward(ids, lm_head=False)     # statistics only need the transformer body
        for lin in self.all_linears():
            lin.calibration_mode = False
            finish_calibration(lin.q_in[bits])

    def calibrate(self, bits: int, batches: Iterable[np.ndarray]):
        self.set_precision(bits)
        self.calibrate_weights(bits)
        self.calibrate_lora(bits)
        self.calibrate_inputs(bits, batches)


def random_state_dict(cfg: dict, vocab_size: int, n_positions: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random-init parameters with the reference's state_dict key names and init scales
    (nn.Linear-like N(0, 0.02) weights, LoRA A ~ U(-1/sqrt(K), 1/sqrt(K)), LoRA B ~ N(0, 0.02) so the
    LoRA branch is non-trivial, LayerNorm weight 1 / bias 0).  Used by bench.py's CPU baseline."""
    rng = np.random.default_rng(seed)
    C = cfg["n_embd"]
    sd = {"transformer.wte.weight": (0.02 * rng.standard_normal((vocab_size, C))).astype(F32),
          "transformer.wpe.weight": (0.01 * rng.standard_normal((n_positions, C))).astype(F32)}

    def ln(prefix):
        for b in cfg["bit_widths"]:
            sd[f"{prefix}.weights.{b}"] = np.ones(C, F32)
            sd[f"{prefix}.biases.{b}"] = np.zeros(C, F32)

    def lin(prefix, k, n):
        sd[prefix + "linear.weight"] = (0.02 * rng.standard_normal((n, k))).astype(F32)
        sd[prefix + "linear.bias"] = np.zeros(n, F32)
        for b in cfg["bit_widths"]:
            r = cfg["lora_rank_per_bit"].get(b, 0)
            if b < 32 and r > 0:
                lim = 1.0 / np.sqrt(k)
                sd[f"{prefix}lora_adapters.{b}bit.lora_A"] = rng.uniform(-lim, lim, (k, r)).astype(F32)
                sd[f"{prefix}lora_adapters.{b}bit.lora_B"] = (0.02 * rng.standard_normal((r, n))).astype(F32)

    for i in range(cfg["n_layer"]):
        p = f"transformer.h.{i}."
        ln(p + "ln_1"); ln(p + "ln_2")
        lin(p + "attn.c_attn.", C, 3 * C); lin(p + "attn.c_proj.", C, C)
        lin(p + "mlp.c_fc.", C, 4 * C); lin(p + "mlp.c_proj.", 4 * C, C)
    ln("transformer.ln_f")
    return sd


def cross_entropy_shifted(logits: np.ndarray, labels: np.ndarray) -> float:
    """Mean next-token cross entropy, as SPLMHeadModel.forward with labels (p1/models_sp.py:441-449)."""
    lg = logits[:, :-1, :].astype(np.float64)
    tgt = labels[:, 1:]
    lg = lg - lg.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(lg).sum(axis=-1))
    picked = np.take_along_axis(lg, tgt[..., None], axis=-1)[..., 0]
    return float((lse - picked).mean())
