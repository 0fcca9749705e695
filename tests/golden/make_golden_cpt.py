#!/usr/bin/env python
"""Golden fixtures for the CPT (part2) variant, produced by the UNMODIFIED reference
(part2_cyclic_precision_training.{quantization, cpt_model}) on CPU float32.  Separate from
make_golden.py because part2 uses bare module names (`quantization`, ...) that clash with part1's.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cpt.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPQ_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "part2_cyclic_precision_training"))
sys.dont_write_bytecode = True

from cpt_model import CPTLinear  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def make_input(shape, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * torch.exp(1.2 * torch.randn(*shape, generator=g)) * 0.5
    x[..., 3] *= 20.0
    flat = x.view(-1)
    flat[::53] = 0.0
    return x.contiguous()


def calibrate(m, bits, x_batches):
    """p2/calibration.py:17-88, 118-152: weight -> inputs (LoRA off) -> LoRA quantiser on A then B."""
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


def main():
    torch.manual_seed(0)
    K, N, r = 64, 96, 8
    m = CPTLinear(K, N, bit_widths=[4, 8, 32], quantizer_per_bit={4: "minmax", 8: "log", 32: None},
                  gradient_bits=8, shared_lora_rank=r, shared_lora_alpha=16)
    with torch.no_grad():
        m.shared_lora.lora_B.normal_(0, 0.05)
        m.linear.bias.normal_(0, 0.1)
    xc = [make_input((2, 24, K), 10 + i) for i in range(2)]
    d = {"x_calib": torch.stack(xc).numpy(), "weight": m.linear.weight.detach().numpy(),
         "bias": m.linear.bias.detach().numpy(), "lora_A": m.shared_lora.lora_A.detach().numpy(),
         "lora_B": m.shared_lora.lora_B.detach().numpy(), "scaling": np.float32(m.shared_lora.scaling),
         "meta": np.array([K, N, r])}
    x = make_input((2, 24, K), 50)
    gy = torch.randn(2, 24, N, generator=torch.Generator().manual_seed(7)) * 0.1
    gy.view(-1)[:3] = torch.tensor([25.0, -40.0, 11.0])
    d["x"], d["grad_y"] = x.numpy(), gy.numpy()
    for p in m.parameters():
        p.requires_grad_(True)
    m.train()
    for bits in (8, 4):
        calibrate(m, bits, xc)
        m.zero_grad()
        xg = x.clone().requires_grad_(True)
        y = m(xg)
        y.backward(gy)
        d[f"y{bits}"] = y.detach().numpy()
        d[f"gx{bits}"] = xg.grad.numpy().copy()
        d[f"gw{bits}"] = m.linear.weight.grad.numpy().copy()
        d[f"gb{bits}"] = m.linear.bias.grad.numpy().copy()
        d[f"gA{bits}"] = m.shared_lora.lora_A.grad.numpy().copy()
        d[f"gB{bits}"] = m.shared_lora.lora_B.grad.numpy().copy()
        for nm, q in (("qw", m.quantizer_weight), ("qi", m.quantizer_input), ("lq", m.lora_weight_quantizers[f"{bits}bit"])):
            d[f"{nm}{bits}_scale"] = q.scales[bits].numpy().copy()
            d[f"{nm}{bits}_zp"] = q.zero_points[bits].numpy().copy()
    # both widths stay calibrated when switching back (no recalibration)
    m.set_precision(8)
    with torch.no_grad():
        d["y8_again"] = m(x).numpy()
    # gradient quantisers: collect on one backward, then quantise the next one (p2/quantization.py:14-26)
    for gq in (m.shared_lora.grad_quantizer_A, m.shared_lora.grad_quantizer_B):
        gq.start_calibration()
    m.zero_grad()
    m(x.clone()).backward(gy)
    for gq in (m.shared_lora.grad_quantizer_A, m.shared_lora.grad_quantizer_B):
        gq.finish_calibration()
    m.zero_grad()
    m(x.clone()).backward(gy)
    d["gA8_gq"] = m.shared_lora.lora_A.grad.numpy().copy()
    d["gB8_gq"] = m.shared_lora.lora_B.grad.numpy().copy()
    d["gqA_scale"] = m.shared_lora.grad_quantizer_A.scales[8].numpy().copy()
    # state_dict keys of the multi-bit quantiser
    keys = sorted(m.state_dict().keys())
    d["state_keys"] = np.array(keys)
    with torch.no_grad():
        m.set_precision(32)
        d["y32"] = m(x).numpy()
    np.savez_compressed(os.path.join(OUT, "cpt_linear.npz"), **d)
    print("cpt_linear:", len(d), "arrays;", len(keys), "state keys")


if __name__ == "__main__":
    with contextlib.redirect_stderr(io.StringIO()):
        main()
