"""BASELINE.json configs[4]: SPLinearWithLoRA forward microbench over GPT-2 XL shapes (n_embd 1600) at 4-bit
(min-max) and 8-bit (log), 4K-64K tokens, against the tensor-pipe and HBM rooflines.
    python tools/qlinear_sweep.py > profiles/<round>_qlinear_sweep.csv"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib
from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
PEAK_TF = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1395.2)))
PEAK_GB = float(peaks.get("hbm_gbs", 6534.8))
dev = torch.device("cuda")
SHAPES = [("c_attn", 1600, 4800), ("c_fc", 1600, 6400), ("c_proj_mlp", 6400, 1600), ("c_proj_attn", 1600, 1600)]
RANK = 64


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3        # us


print("# SPLinearWithLoRA.forward (no_grad, LoRA rank 64 on): quantise activations + LoRA down-projection + fused GEMM")
print(f"# peaks: tensor {PEAK_TF} TFLOP/s (measured sustained bf16), HBM {PEAK_GB} GB/s (measured copy)")
print("layer,K,N,bits,quantizer,tokens,forward_us,gemm_us,gemm_TFLOPs,gemm_frac_of_tensor_peak,quantize_us,quantize_GBs,quantize_frac_of_hbm_peak,forward_Mtokens_per_s")
for name, K, N in SHAPES:
    for bits, qt in ((4, "minmax"), (8, "log")):
        torch.manual_seed(0)
        m = SPLinearWithLoRA(K, N, [4, 8, 32], {4: RANK, 8: RANK, 32: 0}, {4: 2 * RANK, 8: 2 * RANK, 32: 0},
                             {4: "minmax", 8: "log", 32: None}).to(dev)
        m.set_precision(bits)
        key = f"{bits}bit"
        with torch.no_grad():
            m.lora_adapters[key].lora_B.normal_(0, 0.02)
            q = m.quantizers_weight[key]; q.start_calibration(); q(m.linear.weight.data); q.finish_calibration()
            lo = m.lora_adapters[key]
            for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                qq.start_calibration(); qq(w.data); qq.finish_calibration()
        for tokens in (4096, 16384, 65536):
            x = torch.randn(tokens, K, device=dev)
            x[:, 7] *= 25.0
            with torch.no_grad():
                m.calibration_mode = True
                iq = m.quantizers_input[key]; iq.start_calibration(); m(x); iq.finish_calibration()
                m.calibration_mode = False
                reps = 20 if tokens <= 16384 else 8
                fwd = timed(lambda: m(x), reps)
                base, lora = m._operands_for(bits, True)
                act = base['act']
                a_q = torch.empty(tokens, K, device=dev, dtype=torch.float16); a_raw = torch.empty_like(a_q)
                qus = timed(lambda: _lib.quantize_act(x, act['scale'], act['zp'], act['bcast'], act['qtype'], act['bits'],
                                                      act['symmetric'], act['kind'], act['col_mul'], act['mul'], a_q, a_raw,
                                                      act['raw_mul']), reps)
                t16 = torch.zeros(tokens, RANK, device=dev, dtype=torch.float16)
                y = torch.empty(tokens, N, device=dev)
                gus = timed(lambda: _lib.qgemm(a_q, base['B_op'], tokens, N, K, y, A2=t16, B2=lora['Bl_op'], K2=RANK,
                                               col_scale=base['pw'], bias=m.linear.bias.detach()), reps)
            flops = 2.0 * tokens * N * (K + RANK)
            qbytes = tokens * K * 8.0
            print(f"{name},{K},{N},{bits},{qt},{tokens},{fwd:.1f},{gus:.1f},{flops / gus / 1e6:.0f},{flops / gus / 1e6 / PEAK_TF:.3f},"
                  f"{qus:.1f},{qbytes / qus / 1e3:.0f},{qbytes / qus / 1e3 / PEAK_GB:.3f},{tokens / fwd:.2f}")
        del m
        torch.cuda.empty_cache()
