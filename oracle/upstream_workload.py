"""bench.py's CPU arm: BASELINE.json configs[1] driven through the UNMODIFIED upstream classes on torch-CPU.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/upstream.py).  The step is the one the GPU arm times, written
against upstream's own public API:

    for q in input quantisers: q.start_calibration()             p1/train_sp.py:91-94
    model.disable_lora_for_calibration(); model(ids)              p1/train_sp.py:98-110   (statistics pass)
    model.enable_lora_after_calibration(); q.finish_calibration() p1/train_sp.py:112-118
    model(ids, labels=ids)['loss']                                p1/models_sp.py:421-449 (quantised forward + CE)

fp32, autocast off, no_grad, torch.set_num_threads(all host cores).
"""
from __future__ import annotations

import os
import time

from . import upstream


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class UpstreamForwardWorkload:
    def __init__(self, bits: int, model_kw: dict, bit_widths, quantizer_per_bit, rank: int, sample_batch: int, seq: int,
                 seed: int = 0):
        import torch
        torch.set_num_threads(host_threads())
        self.torch = torch
        cfg = upstream.gpt2_config(n_layer=model_kw["n_layer"], n_embd=model_kw["n_embd"], n_head=model_kw["n_head"],
                                   bit_widths=bit_widths, quantizer_per_bit=quantizer_per_bit, rank=rank, alpha=rank,
                                   vocab_size=model_kw["vocab_size"], n_positions=model_kw["n_positions"])
        torch.manual_seed(seed)
        with upstream.quiet():
            self.model = upstream.p1("models_sp").SPLMHeadModel(cfg).eval()
        with torch.no_grad():
            self.model.transformer.wte.weight.normal_(0, 0.02)
            self.model.transformer.wpe.weight.normal_(0, 0.01)
            for n, p in self.model.named_parameters():
                if n.endswith("lora_B"):
                    p.normal_(0, 0.02)
        self.bits, self.key = bits, f"{bits}bit"
        self.B, self.T, self.V = sample_batch, seq, model_kw["vocab_size"]
        self.gen = torch.Generator().manual_seed(1234)
        with upstream.quiet(), torch.no_grad():
            self.model.set_precision(bits)
            self.linears = [m for m in self.model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
            for m in self.linears:                                            # static calibration (untimed)
                qw = m.quantizers_weight[self.key]
                qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
                lo = m.lora_adapters[self.key]
                for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                    qq.start_calibration(); qq(w.data); qq.finish_calibration()
        self.input_q = [m.quantizers_input[self.key] for m in self.linears]

    def step(self) -> float:
        torch = self.torch
        ids = torch.randint(0, self.V, (self.B, self.T), generator=self.gen)
        with upstream.quiet(), torch.no_grad():
            for q in self.input_q:
                q.start_calibration()
            self.model.disable_lora_for_calibration()
            self.model.transformer(ids)
            self.model.enable_lora_after_calibration()
            for q in self.input_q:
                q.finish_calibration()
            out = self.model(ids, labels=ids)
        return float(out["loss"])

    def run(self, steps: int, warmup: int):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        loss = None
        for _ in range(steps):
            loss = self.step()
        dt = time.perf_counter() - t0
        return dt, loss

    def describe(self, steps: int, dt: float) -> str:
        torch = self.torch
        return (f"{steps} step(s) of calibration pass + {self.bits}-bit forward + CE on {self.B} x {self.T} tokens, "
                f"unmodified upstream SPLMHeadModel on torch-CPU {torch.__version__} fp32, "
                f"torch threads {torch.get_num_threads()} of os.cpu_count() {os.cpu_count()}, {dt:.1f} s")
