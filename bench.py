#!/usr/bin/env python
"""Benchmark of the switchable-precision fake-quant linear path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): GPT-2 small SPLMHeadModel, random-init weights, synthetic
tokens, 8-bit log quantisation, per-GPU batch 32 x seq 1024.  One step = the calibration pass of
the 48 input quantisers (start_calibration -> forward in collecting mode with LoRA off ->
[MIN/MAX all-reduce across ranks] -> finish_calibration) followed by one quantised forward with
the LoRA branch on and the next-token cross-entropy.  Tokens are counted once per step.

Prints ONE JSON line (rank 0).  `value` = tokens/s with the token ids resident in HBM; `e2e` =
the same through the public module API with the ids in pinned host memory and the loss read back
to the host every step.  `--impl reference` times the CPU oracle (numpy restatement of the
reference's PyTorch path, oracle/) on a bounded sample of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GPT-2 SP tokens/s at 8-bit (calibration pass + quantised forward)"
UNIT = "tokens/s"
BITS = 8
MODEL = dict(vocab_size=50257, n_positions=1024, n_embd=768, n_layer=12, n_head=12, layer_norm_epsilon=1e-5)
BIT_WIDTHS = [4, 8, 32]
QUANTIZER_PER_BIT = {4: "minmax", 8: "log", 32: None}          # p1/config_sp.py:14-30
LORA_RANK = {4: 64, 8: 64, 32: 0}                              # p1/config_sp.py:36-37
LORA_ALPHA = {4: 64, 8: 64, 32: 0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bits", type=int, default=8, choices=[4, 8],
                    help="precision of the measured step: 8 = log quantisers (the headline configuration), 4 = min-max")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (weak scaling)")
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--cpu-sample-batch", type=int, default=2,
                    help="sequences (of --seq tokens) per step in the bounded CPU-arm sample; the GPU arm runs --batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=20,
                    help="optimizer steps of upstream's train_step (8 micro-steps each) timed after the main metric (0 = skip)")
    ap.add_argument("--train-dropout", type=float, default=0.1, help="embd_pdrop of the training step (p1/config_sp.py:9)")
    ap.add_argument("--train-eager", action="store_true", help="run the micro-steps eagerly (no CUDA graphs): per-phase timing")
    ap.add_argument("--train-strong", type=int, default=1, help="also time the strong-scaling shapes (global 32x256, 32x1024)")
    ap.add_argument("--no-dp-parity", action="store_true")
    ap.add_argument("--cpt-steps", type=int, default=10, help="timed steps of the GPT-2 medium CPT section (configs[3]; 0 = skip)")
    ap.add_argument("--sweep-tokens", type=int, nargs="*", default=[4096, 8192, 16384, 32768, 65536],
                    help="token counts of the QLinear sweep (configs[4]; none = skip; runs on rank 0 at N = 1 only)")
    ap.add_argument("--train-batch", type=int, default=32, help="per-GPU sequences of the training step (p1/config_sp.py:46)")
    ap.add_argument("--train-seq", type=int, default=256, help="sequence length of the training step (p1/config_sp.py:47)")
    ap.add_argument("--profile-train-step", action="store_true",
                    help="for ncu launch lists: run the training section with one NVTX-ranged step")
    ap.add_argument("--eager-step", action="store_true",
                    help="run the headline step eagerly (one launch per kernel) instead of as CUDA-graph replays")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="for ncu launch lists: warm up, run exactly one un-instrumented step, print nothing else")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def measured_traffic():
    """DRAM bytes per qgemm_nt launch from the committed ncu capture of one step (profiles/), or None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02c_qgemm_dram_per_launch.csv")
    try:
        with open(path) as f:
            for line in f:
                if line.startswith("TOTAL("):
                    n = int(line[len("TOTAL("):].split(" ")[0])
                    _, _, rd, wr = line.strip().split(",")
                    return (float(rd) + float(wr)) * 1e6 / n, "profiles/r02c_qgemm_dram_per_launch.csv (ncu, one step)"
    except Exception:
        pass
    return None, "no capture committed"


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the numpy oracle on a bounded sample
# ------------------------------------------------------------------------------------------

def oracle_cfg():
    return dict(n_layer=MODEL["n_layer"], n_head=MODEL["n_head"], n_embd=MODEL["n_embd"],
                layer_norm_epsilon=MODEL["layer_norm_epsilon"], bit_widths=BIT_WIDTHS,
                quantizer_per_bit=QUANTIZER_PER_BIT, lora_rank_per_bit=LORA_RANK, lora_alpha_per_bit=LORA_ALPHA,
                per_channel=True)


class CpuWorkload:
    """Same step as the GPU arm, restated on the CPU oracle."""

    def __init__(self, sample_batch, seq):
        import numpy as np
        from oracle.model_oracle import SPModelOracle, random_state_dict
        self.np = np
        cfg = oracle_cfg()
        self.model = SPModelOracle(cfg, random_state_dict(cfg, MODEL["vocab_size"], MODEL["n_positions"], seed=0))
        self.model.set_precision(BITS)
        self.model.calibrate_weights(BITS)
        self.model.calibrate_lora(BITS)
        self.rng = np.random.default_rng(1)
        self.B, self.T = sample_batch, seq

    def step(self):
        from oracle.model_oracle import cross_entropy_shifted
        ids = self.rng.integers(0, MODEL["vocab_size"], (self.B, self.T))
        self.model.calibrate_inputs(BITS, [ids])
        logits = self.model.forward(ids)
        return cross_entropy_shifted(logits, ids)


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_baseline(sample_batch, seq, steps=1, warmup=0):
    """The CPU arm: the UNMODIFIED upstream SPLMHeadModel on torch-CPU (baseline/_ref, `kind: reference`) on a bounded
    sample of the workload; the numpy port (`kind: port`, ~4x slower than upstream) only when that copy is absent."""
    from oracle import upstream
    if upstream.available() and not os.environ.get("SPQ_CPU_PORT"):
        from oracle.upstream_workload import UpstreamForwardWorkload, host_threads
        w = UpstreamForwardWorkload(BITS, MODEL, BIT_WIDTHS, QUANTIZER_PER_BIT, LORA_RANK[BITS], sample_batch, seq)
        dt, loss = w.run(steps, warmup)
        import torch
        return {"value": sample_batch * seq * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(),
                "os_cpu_count": os.cpu_count(), "kind": "reference", "sample": w.describe(steps, dt)}, dt / max(steps, 1), loss
    w = CpuWorkload(sample_batch, seq)
    for _ in range(warmup):
        w.step()
    t0 = time.perf_counter()
    loss = None
    for _ in range(steps):
        loss = w.step()
    dt = time.perf_counter() - t0
    return {"value": sample_batch * seq * steps / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": f"{steps} step(s) of calibration pass + forward on {sample_batch} x {seq} tokens "
                      f"(numpy oracle -- baseline/_ref absent, BLAS threads = host cores), {dt:.1f} s"}, dt / max(steps, 1), loss


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (torch / numpy / BLAS read
    # these at import time, and nothing numeric has been imported yet in this process)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cpu_threads())
    base, sec_per_step, loss = run_cpu_baseline(args.cpu_sample_batch, args.seq, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": base["sample"],
                   "sample_batch": args.cpu_sample_batch, "seq_len": args.seq},
        "cpu_baseline": base, "loss": loss,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"GPT-2 small (124M) SPLMHeadModel, {BITS}-bit {QUANTIZER_PER_BIT[BITS]} per-channel: calibration pass + quantised forward + CE, "
            f"batch {args.batch} x seq {args.seq} per GPU, LoRA rank 64, random-init weights")


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region: NVML (the source nvidia-smi reads) polled from
    a thread every 100 ms (`nvidia-smi -lms 100` as a child process is the fallback).  Either way the queries
    perturb the device now and then: A/B runs with the sampler off (SPQ_NOCLOCKS=1) never show the single
    +20..80 ms step that about one end-to-end region in two shows with it on (the queries themselves return in
    ~1 ms; `max_query_ms` reports that).  The resident-input loop hides it behind its launch queue."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                    (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()
        self.query_ms = (0.0, 0.0)

    def start(self):
        try:
            # default: the recipe's own sampler, `nvidia-smi -lms 200` in a CHILD process (no driver calls from a second
            # thread of the measuring process: the in-process NVML poll coincided with single +50 ms steps in the
            # end-to-end region of some runs); SPQ_CLOCKS=nvml selects the in-process poll
            if os.environ.get("SPQ_CLOCKS", "smi") != "nvml":
                raise RuntimeError("nvidia-smi child requested")
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[0].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[0].isdigit() else self.index
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(phys), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                q0 = time.perf_counter()
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                q1 = time.perf_counter()
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                q2 = time.perf_counter()
                self.query_ms = (max(self.query_ms[0], (q1 - q0) * 1e3), max(self.query_ms[1], (q2 - q1) * 1e3))
                flags = ["Active" if mask & bit else "Not Active" for bit, _ in self.NVML_REASONS]
                self.rows.append(", ".join([str(sm), str(mx), "0"] + flags))
            except Exception:
                pass
            self._stop.wait(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml (100 ms poll)" if self.nvml is not None else "nvidia-smi -lms 200 (child process)",
                "max_query_ms": [round(v, 2) for v in self.query_ms]}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from transformers import GPT2Config
    from llm_qat_on_gpt2_b200 import SPLMHeadModel, _lib
    from llm_qat_on_gpt2_b200 import dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path for the product)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load_library()

    cfg = GPT2Config(**MODEL, embd_pdrop=0.0)
    cfg.bit_widths = BIT_WIDTHS
    cfg.lora_rank_per_bit = LORA_RANK
    cfg.lora_alpha_per_bit = LORA_ALPHA
    cfg.quantizer_per_bit = QUANTIZER_PER_BIT
    cfg.per_channel_quantization = True
    cfg.attention_dtype = "fp16"       # stock torch SDPA (flash) between the hot-path linears
    # SPQ_MLP_ACT=fp16: gelu(c_fc(.)) held in float16 between the two MLP linears, as under upstream's autocast
    # (p1/train_sp.py:319).  Measured -0.3..-0.7 ms per step (profiles/r02b_*): not worth leaving upstream's fp32
    # semantics for, so the headline keeps float32 there
    cfg.mlp_activation_dtype = os.environ.get("SPQ_MLP_ACT", "fp32")
    torch.manual_seed(0)               # identical replicas on every rank
    model = SPLMHeadModel(cfg).to(dev).eval()
    with torch.no_grad():
        # GPT-2's initializer_range (0.02) for the embeddings / tied LM head: torch's default N(0,1) gives
        # logits of magnitude ~sqrt(768) * 30, a one-hot softmax and vanishing distillation gradients
        model.transformer.wte.weight.normal_(0, 0.02)
        model.transformer.wpe.weight.normal_(0, 0.01)
        for n, p in model.named_parameters():
            if n.endswith("lora_B"):
                p.normal_(0, 0.02)     # non-trivial LoRA branch
    model.set_precision(BITS)
    key = f"{BITS}bit"
    linears = [m for m in model.modules() if m.__class__.__name__ == "SPLinearWithLoRA"]
    # static calibration (untimed): weight and LoRA quantisers, as p1/train_sp.py:58-83, 125-163
    with torch.no_grad():
        for m in linears:
            qw = m.quantizers_weight[key]
            qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
            lo = m.lora_adapters[key]
            for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                qq.start_calibration(); qq(w.data); qq.finish_calibration()
    input_q = [m.quantizers_input[key] for m in linears]
    group = dist.group.WORLD if world > 1 else None

    B, T = args.batch, args.seq
    gen = torch.Generator().manual_seed(1234 + rank)
    n_batches = args.steps + args.warmup
    host_ids = [torch.randint(0, MODEL["vocab_size"], (B, T), generator=gen).pin_memory() for _ in range(n_batches)]
    dev_ids = [h.to(dev) for h in host_ids]

    def step_eager(ids):
        with torch.no_grad():
            for q in input_q:
                q.start_calibration()
            model.disable_lora_for_calibration()
            model.transformer(ids)                       # statistics only need the transformer body
            model.enable_lora_after_calibration()
            dp.finish_calibration_many(input_q, group)   # one MIN/MAX exchange + one flag read
            out = model(ids, labels=ids)
        return out["loss"]

    # the same step as two CUDA-graph replays (statistics pass | finish_calibration + operand rebuild + quantised
    # forward + loss) with the statistics exchange in between and the flag read deferred: `--eager-step` is the A/B
    # switch; the roofline pass (an event pair around every GEMM launch) and --profile-one-step run it eagerly
    from llm_qat_on_gpt2_b200.training import GraphedCalibratedForward
    graphed = None if (args.eager_step or args.profile_one_step) else GraphedCalibratedForward(
        model, group, n_side=int(os.environ.get("SPQ_STEP_SIDE_STREAMS", "8")),
        overlap_stats=os.environ.get("SPQ_STEP_OVERLAP_STATS", "1") != "0")

    def step(ids):
        if graphed is None:
            return step_eager(ids.to(dev, non_blocking=True) if not ids.is_cuda else ids)
        return graphed(ids)["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- instrument the dominant kernel (spq_qgemm) with CUDA events on the launching stream
    gemm_events = []
    orig_qgemm = _lib.qgemm

    def timed_qgemm(A, Bm, M, N, K, out, A2=None, B2=None, K2=0, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_qgemm(A, Bm, M, N, K, out, A2=A2, B2=B2, K2=K2, **kw)
        e1.record()
        # algorithmic bytes of the launch: operands once, output once, residual once
        nbytes = 2.0 * (M + N) * (K + K2) + M * N * out.element_size() + (4.0 * M * N if kw.get("C") is not None else 0.0)
        gemm_events.append((e0, e1, 2.0 * M * N * (K + K2), nbytes))
        return r

    orig_qgemm_lse = _lib.qgemm_lse

    def timed_qgemm_lse(A, Bm, M, N, K, out, **kw):          # the LM head (same kernel, log-sum-exp epilogue)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_qgemm_lse(A, Bm, M, N, K, out, **kw)
        e1.record()
        gemm_events.append((e0, e1, 2.0 * M * N * K, 2.0 * (M + N) * K + 4.0 * M * N + 8.0 * M * r.shape[1]))
        return r

    def patch(fn):
        import llm_qat_on_gpt2_b200.lora as lora_mod
        lora_mod._lib.qgemm = fn
        lora_mod._lib.qgemm_lse = timed_qgemm_lse if fn is timed_qgemm else orig_qgemm_lse

    # the clock sampler attaches to NVML before the warm-up (its one-time driver work stays out of the timed
    # regions); its samples are reset when the first timed region starts
    sampler = ClockSampler(local)
    if rank == 0 and not (args.profile_train_step or args.profile_one_step) and not os.environ.get('SPQ_NOCLOCKS'):
        sampler.start()
    for i in range(args.warmup):
        step(dev_ids[i])
    barrier()
    # a one-off ~50 ms device stall shows up ~0.4 s after sustained load begins (power management settling; with
    # W = 3 it landed in the third end-to-end step of every other run): keep the load up, untimed, for 24 steps
    # in total (~0.9 s; a fixed count so that all ranks issue the same collectives)
    extra_warmup = 0 if (args.profile_train_step or args.profile_one_step) else max(0, 24 - args.warmup)
    for i in range(extra_warmup):
        step(dev_ids[i % len(dev_ids)])
    barrier()
    if args.profile_train_step:
        train_section(args, model, linears, key, dev, world, rank, group, barrier)
        print(json.dumps({"profile_train_step": True}), flush=True)
        return
    if args.profile_one_step:
        n0 = _lib.launch_count()
        torch.cuda.nvtx.range_push("spq_step")
        step(dev_ids[args.warmup])
        torch.cuda.nvtx.range_pop()
        barrier()
        print(json.dumps({"profile_one_step": True, "spq_launches_in_step": _lib.launch_count() - n0}), flush=True)
        return

    # the cyclic GC is parked during the timed regions: a generation-2 pass over the module graph showed up as
    # a single +25..60 ms step in the end-to-end loop, where the host cannot run ahead of the device
    import gc
    gc.collect()
    gc.disable()
    # ---- timed region 1: inputs resident in HBM ("value"), nothing instrumented
    sampler.rows.clear()
    launches0 = _lib.launch_count()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        loss = step(dev_ids[args.warmup + i])
    t1.record()
    barrier()
    ms_value = t0.elapsed_time(t1)
    launches = _lib.launch_count() - launches0
    if graphed is not None:
        # kernels of this library inside the replays (counted while they were captured) + eager launches, if any
        launches += graphed.kernels_per_replay * args.steps
    # ---- roofline pass (separate, untimed for `value`): the same steps with a CUDA-event pair around every GEMM launch
    patch(timed_qgemm)
    barrier()
    t0.record()
    for i in range(args.steps):
        step_eager(dev_ids[args.warmup + i])
    t1.record()
    barrier()
    ms_roofline_pass = t0.elapsed_time(t1)
    patch(orig_qgemm)
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in gemm_events)
    gemm_flops = sum(f for _, _, f, _ in gemm_events)
    gemm_bytes = sum(nb for _, _, _, nb in gemm_events)
    n_gemm = len(gemm_events)

    # ---- timed region 2: end to end through the module API, ids from pinned host memory, loss to host
    barrier()
    e2e_step_ms = []
    t0.record()
    for i in range(args.steps):
        w0 = time.perf_counter()
        # pinned host ids -> device (asynchronous copy into the step's input buffer), loss read back
        loss_host = float(step(host_ids[args.warmup + i]).item())
        e2e_step_ms.append(round((time.perf_counter() - w0) * 1e3, 2))
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    gc.enable()

    # ---- untimed: the graph-replayed step against the eager module calls on the same ids -- calibrated scale /
    # zero-point of all 48 input quantisers and the loss bit for bit, and (N > 1) identical on every rank
    step_parity = None
    if graphed is not None:
        def snapshot():
            return torch.cat([torch.cat([q.scale.reshape(-1), q.zero_point.reshape(-1)]) for q in input_q])
        probe = dev_ids[0]
        loss_g = graphed(probe)["loss"].clone()
        snap_g = snapshot().clone()
        loss_e = step_eager(probe).clone()
        snap_e = snapshot()
        same = bool(torch.equal(snap_g.view(torch.int32), snap_e.view(torch.int32)) and torch.equal(loss_g.reshape(()), loss_e.reshape(())))
        ranks_same = True
        if world > 1:
            gathered = [torch.empty_like(snap_g) for _ in range(world)]
            dist.all_gather(gathered, snap_g)
            ranks_same = all(torch.equal(g.view(torch.int32), gathered[0].view(torch.int32)) for g in gathered)
        step_parity = {"graph_vs_eager_bit_identical": same, "ranks_bit_identical": ranks_same, "parameters": int(snap_g.numel())}
        if not (same and ranks_same):
            raise RuntimeError(f"graph-replayed step differs from the eager step: {step_parity}")

    train = None
    if args.train_steps > 0:
        train = train_section(args, model, linears, key, dev, world, rank, group, barrier)
    cpt = sweep = None
    if args.cpt_steps > 0 or (args.sweep_tokens and world == 1):
        # the GPT-2 small replica (and its graph pools) is not needed any more
        del model, linears, input_q
        gc.collect()
        torch.cuda.empty_cache()
        if args.cpt_steps > 0:
            cpt = cpt_section(args, dev, world, rank, group, barrier)
        if args.sweep_tokens and world == 1:
            sweep = qlinear_sweep_section(args, dev)
    nodata = graphed.nodata_count() if graphed is not None else 0
    wd = _lib.debug_status()
    if wd != 0:
        raise RuntimeError(f"GEMM pipeline watchdog flag {wd}: results of this run are invalid")

    if world > 1:
        t = torch.tensor([ms_value, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_value, ms_e2e = t.tolist()

    if rank == 0:
        peaks = measured_peaks()
        tokens = B * T * world * args.steps
        value = tokens / (ms_value / 1e3)
        e2e = tokens / (ms_e2e / 1e3)
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        peak = peaks["tflops_sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands (exact integer codes / dequantised values), f32 accumulate",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "per_gpu_batch": B, "seq_len": T, "bits": BITS,
                       "parallelism": f"dp{world} (replicas, batch-sharded; MIN/MAX all-reduce of calibration statistics)",
                       "l2": "per-step working set (~10 GB of activations + 6.6 GB of logits) >> 126 MB L2; fresh token ids every step",
                       "attention": "torch SDPA fp16 (outside the hot path)",
                       "mlp_activation": f"gelu(c_fc) stored as {cfg.mlp_activation_dtype} between the MLP linears",
                       "extra_untimed_warmup_steps": extra_warmup,
                       "launch_mode": ("eager (one launch per kernel, one device->host flag read per step)" if graphed is None else
                                       f"2 CUDA-graph replays per step ({graphed.kernels_per_replay} kernels of this library "
                                       f"captured); quantiser calibrations without data: {nodata}")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "loss": loss_host, "step_wall_ms": e2e_step_ms},
            "gpu_launches": launches,
            "step_parity": step_parity,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "spq::gemm::qgemm_nt_kernel (all launches in the timed region)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if (achieved and peak) else None,
                         "traffic": measured_traffic()[0], "traffic_unit": "bytes per launch (dram__bytes_read.sum + "
                         "dram__bytes_write.sum averaged over every qgemm_nt launch of one step)",
                         "traffic_source": measured_traffic()[1],
                         "algorithmic_bytes_per_launch": gemm_bytes / max(n_gemm, 1),
                         "peak_source": peaks["source"] + ", bf16 dense sustained",
                         "launches": n_gemm, "share_of_step": gemm_ms / ms_roofline_pass if ms_roofline_pass else None,
                         "measured_in": "separate instrumented pass of the same steps (CUDA-event pair per launch), "
                                        f"{ms_roofline_pass / args.steps:.2f} ms per step"},
        }
        if train is not None:
            line["train"] = train
        if cpt is not None:
            line["cpt_medium"] = cpt
        if sweep is not None:
            line["qlinear_sweep"] = sweep
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = run_cpu_baseline(args.cpu_sample_batch, args.seq, steps=1, warmup=0)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


GRAD_ACCUM = 8                      # p1/config_sp.py:58
# GEMM work of one optimizer step of upstream's train_step, MFLOP per token of the batch:
#   teacher micro-step : CE forward (base 169.87 + LM head 77.19) + backward dX of both + the cache forward
#   student micro-step : forward (base + LoRA 18.87 + LM head) + backward (dX 169.87 + LoRA 37.75 + LM head dX)
TRAIN_MFLOP_TEACHER = 3 * (169.87 + 77.19)
TRAIN_MFLOP_STUDENT = (169.87 + 18.87 + 77.19) + (169.87 + 37.75 + 77.19)


def _calibrate_width(model, linears, bits, ids, group):
    """Untimed: weight, LoRA and input quantisers of one student width (p1/train_sp.py:47-163); the input statistics
    are MIN/MAX all-reduced across the ranks."""
    import torch
    from llm_qat_on_gpt2_b200 import dp
    key = f"{bits}bit"
    with torch.no_grad():
        model.set_precision(bits)
        for m in linears:
            qw = m.quantizers_weight[key]
            qw.start_calibration(); qw(m.linear.weight.data); qw.finish_calibration()
            lo = m.lora_adapters[key]
            for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                qq.start_calibration(); qq(w.data); qq.finish_calibration()
        iq = [m.quantizers_input[key] for m in linears]
        for q in iq:
            q.start_calibration()
        model.disable_lora_for_calibration()
        model.transformer(ids)
        model.enable_lora_after_calibration()
        dp.finish_calibration_many(iq, group)


def _timed_train(trainer, host_batches, dev, steps, warmup, world, barrier):
    """(ms per optimizer step resident, ms per step end to end, last result) -- max over ranks is taken by the caller."""
    import torch
    dev_batches = [h.to(dev) for h in host_batches]
    for i in range(warmup):
        res = trainer.train_step(dev_batches[i % len(dev_batches)])
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        res = trainer.train_step(dev_batches[(warmup + i) % len(dev_batches)])
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1) / steps
    t0.record()
    for i in range(steps):
        res = trainer.train_step(host_batches[(warmup + i) % len(host_batches)])       # pinned host ids -> device
    t1.record()
    barrier()
    return ms, t0.elapsed_time(t1) / steps, res


def train_section(args, model, linears, key, dev, world, rank, group, barrier):
    """BASELINE.json configs[2]: upstream's `train_step` (p1/train_sp.py:341-397) through training.SPTrainer -- per
    optimizer step 8 micro-steps on one batch: teacher (32-bit) CE forward + backward + cache forward, then 7 student
    micro-steps at random.choice([4, 8]) with LoRA recalibration, KL(T=3) + 1e-7 MSE, backward; gradient all-reduce
    of the touched LoRA / LayerNorm segments (overlapped), clip 1.0, AdamW.  tokens/s = B*T*8*world / step time
    (SURVEY section 8d).  Weak scaling: the reference's 32 x 256 per GPU; strong scaling: global 32 x 256 and
    32 x 1024 sharded over the ranks."""
    import random
    import torch
    import torch.distributed as dist
    from llm_qat_on_gpt2_b200 import _lib
    from llm_qat_on_gpt2_b200.training import SPTrainer
    V = MODEL["vocab_size"]
    other = [b for b in BIT_WIDTHS if b < 32 and b != BITS]
    gen = torch.Generator().manual_seed(99 + rank)
    calib_ids = torch.randint(0, V, (args.train_batch, args.train_seq), generator=gen).to(dev)
    for b in other:
        _calibrate_width(model, linears, b, calib_ids, group)
    model.train()
    model.transformer.drop.p = args.train_dropout            # p1/config_sp.py:9 embd_pdrop = 0.1
    out = {}

    def run(tag, B, T, steps, warmup, phases=False):
        trainer = SPTrainer(model, BIT_WIDTHS, grad_accum=GRAD_ACCUM, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0,
                            temperature=3.0, alpha_kl=1.0, alpha_feature=1e-7, total_lr_steps=550 * GRAD_ACCUM, group=group,
                            rng=random.Random(7), use_graphs=not args.train_eager)
        hosts = [torch.randint(0, V, (B, T), generator=gen).pin_memory() for _ in range(4)]
        if args.profile_train_step:
            # for ncu launch lists (--profile-from-start off): warm up / capture, then exactly one profiled step
            d0 = hosts[0].to(dev)
            for _ in range(3):
                trainer.train_step(d0)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            trainer.train_step(d0)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            return {"profiled": tag}
        l0 = _lib.launch_count()
        ms, ms_e2e, res = _timed_train(trainer, hosts, dev, steps, warmup, world, barrier)
        launches = (_lib.launch_count() - l0)
        ph = None
        if phases:
            trainer.phase_events = []
            trainer.train_step(hosts[0].to(dev))
            ph = trainer.phase_times_ms()
            trainer.phase_events = None
        if world > 1:
            tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, ms_e2e = tt.tolist()
        tokens = B * T * GRAD_ACCUM * world
        n_student = GRAD_ACCUM - 1
        flop = (TRAIN_MFLOP_TEACHER + n_student * TRAIN_MFLOP_STUDENT) * 1e6 * B * T       # per GPU per optimizer step
        peak = measured_peaks()["tflops_sustained"]
        r = {"value": tokens / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
             "per_gpu_batch": B, "seq_len": T, "micro_steps": GRAD_ACCUM, "loss": res["loss"], "precisions_last_step": res["precisions"],
             "e2e": {"value": tokens / (ms_e2e / 1e3), "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * T * 8,
                     "d2h_bytes_per_step": int(trainer.loss_buf.numel() * 4)},
             "trainable_params": int(trainer.state.numel), "cuda_graphs": sorted(trainer.graphs),
             # one graph replay + one loss copy (two for a student) per micro-step, the gradient memset, the
             # norm (2 kernels) and one AdamW launch per touched segment, one all-reduce per touched segment
             "host_launches_per_step": (GRAD_ACCUM * 3 - 1 + 1 + 2 + 3 + (3 if world > 1 else 0)) if trainer.graphs else None,
             "lib_launches_outside_graphs": launches,
             "roofline": {"bound": "tensor", "kernel": "all GEMMs of the optimizer step (algorithmic FLOP / whole step time)",
                          "achieved": flop / (ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                          "frac": flop / (ms / 1e3) / 1e12 / peak if peak else None}}
        if ph is not None:
            agg = {}
            for name, t in ph:
                agg[name] = agg.get(name, 0.0) + t
            r["phases_ms"] = {k: round(v, 3) for k, v in agg.items()}
        del trainer
        torch.cuda.empty_cache()
        return r

    B, T = args.train_batch, args.train_seq
    out = run("weak", B, T, args.train_steps, 3, phases=True)
    if args.profile_train_step:
        return out
    out["metric"] = ("GPT-2 SP train_step tokens/s (teacher CE fwd+bwd + cache fwd @32, 7 student micro-steps @random{4,8} with "
                     "LoRA recalibration + KL/MSE distillation fwd+bwd, grad all-reduce, clip, AdamW); tokens = B*T*8 per optimizer step")
    out["scaling"] = "weak"
    if args.train_strong and B % world == 0:
        for tag, gb, gt, st in (("strong_32x256", 32, 256, max(5, args.train_steps // 2)), ("strong_32x1024", 32, 1024, 5)):
            if gb % world == 0:
                r = run(tag, gb // world, gt, st, 2)
                r["global_batch"], r["scaling"] = gb, "strong"
                out[tag] = r
    if world > 1 and not args.no_dp_parity:
        out["dp_parity"] = dp_parity_section(args, model, linears, dev, world, rank, group)
    model.transformer.drop.p = 0.0
    model.eval()
    return out


def dp_parity_section(args, model, linears, dev, world, rank, group):
    """Untimed, N > 1: (1) the calibrated input-quantiser parameters are bit-identical on every rank and equal to a
    single-process calibration of the gathered global batch; (2) the all-reduced gradient of one optimizer step
    equals the gradient a single process computes on the global batch."""
    import random
    import torch
    import torch.distributed as dist
    from llm_qat_on_gpt2_b200.training import SPTrainer
    V, Bp, T = MODEL["vocab_size"], 4, args.train_seq
    p_drop, model.transformer.drop.p = model.transformer.drop.p, 0.0      # parity needs the same function on every rank
    gen = torch.Generator().manual_seed(4242 + rank)
    ids = torch.randint(0, V, (Bp, T), generator=gen).to(dev)
    gathered = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(gathered, ids)
    global_ids = torch.cat(gathered, dim=0)
    res = {}
    # ---- (1) calibration
    def params():
        return torch.cat([torch.cat([q.scale.reshape(-1), q.zero_point.reshape(-1)])
                          for m in linears for q in (m.quantizers_input[f"{BITS}bit"],)]).view(torch.int32).clone()
    _calibrate_width(model, linears, BITS, ids, group)                 # sharded, MIN/MAX all-reduced
    mine = params()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = torch.tensor([int(torch.equal(mine, ref0))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    _calibrate_width(model, linears, BITS, global_ids, None)           # single-process on the whole batch
    single = params()
    d = (single.long() - mine.long()).abs()
    res["calibration"] = {"ranks_bit_identical": bool(same.item()), "vs_single_process_max_ulp": int(d.max()),
                          "vs_single_process_differing": int((d > 0).sum()), "parameters": int(d.numel())}
    # ---- (2) gradients: lr = 0 keeps the replicas' parameters untouched.  The library attention between the hot-path
    # linears (cuDNN fp16 flash SDPA) is not batch-invariant: its results for a sequence depend on how many sequences
    # are in the call (measured: 2-4e-2 on these gradients between B = 4 and B = 8, tools/diag_dp.py), which would
    # mask what is being checked.  The check therefore runs the exact fp32 attention path; every kernel of this repo
    # is batch-invariant, so shard-average and global gradient then differ only by fp32 summation order.
    att_modules = [m for m in model.modules() if hasattr(m, "attention_dtype")]
    att_saved = [m.attention_dtype for m in att_modules]
    for m in att_modules:
        m.attention_dtype = "fp32"
    def grads(batch, data_parallel):
        tr = SPTrainer(model, BIT_WIDTHS, grad_accum=3, lr=0.0, weight_decay=0.0, group=group, rng=random.Random(3),
                       use_graphs=False)
        if not data_parallel:
            tr.world = 1
        tr.train_step(batch)
        g = tr.state.flat_grad.clone() / tr.world
        del tr
        return g
    g_dp = grads(ids, True)
    g_single = grads(global_ids, False)
    rel = float((g_dp.double() - g_single.double()).norm() / g_single.double().norm())
    model.transformer.drop.p = p_drop
    for m, a in zip(att_modules, att_saved):
        m.attention_dtype = a
    res["gradient"] = {"rel_allreduced_vs_single_process": rel, "elements": int(g_dp.numel()),
                       "per_rank_batch": Bp, "global_batch": Bp * world, "micro_steps": 3, "attention": "fp32 (batch-invariant)"}
    ok = res["calibration"]["ranks_bit_identical"] and res["calibration"]["vs_single_process_max_ulp"] == 0 and rel <= 1e-5
    res["ok"] = bool(ok)
    torch.cuda.empty_cache()
    return res


def cpt_section(args, dev, world, rank, group, barrier):
    """BASELINE.json configs[3]: CPTModel GPT-2 medium (n_embd 1024, 24 layers, 16 heads; shared LoRA r = 16, alpha = 32:
    p2/config_cpt.py:11-12), every width 3..8 pre-calibrated (p2/calibration.py), the width cycling 3 -> 8 -> 3 PER
    STEP along CyclicPrecisionScheduler's cosine (p2/cyclic_scheduler.py:24-43), one optimizer step per batch of
    32 x 256 per GPU (p2/main_cpt.py:45-60: CE loss, backward, clip 1.0, AdamW), batch-sharded data parallel."""
    import types
    import torch
    import torch.distributed as dist
    from llm_qat_on_gpt2_b200.cpt import CPTModel, CyclicPrecisionScheduler, CalibrationManager, CPTTrainer
    widths = [3, 4, 5, 6, 7, 8]
    mc = types.SimpleNamespace(vocab_size=50257, n_positions=1024, n_embd=1024, n_layer=24, n_head=16, layer_norm_epsilon=1e-5,
                               embd_pdrop=args.train_dropout, bit_widths=widths + [32], shared_lora_rank=16, shared_lora_alpha=32,
                               quantizer_per_bit={**{b: "log" for b in widths}, 32: None}, gradient_bits=8,
                               attention_dtype="fp16")        # library flash attention between the hot-path linears, as in the SP arm
    cfg = {"model": mc, "training": types.SimpleNamespace(target_bits=5)}
    torch.manual_seed(0)
    model = CPTModel(cfg).to(dev)
    with torch.no_grad():
        for m in model.modules():
            if m.__class__.__name__ == "LoRAAdapter" and m.lora_B is not None:
                m.lora_B.normal_(0, 0.02)                      # non-trivial adapter (upstream starts at zero)
    B, T, V = args.train_batch, args.train_seq, mc.vocab_size
    gen = torch.Generator().manual_seed(555 + rank)
    loader = [{"input_ids": torch.randint(0, V, (B, T), generator=gen)} for _ in range(2)]
    mgr = CalibrationManager(model, loader, dev, data_parallel_group=group)
    mgr.calibrate_gradient_quantizers()
    for b in widths:
        mgr.ensure_calibrated(b, num_batches=1)
    model.train()
    steps, warm = args.cpt_steps, 2 * len(widths)
    sched = CyclicPrecisionScheduler(bit_widths=widths, schedule_type="cosine", total_epochs=2 * (len(widths) - 1) * 10,
                                     total_cycles=10)            # 10 steps per cycle: 3 -> 8 -> 3
    trainer = CPTTrainer(model, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0, total_lr_steps=10000, group=group,
                         use_graphs=not args.train_eager)
    hosts = [torch.randint(0, V, (B, T), generator=gen).pin_memory() for _ in range(4)]
    devs = [h.to(dev) for h in hosts]
    for w in widths:                                            # capture every width's graph (untimed)
        trainer.train_step(devs[0], w, read_loss=False)
    seq = [sched.get_precision_for_epoch(i) for i in range(warm + 2 * steps)]
    for i in range(warm):
        trainer.train_step(devs[i % 4], seq[i], read_loss=False)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        r = trainer.train_step(devs[i % 4], seq[warm + i], read_loss=False)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1) / steps
    t0.record()
    for i in range(steps):
        r = trainer.train_step(hosts[i % 4], seq[warm + steps + i], read_loss=True)     # ids from pinned host, loss to host
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1) / steps
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
    tokens = B * T * world
    # base GEMMs 603.98 + quantised LM head 2*1024*50257 = 102.93 MFLOP/token per pass; forward + dX; LoRA r = 16 adds
    # 2*16*(K+N) per linear per pass (x3 with backward)
    lora = 2 * 16 * (24 * ((1024 + 3072) + (1024 + 1024) + (1024 + 4096) + (4096 + 1024)) + (1024 + 50257)) / 1e6
    flop = (2 * (603.98 + 102.93) + 3 * lora) * 1e6 * B * T
    peak = measured_peaks()["tflops_sustained"]
    out = {"metric": "GPT-2 medium CPT training tokens/s, width cycling 3->8 per step (p2 train_epoch_with_cpt step)",
           "value": tokens / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warm, "per_gpu_batch": B, "seq_len": T,
           "widths_timed": seq[warm:warm + steps], "loss": r["loss"], "scaling": "weak",
           "e2e": {"value": tokens / (ms_e2e / 1e3), "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4},
           "trainable_params": int(trainer.numel), "cuda_graphs": sorted(trainer.graphs),
           "roofline": {"bound": "tensor", "kernel": "all GEMMs of the step (algorithmic FLOP / whole step time)",
                        "achieved": flop / (ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flop / (ms / 1e3) / 1e12 / peak if peak else None}}
    del trainer, model
    torch.cuda.empty_cache()
    return out


def qlinear_sweep_section(args, dev):
    """BASELINE.json configs[4]: SPLinearWithLoRA over GPT-2 XL shapes (1600 -> 4800 / 6400, 6400 -> 1600), 4-bit min-max and
    8-bit log, 4K-64K tokens, forward (no_grad) and forward + STE backward (LoRA A/B and dX), LoRA rank 64; x ~ N(0,1)
    with x20 outlier channels, W ~ N(0, 0.02).  Reported per case: time, tokens/s, algorithmic
    TFLOP/s of the whole call against the measured bf16 peak (forward 2*M*N*K + LoRA 2*M*r*(K+N); backward adds dX
    2*M*N*K and 2x the LoRA work)."""
    import torch
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    peak = measured_peaks()["tflops_sustained"]
    R = 64
    rows = []

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
        return e0.elapsed_time(e1) / reps

    for K, N in ((1600, 4800), (1600, 6400), (6400, 1600)):
        for bits, qt in ((4, "minmax"), (8, "log")):
            torch.manual_seed(0)
            m = SPLinearWithLoRA(K, N, [4, 8, 32], {4: R, 8: R, 32: 0}, {4: R, 8: R, 32: 0}, {4: "minmax", 8: "log", 32: None}).to(dev)
            m.set_precision(bits)
            key = f"{bits}bit"
            lo = m.lora_adapters[key]
            with torch.no_grad():
                m.linear.weight.normal_(0, 0.02)
                lo.lora_B.normal_(0, 0.02)
                q = m.quantizers_weight[key]; q.start_calibration(); q(m.linear.weight.data); q.finish_calibration()
                for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                    qq.start_calibration(); qq(w.data); qq.finish_calibration()
            m.linear.weight.requires_grad_(False); m.linear.bias.requires_grad_(False)
            for tokens in args.sweep_tokens:
                x = torch.randn(tokens, K, device=dev)
                x[:, 7::97] *= 20.0
                gy = torch.randn(tokens, N, device=dev) * 1e-3
                with torch.no_grad():
                    m.calibration_mode = True
                    iq = m.quantizers_input[key]; iq.start_calibration(); m(x); iq.finish_calibration()
                    m.calibration_mode = False
                reps = max(3, min(20, int(2.0e11 / (tokens * N * K))))

                def fwd():
                    with torch.no_grad():
                        m(x)

                xg = x.clone().requires_grad_(True)

                def fwd_bwd():
                    xg.grad = None; lo.lora_A.grad = None; lo.lora_B.grad = None
                    m(xg).backward(gy)
                t_f, t_fb = timed(fwd, reps), timed(fwd_bwd, reps)
                f_fwd = 2.0 * tokens * N * K + 2.0 * tokens * R * (K + N)
                f_fb = 2 * 2.0 * tokens * N * K + 3 * 2.0 * tokens * R * (K + N)
                rows.append({"K": K, "N": N, "bits": bits, "quantizer": qt, "tokens": tokens, "fwd_ms": round(t_f, 4),
                             "fwd_bwd_ms": round(t_fb, 4), "fwd_Mtok_s": round(tokens / t_f / 1e3, 2),
                             "fwd_bwd_Mtok_s": round(tokens / t_fb / 1e3, 2),
                             "fwd_frac_of_peak": round(f_fwd / (t_f / 1e3) / 1e12 / peak, 3),
                             "fwd_bwd_frac_of_peak": round(f_fb / (t_fb / 1e3) / 1e12 / peak, 3)})
                del x, gy, xg
            del m
            torch.cuda.empty_cache()
    best = max(rows, key=lambda r: r["fwd_frac_of_peak"])
    out = {"metric": "SPLinearWithLoRA microbench, GPT-2 XL shapes (whole module call: quantise + LoRA + fused GEMM [+ STE backward])",
           "peak_tflops": peak, "lora_rank": R, "rows": rows, "best_fwd_frac_of_peak": best["fwd_frac_of_peak"]}
    out["fp8"] = fp8_section(args, dev, timed)
    return out


def fp8_section(args, dev, timed):
    """The e4m3 integer-code path (tcgen05.mma.kind::f8f6f4): 4-bit min-max with per-tensor scales, the reference's evaluation
    configuration (p1/deploy.py:210,238).  In-repo fp8 peak = torch._scaled_mm (cuBLASLt e4m3, 8192^3) on this GPU; GEMM-only
    time of spq_qgemm_f8 with the LoRA segment, and the whole SPLinearWithLoRA.forward."""
    import torch
    from llm_qat_on_gpt2_b200 import _lib
    from llm_qat_on_gpt2_b200.lora import SPLinearWithLoRA
    res = {"peak_fp8_tflops": None, "rows": []}
    try:
        n = 8192
        a = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn)
        b = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn).t()
        one = torch.ones((), device=dev)
        t = timed(lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16), 20)
        res["peak_fp8_tflops"] = 2.0 * n ** 3 / (t / 1e3) / 1e12
        res["peak_fp8_source"] = "torch._scaled_mm e4m3 x e4m3 -> bf16, 8192^3, CUDA events, this run"
        del a, b
    except Exception as e:                                  # library fp8 GEMM unavailable: report against the bf16 peak only
        res["peak_fp8_source"] = f"unavailable ({type(e).__name__})"
    R = 64
    for K, N in ((1600, 6400), (6400, 1600), (768, 2304)):
        torch.manual_seed(0)
        m = SPLinearWithLoRA(K, N, [4, 32], {4: R, 32: 0}, {4: R, 32: 0}, {4: "minmax", 32: None}, per_channel=False).to(dev)
        m.set_precision(4)
        lo = m.lora_adapters["4bit"]
        with torch.no_grad():
            m.linear.weight.normal_(0, 0.02); lo.lora_B.normal_(0, 0.02)
            q = m.quantizers_weight["4bit"]; q.start_calibration(); q(m.linear.weight.data); q.finish_calibration()
            for qq, w in ((lo.quantize_A, lo.lora_A), (lo.quantize_B, lo.lora_B)):
                qq.start_calibration(); qq(w.data); qq.finish_calibration()
        for tokens in (args.sweep_tokens[0], args.sweep_tokens[-1]):
            x = torch.randn(tokens, K, device=dev)
            with torch.no_grad():
                m.calibration_mode = True
                iq = m.quantizers_input["4bit"]; iq.start_calibration(); m(x); iq.finish_calibration()
                m.calibration_mode = False
                base, lora = m._operands_for(4, True)
                if base.get("f8") is None:
                    continue
                reps = max(3, min(20, int(2.0e11 / (tokens * N * K))))
                t_mod = timed(lambda: m(x), reps)
                a8 = torch.randint(-7, 8, (tokens, K), device=dev).float().to(torch.float8_e4m3fn).view(torch.uint8)
                t16 = torch.randn(tokens, R, device=dev).half()
                y = torch.empty(tokens, N, device=dev)
                t_g = timed(lambda: _lib.qgemm_f8(a8, base["f8"]["B8"], tokens, N, K, y, A2=t16, B2=lora["Bl_op8"], K2=R,
                                                  col_scale=base["f8"]["cs"], bias=m.linear.bias.detach()), reps)
            fl = 2.0 * tokens * N * (K + R)
            tf = fl / (t_g / 1e3) / 1e12
            res["rows"].append({"K": K, "N": N, "tokens": tokens, "module_fwd_ms": round(t_mod, 4), "gemm_ms": round(t_g, 4),
                                "gemm_tflops": round(tf, 1),
                                "gemm_frac_of_fp8_peak": round(tf / res["peak_fp8_tflops"], 3) if res["peak_fp8_tflops"] else None,
                                "gemm_frac_of_bf16_peak": round(tf / measured_peaks()["tflops_sustained"], 3)})
            del x
        del m
        torch.cuda.empty_cache()
    return res


def main():
    global BITS, METRIC
    args = parse_args()
    BITS = args.bits
    METRIC = f"GPT-2 SP tokens/s at {BITS}-bit (calibration pass + quantised forward)"
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
