"""bench.py contract on CPU: the reference arm (the unmodified upstream model from baseline/_ref on torch-CPU; the
numpy port only when that copy is absent) prints ONE JSON line with the keys
the driver reads, on the GPU arm's metric / unit / config; ranks other than 0 stay silent; the GPU arm fails
loudly without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seq", "32", "--cpu-sample-batch", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("GPT-2 SP tokens/s at 8-bit") and d["vs_baseline"] is None and d["data"] == "synthetic"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    from oracle import upstream
    assert cb["kind"] == ("reference" if upstream.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["warmup"] == 0 and d["steps"] == 1                      # the arm runs exactly the K / W it was given
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_port_fallback():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seq", "32", "--cpu-sample-batch", "1"],
             env={"SPQ_CPU_PORT": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_are_silent():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--train-steps", "0"], timeout=300)
    assert p.returncode != 0
    assert p.stdout.strip() == ""                     # no JSON line from a CPU fallback
