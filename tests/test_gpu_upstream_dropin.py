"""Upstream's own callers running UNMODIFIED on top of the drop-in modules (VERDICT r01
"boundary" row): the part5_squad pytest suite (model / training-step / loss / distillation tests) and part1's
CalibrationManager + train_step.  Each runs in a fresh process (tests/upstream_dropin_runner.py) because the drop-in
is installed by aliasing module names.  pytest -m gpu."""
import json
import os
import subprocess
import sys

import pytest

from oracle import upstream as up

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not up.available(), reason="baseline/_ref not installed")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(what, timeout=1500):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "upstream_dropin_runner.py"), what], cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, (p.stdout[-1500:], p.stderr[-3000:])
    out = json.loads(lines[-1][7:])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "upstream_parity.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": f"dropin_{what}", **out}) + "\n")
    except OSError:
        pass
    return out


def test_part5_suite_runs_on_the_dropin():
    out = _run("part5_tests")
    assert out["linear_class_module"] == "llm_qat_on_gpt2_b200.lora"           # upstream's model was built on OUR linears
    assert "baseline/_ref" in out["models_squad_file"].replace(os.sep, "/")    # ... and is upstream's own file
    assert out["kernel_launches"] > 1000 and out["watchdog"] == 0
    failed = {k: v for k, v in out["results"].items() if v != "passed"}
    assert len(out["results"]) >= 17 and not failed, failed


def test_calibration_manager_and_train_step_on_the_dropin():
    out = _run("calibration")
    cmp = out["calibration_vs_upstream_cuda"]
    # weights: identical tensors on both sides.  min-max statistics / scales: <= 1 ulp (torch-CUDA divides by a scalar
    # through a reciprocal); log: <= 4 ulp on the range
    assert cmp["minmax_weight"]["max_ulp"] <= 1 and cmp["log_weight"]["max_ulp"] <= 4, cmp
    # inputs: statistics of activations that went through the fp16-operand GEMMs of the preceding layers
    assert "minmax_input" in cmp and "log_input" in cmp
    assert all("shape" not in v for v in cmp.values()), cmp
    assert out["params_moved"] == out["params_trainable"] > 0
    assert all(l == l and 0 < l < 100 for l in out["train_step_losses"]), out["train_step_losses"]
    assert out["watchdog"] == 0 and out["kernel_launches"] > 500
