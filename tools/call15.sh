#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py tests/test_gpu_upstream.py -x -q -k "mse_select or train or Train or step" > gpurun_out/c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest.log; tail -4 gpurun_out/c15_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --cpt-steps 0 --sweep-tokens --train-strong 0 > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c15_bench.json").read().strip().splitlines()[-1])
t=d["train"]; print("train", round(t["value"]), t["ms_per_step"], t["loss"], t["phases_ms"])
PY
tail -3 gpurun_out/c15_bench.err
