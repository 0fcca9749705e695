#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -x -q -k "softmax_loss or lm_head_loss or train" > gpurun_out/c16_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c16_pytest.log; tail -4 gpurun_out/c16_pytest.log
python tools/loss_bench.py > gpurun_out/c16_loss_bench.log 2>&1; cat gpurun_out/c16_loss_bench.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --cpt-steps 0 --sweep-tokens --train-strong 0 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
t=d["train"]; print("train", round(t["value"]), t["ms_per_step"], t["loss"], t["phases_ms"])
PY
