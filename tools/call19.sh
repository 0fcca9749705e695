#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cpt.py tests/test_gpu_training.py -x -q > gpurun_out/c19_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest.log; tail -3 gpurun_out/c19_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
for p in 1 0 1; do
SPQ_E2E_PIPELINE=$p python bench.py $B > gpurun_out/c19_bench_$p.json 2> gpurun_out/c19_bench_$p.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c19_bench_$p.json").read().strip().splitlines()[-1])
print("pipeline=$p", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["value"]), d["e2e"]["loss"], d["e2e"]["step_wall_ms"][:4])
PY
done
