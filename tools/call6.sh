#!/bin/bash
# fp16 MLP activation option: test + bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "graphed or mlp_activation" > gpurun_out/c6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest.log
tail -12 gpurun_out/c6_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py $B > gpurun_out/c6_bench_$label.json 2> gpurun_out/c6_bench_$label.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c6_bench_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"])
except Exception as e:
    print("$label failed rc=$rc", e)
PY
}
run f16 SPQ_X=1
run f32 SPQ_MLP_ACT=fp32
run f16b SPQ_X=1
tail -3 gpurun_out/c6_bench_f16.err
