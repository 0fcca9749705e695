#!/bin/bash
# LayerNorm fused into the activation-side kernels: kernel tests, module tests, full suite, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "ln_quantize_act_fused or ln_rowscale_stats_fused" > gpurun_out/c7_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/c7_kernels.log
tail -25 gpurun_out/c7_kernels.log
timeout 900 python -m pytest tests/test_gpu_modules.py -x -q -k "layernorm_fused or graphed or refresher" > gpurun_out/c7_modules.log 2>&1; echo "rc=$?" >> gpurun_out/c7_modules.log
tail -25 gpurun_out/c7_modules.log
if grep -q "rc=0" gpurun_out/c7_kernels.log && grep -q "rc=0" gpurun_out/c7_modules.log; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest.log
  tail -6 gpurun_out/c7_pytest.log
fi
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py $B > gpurun_out/c7_bench_$label.json 2> gpurun_out/c7_bench_$label.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c7_bench_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"])
except Exception as e:
    print("$label failed rc=$rc", e)
PY
}
run fused SPQ_X=1
run unfused SPQ_FUSE_LN=0
run fusedb SPQ_X=1
tail -3 gpurun_out/c7_bench_fused.err
