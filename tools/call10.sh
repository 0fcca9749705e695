#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "act or quantize or fused" > gpurun_out/c10_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/c10_kernels.log
tail -5 gpurun_out/c10_kernels.log
for tab in 1 0; do for K in 768 3072; do SPQ_QACT_TAB=$tab python tools/ln_fused_bench.py 32768 $K 2>&1 | grep -E "M=|quantize_act " | tr '\n' ' '; echo " tab=$tab"; done; done | tee gpurun_out/c10_qact.log
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py $B > gpurun_out/c10_bench_$label.json 2> gpurun_out/c10_bench_$label.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c10_bench_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"])
except Exception as e:
    print("$label failed rc=$rc", e)
PY
}
run tab SPQ_X=1
run reg SPQ_QACT_TAB=0
run tabb SPQ_X=1
