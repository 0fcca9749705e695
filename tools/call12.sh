#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S="only=attn_c_proj_res only=attn_c_proj_res_lora only=mlp_c_proj_res only=mlp_c_proj_res_lora"
python tools/gemm_shapes.py 20 $S > gpurun_out/c12_prefetch.csv 2>&1
SPQ_GEMM_DEBUG=32 python tools/gemm_shapes.py 20 $S > gpurun_out/c12_noprefetch.csv 2>&1
paste -d' ' <(cut -d, -f1,7,8 gpurun_out/c12_prefetch.csv) <(cut -d, -f7,8,9 gpurun_out/c12_noprefetch.csv)
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_cluster.py -x -q -k "gemm" > gpurun_out/c12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c12_pytest.log; tail -3 gpurun_out/c12_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
for label in pf nopf pfb; do
  if [ $label = nopf ]; then export SPQ_GEMM_DEBUG=32; else unset SPQ_GEMM_DEBUG; fi
  python bench.py $B > gpurun_out/c12_bench_$label.json 2> gpurun_out/c12_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c12_bench_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"], d["step_parity"])
except Exception as e:
    print("$label failed", e)
PY
done
tail -3 gpurun_out/c12_bench_pf.err
