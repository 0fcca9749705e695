#!/bin/bash
# quad-cluster (B multicast) GEMM: bit-identity test, shapes A/B, full suite, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_cluster.py -x -q > gpurun_out/c3_cluster_test.log 2>&1; echo "rc=$?" >> gpurun_out/c3_cluster_test.log
tail -15 gpurun_out/c3_cluster_test.log
timeout 300 python tools/gemm_shapes.py 20 > gpurun_out/c3_shapes_quad.csv 2>&1
SPQ_GEMM_CLUSTER4=0 timeout 300 python tools/gemm_shapes.py 20 > gpurun_out/c3_shapes_pair.csv 2>&1
paste -d' ' <(cut -d, -f1,7,8 gpurun_out/c3_shapes_quad.csv) <(cut -d, -f7,8,9 gpurun_out/c3_shapes_pair.csv)
SPQ_GEMM_DEBUG=1 timeout 300 python tools/gemm_shapes.py 20 only=lm_head_lse only=c_fc_gelu only=c_fc_plain_f32 only=c_attn_f16 only=attn_c_proj_res only=mlp_c_proj_res > gpurun_out/c3_shapes_quad_noepi.csv 2>&1
cut -d, -f1,7,8 gpurun_out/c3_shapes_quad_noepi.csv
if grep -q "rc=0" gpurun_out/c3_cluster_test.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest.log
  tail -4 gpurun_out/c3_pytest.log
  B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
  python bench.py $B > gpurun_out/c3_bench_quad.json 2> gpurun_out/c3_bench_quad.err; echo "quad rc=$?"
  SPQ_GEMM_CLUSTER4=0 python bench.py $B > gpurun_out/c3_bench_pair.json 2> gpurun_out/c3_bench_pair.err; echo "pair rc=$?"
  python bench.py $B > gpurun_out/c3_bench_quadb.json 2>> gpurun_out/c3_bench_quad.err; echo "quadb rc=$?"
  for f in quad pair quadb; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"])
except Exception as e:
    print("$f failed", e)
PY
  done
fi
