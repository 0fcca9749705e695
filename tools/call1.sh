#!/bin/bash
# diagnostics call: graphed step test + A/B, GEMM shapes with epilogue switches, fresh launch list, ncu of the LM head
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_modules.py -x -q -k "graphed" > gpurun_out/c1_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
python tools/gemm_shapes.py 20 > gpurun_out/c1_shapes_default.csv 2>&1
SPQ_GEMM_STORE_HINT=0 python tools/gemm_shapes.py 20 only=lm_head_lse only=c_fc_gelu > gpurun_out/c1_shapes_hint0.csv 2>&1
SPQ_GEMM_STORE_HINT=1 python tools/gemm_shapes.py 20 only=lm_head_lse only=c_fc_gelu only=c_fc_plain_f32 only=c_attn_f16 only=attn_c_proj_res only=mlp_c_proj_res > gpurun_out/c1_shapes_hint1.csv 2>&1
SPQ_GEMM_DEBUG=8 python tools/gemm_shapes.py 20 only=lm_head_lse only=c_fc_gelu only=c_fc_plain_f32 only=c_attn_f16 only=attn_c_proj_res only=mlp_c_proj_res > gpurun_out/c1_shapes_nostore.csv 2>&1
SPQ_GEMM_DEBUG=1 python tools/gemm_shapes.py 20 only=lm_head_lse only=c_fc_gelu only=c_fc_plain_f32 only=c_attn_f16 only=attn_c_proj_res only=mlp_c_proj_res > gpurun_out/c1_shapes_noepi.csv 2>&1
cat gpurun_out/c1_shapes_default.csv
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
python bench.py $B > gpurun_out/c1_bench_graphed.json 2> gpurun_out/c1_bench_graphed.err; echo "graphed rc=$?"
python bench.py $B --eager-step > gpurun_out/c1_bench_eager.json 2> gpurun_out/c1_bench_eager.err; echo "eager rc=$?"
python bench.py $B > gpurun_out/c1_bench_graphed2.json 2>> gpurun_out/c1_bench_graphed.err; echo "graphed2 rc=$?"
for f in graphed eager graphed2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c1_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f failed", e)
PY
done
P="--profile-one-step --steps 1 --warmup 2 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
python bench.py $P > gpurun_out/c1_p1_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "spq_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv python bench.py $P > gpurun_out/c1_p1_ncu.log 2>&1
echo "launch list rc=$?"
python tools/gemm_shapes.py 4 only=lm_head_lse > gpurun_out/c1_lm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qgemm_nt -s 2 -c 1 -o gpurun_out/c1_lmhead python tools/gemm_shapes.py 4 only=lm_head_lse > gpurun_out/c1_lm_ncu.log 2>&1
echo "lm ncu rc=$?"
python tools/gemm_shapes.py 4 only=c_fc_gelu > gpurun_out/c1_fc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qgemm_nt -s 2 -c 1 -o gpurun_out/c1_cfc python tools/gemm_shapes.py 4 only=c_fc_gelu > gpurun_out/c1_fc_ncu.log 2>&1
echo "cfc ncu rc=$?"
ls -la gpurun_out | tail -20
