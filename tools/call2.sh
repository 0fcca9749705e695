#!/bin/bash
# 16 epilogue warps + chunk-level LSE vs the 8-warp build: full GPU suite, GEMM shapes, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest.log
tail -5 gpurun_out/c2_pytest.log
python tools/gemm_shapes.py 20 > gpurun_out/c2_shapes_epi16.csv 2>&1
SPQ_LIB=$PWD/llm_qat_on_gpt2_b200/libspq_epi8.so python tools/gemm_shapes.py 20 > gpurun_out/c2_shapes_epi8.csv 2>&1
paste -d' ' <(cut -d, -f1,7,8 gpurun_out/c2_shapes_epi16.csv) <(cut -d, -f7,8,9 gpurun_out/c2_shapes_epi8.csv)
B="--steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
python bench.py $B > gpurun_out/c2_bench_epi16.json 2> gpurun_out/c2_bench_epi16.err; echo "epi16 rc=$?"
SPQ_LIB=$PWD/llm_qat_on_gpt2_b200/libspq_epi8.so python bench.py $B > gpurun_out/c2_bench_epi8.json 2> gpurun_out/c2_bench_epi8.err; echo "epi8 rc=$?"
python bench.py $B > gpurun_out/c2_bench_epi16b.json 2>> gpurun_out/c2_bench_epi16.err; echo "epi16b rc=$?"
for f in epi16 epi8 epi16b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],3), d["gpu_launches"], d["clocks"]["sm_mhz"], d["e2e"]["loss"])
except Exception as e:
    print("$f failed", e)
PY
done
tail -3 gpurun_out/c2_bench_epi16.err
