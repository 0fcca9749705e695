#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cpt.py -x -q > gpurun_out/c17_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c17_pytest.log; tail -15 gpurun_out/c17_pytest.log
for f in 1 0; do
SPQ_CPT_FUSED_CE=$f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --train-steps 0 --sweep-tokens > gpurun_out/c17_bench_$f.json 2> gpurun_out/c17_bench_$f.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c17_bench_$f.json").read().strip().splitlines()[-1])
c=d["cpt_medium"]; print("fused=$f cpt", round(c["value"]), c["ms_per_step"], c["loss"])
PY
done
tail -3 gpurun_out/c17_bench_1.err
