#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_cpt.py tests/test_gpu_upstream.py -x -q > gpurun_out/c18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c18_pytest.log; tail -5 gpurun_out/c18_pytest.log
for f in 1 0; do
SPQ_GRAD_SIDE=$f python bench.py --steps 3 --warmup 3 --train-steps 8 --no-cpu-baseline --sweep-tokens --train-strong 0 > gpurun_out/c18_bench_$f.json 2> gpurun_out/c18_bench_$f.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c18_bench_$f.json").read().strip().splitlines()[-1])
t=d["train"]; c=d["cpt_medium"]; print("side=$f train", round(t["value"]), round(t["ms_per_step"],2), t["loss"], "cpt", round(c["value"]), round(c["ms_per_step"],2), c["loss"])
PY
done
tail -3 gpurun_out/c18_bench_1.err
