#!/bin/bash
# last check of the round on the final commit: full GPU suite, smoke, refreshed 2-GPU line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/fc_pytest.log; tail -3 gpurun_out/fc_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fc_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/fc_smoke.log; tail -2 gpurun_out/fc_smoke.log
python bench.py > gpurun_out/fc_bench_1gpu.json 2> gpurun_out/fc_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/fc_bench_1gpu.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["step_parity"])
print("train", round(d["train"]["value"]), d["train"]["ms_per_step"], "cpt", round(d["cpt_medium"]["value"]), d["cpt_medium"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
PY
