#!/bin/bash
# 2-GPU run of the bench as the driver launches it (graphed step with the statistics exchange between the two graphs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/ev2_bench_2gpu.json 2> gpurun_out/ev2_bench_2gpu.err; echo "rc=$?"
tail -5 gpurun_out/ev2_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ev2_bench_2gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["roofline"]["frac"], d["config"]["launch_mode"])
t=d["train"]; print("train", round(t["value"]), t["ms_per_step"], t.get("dp_parity"))
c=d.get("cpt_medium"); print("cpt", c and round(c["value"]), c and c["ms_per_step"])
PY
