#!/bin/bash
# 8-GPU run of the bench as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/ev8_bench_8gpu.json 2> gpurun_out/ev8_bench_8gpu.err; echo "rc=$?"
tail -3 gpurun_out/ev8_bench_8gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ev8_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["roofline"]["frac"], d["step_parity"])
t=d["train"]; print("train", round(t["value"]), t["ms_per_step"], t.get("dp_parity",{}).get("ok"), t.get("phases_ms"))
c=d.get("cpt_medium"); print("cpt", c and round(c["value"]), c and c["ms_per_step"])
PY
