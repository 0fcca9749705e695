#!/bin/bash
# ncu launch list of ONE optimizer step of the training bench (config 3) + the new fused-LN test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "layernorm_fused" > gpurun_out/pt_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pt_pytest.log; tail -3 gpurun_out/pt_pytest.log
P="--profile-train-step --steps 1 --warmup 1 --no-cpu-baseline --cpt-steps 0 --sweep-tokens"
python bench.py $P > gpurun_out/pt_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pt_launches.csv python bench.py $P > gpurun_out/pt_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/pt_ncu.log
