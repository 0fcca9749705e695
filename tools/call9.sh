#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "ln_quantize_act_fused or ln_rowscale_stats_fused" > gpurun_out/c9_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/c9_kernels.log
tail -8 gpurun_out/c9_kernels.log
for cap in 0 1 2; do SPQ_LN_CTAS_PER_SM=$cap python tools/ln_fused_bench.py; done 2>&1 | tee gpurun_out/c9_ln.log
python tools/ln_fused_bench.py 32768 1024 2>&1 | tee -a gpurun_out/c9_ln.log
