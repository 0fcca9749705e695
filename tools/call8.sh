#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cap in 0 4 6 10; do SPQ_LN_CTAS_PER_SM=$cap python tools/ln_fused_bench.py; done 2>&1 | tee gpurun_out/c8_ln.log
for l in rpi4 rpi1; do for cap in 0 4; do SPQ_LIB=$PWD/llm_qat_on_gpt2_b200/libspq_$l.so SPQ_LN_CTAS_PER_SM=$cap python tools/ln_fused_bench.py; done; done 2>&1 | tee -a gpurun_out/c8_ln.log
