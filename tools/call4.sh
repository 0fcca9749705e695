#!/bin/bash
# TMA-store epilogue vs LSU-store epilogue (SPQ_GEMM_DEBUG=4 disables the TMA-store path) on the pair kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S="only=c_attn_f16 only=attn_c_proj_res only=c_fc_gelu only=c_fc_plain_f32 only=c_fc_plain_f16 only=mlp_c_proj_res"
SPQ_GEMM_CLUSTER4=0 timeout 300 python tools/gemm_shapes.py 20 $S > gpurun_out/c4_tma.csv 2>&1
SPQ_GEMM_CLUSTER4=0 SPQ_GEMM_DEBUG=4 timeout 300 python tools/gemm_shapes.py 20 $S > gpurun_out/c4_lsu.csv 2>&1
paste -d' ' <(cut -d, -f1,7,8 gpurun_out/c4_tma.csv) <(cut -d, -f7,8,9 gpurun_out/c4_lsu.csv)
python tools/write_bw.py > gpurun_out/c4_write_bw.log 2>&1; cat gpurun_out/c4_write_bw.log
