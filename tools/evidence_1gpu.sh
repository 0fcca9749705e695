#!/bin/bash
# end-of-round evidence, 1 GPU: full suite, smoke, default bench, reference arm, launch list, DRAM per GEMM launch, ncu --set full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ev1_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ev1_pytest.log; tail -3 gpurun_out/ev1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev1_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ev1_smoke.log; tail -4 gpurun_out/ev1_smoke.log
python bench.py > gpurun_out/ev1_bench_1gpu.json 2> gpurun_out/ev1_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/ev1_bench_ref.json 2> gpurun_out/ev1_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/ev1_bench_1gpu.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["roofline"]["frac"], d["clocks"], d["step_parity"])
print("train", round(d["train"]["value"]), d["train"]["ms_per_step"], "cpt", round(d["cpt_medium"]["value"]), d["cpt_medium"]["ms_per_step"], "sweep best", d["qlinear_sweep"]["best_fwd_frac_of_peak"])
r=json.loads(open("gpurun_out/ev1_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
P="--profile-one-step --steps 1 --warmup 2 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens"
python bench.py $P > gpurun_out/ev1_p1_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "spq_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev1_launches.csv python bench.py $P > gpurun_out/ev1_p1_ncu.log 2>&1
echo "launch list rc=$?"
python bench.py $P > gpurun_out/ev1_p2_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "spq_step/" -k regex:qgemm_nt --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev1_qgemm_dram.csv python bench.py $P > gpurun_out/ev1_p2_ncu.log 2>&1
echo "dram rc=$?"
python tools/gemm_shapes.py 4 only=lm_head_lse > gpurun_out/ev1_lm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qgemm_nt -s 2 -c 1 -o gpurun_out/ev1_lmhead python tools/gemm_shapes.py 4 only=lm_head_lse > gpurun_out/ev1_lm_ncu.log 2>&1
echo "lm ncu rc=$?"
python tools/gemm_shapes.py 4 only=c_attn_f16_lora > gpurun_out/ev1_ca_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qgemm_nt -s 2 -c 1 -o gpurun_out/ev1_cattn python tools/gemm_shapes.py 4 only=c_attn_f16_lora > gpurun_out/ev1_ca_ncu.log 2>&1
echo "c_attn ncu rc=$?"
python tools/ln_fused_bench.py > gpurun_out/ev1_ln_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ln_quantize_act -s 3 -c 1 -o gpurun_out/ev1_lnq python tools/ln_fused_bench.py > gpurun_out/ev1_lnq_ncu.log 2>&1
echo "lnq ncu rc=$?"
python tools/gemm_shapes.py 20 > gpurun_out/ev1_shapes.csv 2>&1; cat gpurun_out/ev1_shapes.csv | cut -d, -f1,7,8,9
