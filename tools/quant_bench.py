"""Time spq_quantize_act / spq_minmax_stats / layernorm in isolation against the HBM roofline.
    python tools/quant_bench.py M K [log|minmax|raw] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib
M, K = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "log"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda")
x[:, 5] *= 20
a_q = torch.empty(M, K, device="cuda", dtype=torch.float16)
a_raw = torch.empty(M, K, device="cuda", dtype=torch.float16)
rs = torch.empty(M, device="cuda")
rawm = torch.full((K,), 0.5, device="cuda")
if mode == "log":
    lmin = torch.full((K,), -16.6, device="cuda"); lrng = torch.full((K,), 19.0, device="cuda")
    cm = torch.full((K,), 2.0 ** 5, device="cuda")
    call = lambda: _lib.quantize_act(x, lrng, lmin, _lib.PER_COL, _lib.LOG, 8, True, _lib.OPERAND_DEQUANT, cm, 1.0, a_q, a_raw, rawm)
    nbytes = M * K * 8
elif mode == "minmax":
    sc = torch.full((K,), 4.0 / 7, device="cuda"); zp = torch.zeros(K, device="cuda")
    call = lambda: _lib.quantize_act(x, sc, zp, _lib.PER_COL, _lib.MINMAX, 4, True, _lib.OPERAND_CODE, None, 1.0, a_q, a_raw, rawm)
    nbytes = M * K * 8
elif mode == "stats":
    smin = torch.empty(K, device="cuda"); smax = torch.empty(K, device="cuda"); stt = torch.zeros(1, dtype=torch.int32, device="cuda")
    call = lambda: _lib.minmax_stats(x, _lib.PER_COL, True, 1e-5, smin, smax, accumulate=False, state=stt)
    nbytes = M * K * 4
else:
    call = lambda: _lib.rowscale_f16(x, a_raw, rs)
    nbytes = M * K * 6
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"quantize_act[{mode}] {M}x{K}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s (algorithmic {nbytes/1e6:.0f} MB)  SPQ_QUANT_DEBUG={os.environ.get('SPQ_QUANT_DEBUG','0')}")
