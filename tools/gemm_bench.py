"""Time / profile one spq_qgemm shape:  python tools/gemm_bench.py M N K [reps] [half|f32] [resid|gelu]
(resid: the epilogue adds a float32 residual C, in place: D = C)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib
M, N, K = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
half = len(sys.argv) > 5 and sys.argv[5] == "half"
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
out = torch.empty(M, N, device="cuda", dtype=torch.float16 if half else torch.float32)
bias = torch.randn(N, device="cuda"); cs = torch.rand(N, device="cuda")
resid = len(sys.argv) > 6 and sys.argv[6] == "resid"
gelu = 1 if (len(sys.argv) > 6 and sys.argv[6] == "gelu") else 0
Cres = out if resid else None
for _ in range(3):
    _lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias, C=Cres, activation=gelu)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias, C=Cres, activation=gelu)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
Bt = B.t().contiguous()
for _ in range(3):
    ref = A @ Bt
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ref = A @ Bt
e1.record(); torch.cuda.synchronize()
ms_ref = e0.elapsed_time(e1) / reps
ref32 = (A[:2048].float() @ B.float().t()) * cs + bias
if resid:
    ref32 = None
if ref32 is not None and not gelu:
    out.zero_(); _lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias)
    err = ((out[:2048].float() - ref32).norm() / ref32.norm()).item()
    bad = (out.float() - ((A.float() @ B.float().t()) * cs + bias)).abs().max().item() if M * N <= 2 ** 28 else float('nan')
    print(f"rel err (first 2048 rows) {err:.2e}   max abs err (all rows) {bad:.3e}")
print(f"cuBLAS   {M}x{N}x{K} out=f16: {ms_ref*1e3:.1f} us  {2.0*M*N*K/ms_ref/1e9:.0f} TFLOP/s")
print(f"spq_qgemm {M}x{N}x{K} out={'f16' if half else 'f32'}{' +C' if resid else ''}{' +gelu' if gelu else ''}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s  watchdog {_lib.debug_status()}")
