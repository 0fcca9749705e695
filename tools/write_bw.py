"""Pure HBM write bandwidth and cuBLAS fp16-in / fp32-out, as yardsticks for the fp32-output GEMM epilogue."""
import torch
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
M, N, K = 32768, 2304, 768
out = torch.empty(M, N, device="cuda")
big = torch.empty(1 << 28, device="cuda")            # 1 GiB
ms = t(lambda: out.fill_(1.0)); print(f"fill 302 MB: {ms*1e3:.1f} us  {out.numel()*4/ms/1e6:.0f} GB/s")
ms = t(lambda: big.fill_(1.0)); print(f"fill 1 GiB: {ms*1e3:.1f} us  {big.numel()*4/ms/1e6:.0f} GB/s")
src = torch.empty(M, N, device="cuda")
ms = t(lambda: out.copy_(src)); print(f"copy 302 MB: {ms*1e3:.1f} us  {2*out.numel()*4/ms/1e6:.0f} GB/s (r+w)")
A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(K, N, device="cuda") * 0.1).half()
try:
    ms = t(lambda: torch.mm(A, B, out_dtype=torch.float32)); print(f"cuBLAS f16 in / f32 out: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
except Exception as e:
    print("mm out_dtype unsupported:", e)
ms = t(lambda: torch.mm(A, B)); print(f"cuBLAS f16 in / f16 out: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
A32 = A.float(); B32 = B.float()
torch.backends.cuda.matmul.allow_tf32 = True
ms = t(lambda: torch.mm(A32, B32)); print(f"cuBLAS tf32 (f32 in/out): {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
