"""First-contact probe for the GPU box: verbose GEMM correctness + timing (not a pytest file)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from llm_qat_on_gpt2_b200 import _lib

def main():
    _lib.load_library()
    print("device", torch.cuda.get_device_name(0), _lib.device_info(), flush=True)
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 64, 64), (128, 128, 128), (256, 512, 192), (300, 200, 72), (2048, 2304, 768)]:
        try:
            A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
            out = torch.full((M, N), float("nan"), device="cuda")
            _lib.qgemm(A, B, M, N, K, out)
            torch.cuda.synchronize()
            st = _lib.debug_status()
            ref = A.double() @ B.double().t()
            err = ((out.double() - ref).norm() / ref.norm()).item()
            nan = torch.isnan(out).sum().item()
            print(f"NT {M}x{N}x{K}: rel_err {err:.3e} nan {nan} watchdog {st}", flush=True)
            if err > 1e-4 and M <= 300:
                d = (out.double() - ref).abs()
                print("   worst rows", d.max(dim=1).values.topk(4).indices.tolist(), "worst cols", d.max(dim=0).values.topk(4).indices.tolist())
                print("   out[0,:8]", out[0, :8].tolist()); print("   ref[0,:8]", ref[0, :8].tolist())
        except Exception:
            traceback.print_exc()
    for (Mred, I, J) in [(64, 128, 64), (256, 128, 64), (1000, 768, 64), (333, 200, 136), (2048, 768, 768)]:
        try:
            P = torch.randn(Mred, I, device="cuda").half(); Q = (torch.randn(Mred, J, device="cuda") * 0.1).half()
            out = torch.empty(I, J, device="cuda")
            _lib.gemm_tn(P, Q, out)
            torch.cuda.synchronize()
            st = _lib.debug_status()
            ref = P.double().t() @ Q.double()
            err = ((out.double() - ref).norm() / ref.norm()).item()
            print(f"TN {Mred}x{I}x{J}: rel_err {err:.3e} watchdog {st}", flush=True)
            if err > 1e-4 and I <= 300:
                print("   out[0,:8]", out[0, :8].tolist()); print("   ref[0,:8]", ref[0, :8].tolist())
        except Exception:
            traceback.print_exc()
    # timing vs cuBLAS fp16
    for (M, N, K) in [(32768, 2304, 768), (32768, 768, 768), (32768, 3072, 768), (32768, 768, 3072), (8192, 2304, 768),
                      (16384, 4800, 1600), (16384, 6400, 1600), (32768, 50257, 768)]:
        try:
            A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
            out = torch.empty(M, N, device="cuda")
            bias = torch.randn(N, device="cuda"); cs = torch.rand(N, device="cuda")
            for _ in range(3):
                _lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                _lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * M * N * K / ms / 1e9
            Bt = B.t().contiguous()
            for _ in range(3):
                ref = A @ Bt
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ref = A @ Bt
            e1.record(); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / reps
            print(f"time {M}x{N}x{K}: spq_qgemm(fp32 out) {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS fp16 out {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.0f} TFLOP/s  watchdog {_lib.debug_status()}", flush=True)
            del A, B, out, ref, Bt
        except Exception:
            traceback.print_exc()

if __name__ == "__main__":
    main()
