"""spq_qgemm on the GEMM shapes of the headline step (GPT-2 small, 32 x 1024 tokens), each with the epilogue the model
uses, next to torch.matmul (cuBLAS fp16 -> fp16) on the same operands:

    python tools/gemm_shapes.py [reps] [only=<name>]

SPQ_GEMM_DEBUG (1 skip epilogue, 2 skip MMA, 8 skip the TMA store, 16 skip the staging writes) is read by the library
at its first launch: run once per setting.  Prints one line per shape (CSV) for profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("only=")]
M = int(os.environ.get("SPQ_SHAPES_M", 32768))
SHAPES = [
    # name, N, K, out dtype, residual, gelu, lse, K2 (LoRA segment)
    ("c_attn_f16", 2304, 768, "f16", False, False, False, 0),
    ("c_attn_f16_lora", 2304, 768, "f16", False, False, False, 64),
    ("attn_c_proj_res", 768, 768, "f32", True, False, False, 0),
    ("attn_c_proj_res_lora", 768, 768, "f32", True, False, False, 64),
    ("c_fc_gelu", 3072, 768, "f32", False, True, False, 0),
    ("c_fc_gelu_lora", 3072, 768, "f32", False, True, False, 64),
    ("c_fc_plain_f32", 3072, 768, "f32", False, False, False, 0),
    ("c_fc_plain_f16", 3072, 768, "f16", False, False, False, 0),
    ("mlp_c_proj_res", 768, 3072, "f32", True, False, False, 0),
    ("mlp_c_proj_res_lora", 768, 3072, "f32", True, False, False, 64),
    ("lora_down_768", 64, 768, "f16", False, False, False, 0),
    ("lora_down_3072", 64, 3072, "f16", False, False, False, 0),
    ("lm_head_lse", 50257, 768, "f32", False, False, True, 0),
]
torch.manual_seed(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                       # cold L2, as inside the step (working set >> L2)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


print("name,M,N,K,K2,out,us,TFLOP/s,cublas_us,cublas_TFLOP/s,GB_moved,GB/s")
for name, N, K, od, res, gelu, lse, K2 in SHAPES:
    if only and name not in only:
        continue
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") * 0.1).half()
    A2 = torch.randn(M, K2, device="cuda").half() if K2 else None
    B2 = (torch.randn(N, K2, device="cuda") * 0.1).half() if K2 else None
    ld = N if N % 4 == 0 else (N + 31) // 32 * 32
    buf = torch.empty(M, ld, device="cuda", dtype=torch.float16 if od == "f16" else torch.float32)
    out = buf[:, :N] if ld != N else buf
    bias = torch.randn(N, device="cuda"); cs = torch.rand(N, device="cuda") + 0.5
    rs = torch.rand(M, device="cuda") + 0.5
    C = torch.randn(M, N, device="cuda") if res else None
    if lse:
        fn = lambda: _lib.qgemm_lse(A, B, M, N, K, out, row_scale=rs, col_scale=cs)
    else:
        fn = lambda: _lib.qgemm(A, B, M, N, K, out, A2=A2, B2=B2, K2=K2, row_scale=rs, col_scale=cs, bias=bias, C=C,
                                activation=1 if gelu else 0)
    n = max(3, reps // 4) if lse else reps
    us = timed(fn, n) * 1e3
    Bt = B.t().contiguous()
    us_ref = timed(lambda: torch.matmul(A, Bt), n) * 1e3 if N * M * 2 < (8 << 30) else float("nan")
    fl = 2.0 * M * N * (K + K2)
    gb = (2.0 * (M + N) * (K + K2) + M * N * out.element_size() + (4.0 * M * N if res else 0.0)) / 1e9
    print(f"{name},{M},{N},{K},{K2},{od},{us:.1f},{fl / us / 1e6:.0f},{us_ref:.1f},{fl / us_ref / 1e6:.0f},{gb:.3f},{gb / us * 1e6:.0f}",
          flush=True)
    del A, B, buf, out, C, Bt
print("watchdog", _lib.debug_status())
