"""LayerNorm-fused activation-side kernels against the kernels they replace, at the headline shape (32768 x 768):
    python tools/ln_fused_bench.py [M] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib as lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 768
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda") * 3
w = torch.rand(K, device="cuda") + 0.5; b = torch.randn(K, device="cuda") * 0.1
y = torch.empty_like(x); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
lib.layernorm_fwd(x, w, b, 1e-5, y, mean, rstd)
smin = torch.empty(K, device="cuda"); smax = torch.empty(K, device="cuda"); state = torch.zeros(1, dtype=torch.int32, device="cuda")
lib.minmax_stats(y, lib.PER_COL, True, 1e-5, smin, smax, accumulate=False, state=state)
sc = torch.empty(K, device="cuda"); zp = torch.empty(K, device="cuda")
lib.finish_calibration(smin, smax, lib.LOG, True, 8, 1e-5, sc, zp)
cm = torch.ones(K, device="cuda"); rm = torch.ones(K, device="cuda")
a_q = torch.empty(M, K, dtype=torch.float16, device="cuda"); a_raw = torch.empty_like(a_q); rs = torch.empty(M, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
t_ln = timed(lambda: lib.layernorm_fwd(x, w, b, 1e-5, y, mean, rstd))
t_q = timed(lambda: lib.quantize_act(y, sc, zp, lib.PER_COL, lib.LOG, 8, True, lib.OPERAND_DEQUANT, cm, 1.0, a_q, a_raw, rm))
t_f = timed(lambda: lib.ln_quantize_act(x, w, b, 1e-5, sc, zp, lib.PER_COL, lib.LOG, 8, True, lib.OPERAND_DEQUANT, cm, 1.0, a_q, a_raw, rm))
t_rs = timed(lambda: lib.rowscale_f16(y, a_raw, rs))
t_st = timed(lambda: lib.minmax_stats(y, lib.PER_COL, True, 1e-5, smin, smax, accumulate=False, state=state))
t_fr = timed(lambda: lib.ln_rowscale_stats(x, w, b, 1e-5, a_raw, rs, stats_mode=2, stat_eps=1e-5, stat_min=smin, stat_max=smax, accumulate=False, state=state))
t_fr0 = timed(lambda: lib.ln_rowscale_stats(x, w, b, 1e-5, a_raw, rs))
gb = M * K / 1e9
print(f"M={M} K={K} CTAs-per-SM cap: {os.environ.get('SPQ_LN_CTAS_PER_SM', '-')}  lib: {os.environ.get('SPQ_LIB', 'default')}")
print(f"layernorm_fwd {t_ln:.1f} us ({8 * gb / t_ln * 1e3:.2f} TB/s)   quantize_act {t_q:.1f} us ({8 * gb / t_q * 1e3:.2f} TB/s)   sum {t_ln + t_q:.1f}")
print(f"ln_quantize_act {t_f:.1f} us ({8 * gb / t_f * 1e3:.2f} TB/s)")
print(f"rowscale {t_rs:.1f} us   colstats {t_st:.1f} us   LN + rowscale + stats {t_ln + t_rs + t_st:.1f}")
print(f"ln_rowscale_stats {t_fr:.1f} us ({6 * gb / t_fr * 1e3:.2f} TB/s)   without stats {t_fr0:.1f} us")
