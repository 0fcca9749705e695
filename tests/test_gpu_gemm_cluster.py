"""The quad-cluster GEMM (two CTA pairs per cluster, B tile fetched once per quad by TMA multicast) against the
pair-cluster GEMM of the same library: same tiles, same accumulation order, so every output must be BIT-identical.
The cluster mode is read from SPQ_GEMM_CLUSTER4 once per process, hence one subprocess per mode."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import hashlib, json, sys, torch
sys.path.insert(0, %r)
from llm_qat_on_gpt2_b200 import _lib as lib
res = {}
def digest(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
# (M, N, K, K2): ragged M (last quad half empty / one pair fully out of range), ragged N, LoRA segment
for (M, N, K, K2) in [(4096, 2304, 768, 0), (4100, 3072, 256, 64), (8192, 768, 3072, 64), (5000, 1000, 128, 0), (4609, 4800, 1600, 16)]:
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    A2 = torch.randn(M, max(K2, 8), device="cuda").half()[:, :K2] if K2 else None
    B2 = (torch.randn(N, max(K2, 8), device="cuda") * 0.1).half()[:, :K2] if K2 else None
    rs = torch.rand(M, device="cuda") + 0.5; cs = torch.rand(N, device="cuda") + 0.5; bias = torch.randn(N, device="cuda")
    C = torch.randn(M, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.qgemm(A, B, M, N, K, out, A2=A2, B2=B2, K2=K2, row_scale=rs, col_scale=cs, bias=bias)
    ref = (A.double() @ B.double().t() + (A2.double() @ B2.double().t() if K2 else 0)) * rs.double()[:, None] * cs.double() + bias.double()
    res[f"f32_{M}_{N}_{K}_{K2}"] = (digest(out), float((out.double() - ref).norm() / ref.norm()))
    lib.qgemm(A, B, M, N, K, out, A2=A2, B2=B2, K2=K2, row_scale=rs, col_scale=cs, bias=bias, C=C)
    res[f"res_{M}_{N}_{K}_{K2}"] = (digest(out), float((out.double() - ref - C.double()).norm() / ref.norm()))
    oh = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    lib.qgemm(A, B, M, N, K, oh, A2=A2, B2=B2, K2=K2, row_scale=rs, col_scale=cs, bias=bias, activation=1)
    res[f"f16gelu_{M}_{N}_{K}_{K2}"] = (digest(oh), 0.0)
# LM-head shape with the log-sum-exp epilogue (grouped tile order, odd vocabulary)
M, V, K = 2048, 50257, 128
torch.manual_seed(3)
A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(V, K, device="cuda") * 0.3).half()
ld = (V + 31) // 32 * 32
buf = torch.zeros((M, ld), device="cuda")
parts = lib.qgemm_lse(A, B, M, V, K, buf[:, :V])
lse = torch.logsumexp(buf[:, :V].double(), dim=1)
m = parts[..., 0]; s_ = parts[..., 1]; mm = m.max(dim=1).values
got = mm + torch.log((s_ * torch.exp(m - mm[:, None])).sum(dim=1))
res["lse"] = (digest(buf), float((got.double() - lse).abs().max()))
# e4m3 integer codes
M, N, K = 4096, 2304, 768
ca = torch.randint(-7, 8, (M, K), device="cuda"); cb = torch.randint(-7, 8, (N, K), device="cuda")
out = torch.empty(M, N, device="cuda")
lib.qgemm_f8(ca.float().to(torch.float8_e4m3fn).view(torch.uint8), cb.float().to(torch.float8_e4m3fn).view(torch.uint8), M, N, K, out)
res["f8"] = (digest(out), float((out.double() - ca.double() @ cb.double().t()).abs().max()))
res["watchdog"] = (str(lib.debug_status()), 0.0)
print("RESULT " + json.dumps(res))
'''


def _run(mode):
    env = dict(os.environ, SPQ_GEMM_CLUSTER4=mode)
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_quad_cluster_bit_identical_to_pair_cluster():
    pair, quad = _run("0"), _run("1")
    assert pair["watchdog"][0] == "0" and quad["watchdog"][0] == "0"
    for k in pair:
        assert quad[k][0] == pair[k][0], f"{k}: quad-cluster output differs from the pair-cluster output"
    for k, (_, err) in quad.items():
        if k.startswith(("f32_", "res_")):
            assert err <= 1e-5, (k, err)
    assert quad["lse"][1] <= 1e-4 and quad["f8"][1] == 0.0
