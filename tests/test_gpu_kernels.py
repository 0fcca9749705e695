"""GPU parity tests of the CUDA kernels, called through the C ABI (ctypes), against the numpy
oracle and the reference-generated golden fixtures.  Run on the B200 box: pytest -m gpu.

Bars (BASELINE.json): min/max statistics, scale / zero-point and integer codes bit-exact;
dequantised values, GEMM outputs and gradients within rel 1e-3 of float32 (tolerances are
written next to each assertion and are usually much tighter).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import (QuantizerState, collect_statistics, finish_calibration, log_quantize, minmax_quantize,
                    switchable_layernorm_backward, switchable_layernorm_forward)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    from llm_qat_on_gpt2_b200 import _lib
    _lib.load_library()
    return _lib


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def rel_fro(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def heavy_tailed(shape, seed, zeros=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape) * np.exp(1.5 * rng.standard_normal(shape))
    if shape[-1] >= 8:
        x[..., 3] *= 20.0
        x[..., -2] *= 0.01
    if zeros:
        flat = x.reshape(-1)
        idx = rng.permutation(flat.size)[: max(2, flat.size // 50)]
        flat[idx[: len(idx) // 2]] = 0.0
        flat[idx[len(idx) // 2:]] = 3e-6
    return x.astype(np.float32)


def test_device_is_blackwell(lib):
    sms, major, minor = lib.device_info()
    assert major == 10, f"expected an sm_100 device, got sm_{major}{minor}"
    assert sms >= 100


# --------------------------------------------------------------------------- statistics
@pytest.mark.parametrize("shape,channel_dim,per_channel", [
    ((4, 96, 768), -1, True), ((2304, 768), 0, True), ((768, 64), 1, True), ((64, 2304), 1, True),
    ((3, 50, 40), -1, False), ((37, 13), 0, True), ((5, 7, 33), -1, True), ((1, 1, 8), -1, True),
])
@pytest.mark.parametrize("qtype", ["minmax", "log"])
def test_calibration_statistics_bit_exact(lib, shape, channel_dim, per_channel, qtype):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize
    q = LearnableFakeQuantize(8, channel_dim=channel_dim, quantizer_type=qtype, per_channel=per_channel).cuda()
    o = QuantizerState(8, channel_dim=channel_dim, quantizer_type=qtype, per_channel=per_channel)
    q.start_calibration(); o.start_calibration()
    for b in range(3):
        x = heavy_tailed(shape, 10 + b)
        if b == 1 and qtype == "log":
            x[:] = 1e-7                      # a batch with nothing above eps leaves the statistics alone
        xd = dev(x)
        assert q(xd) is xd                   # collecting mode passes x through
        collect_statistics(o, x)
    q.finish_calibration(); finish_calibration(o)
    assert q.calibrated and q.num_batches_collected == 3
    for name in ("running_min", "running_max", "scale", "zero_point"):
        got = getattr(q, name).cpu().numpy()
        want = getattr(o, name)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert np.array_equal(got, want), f"{name} not bit-exact (max |d| {np.abs(got - want).max()})"


def test_statistics_nan_and_allzero(lib):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize
    # NaN propagates per channel, like torch.min/max
    x = heavy_tailed((64, 40), 3, zeros=False)
    x[5, 7] = np.nan
    q = LearnableFakeQuantize(8, channel_dim=-1, quantizer_type="minmax").cuda()
    q.start_calibration(); q(dev(x)); q.finish_calibration()
    rmin = q.running_min.cpu().numpy().reshape(-1)
    assert np.isnan(rmin[7]) and np.isfinite(np.delete(rmin, 7)).all()
    # all-zero tensor through a log quantiser: the reference's default-shape quirk
    g = np.load(os.path.join(GOLDEN, "quant_log_allzero.npz"))
    q = LearnableFakeQuantize(8, channel_dim=1, quantizer_type="log").cuda()
    q.start_calibration(); q(dev(g["x"])); q.finish_calibration()
    for name in ("running_min", "running_max", "scale", "zero_point"):
        got = getattr(q, name).cpu().numpy()
        assert got.shape == g[name].shape and np.array_equal(got, g[name]), name
    assert np.array_equal(q(dev(g["x"])).cpu().numpy(), g["out"])
    with pytest.raises(RuntimeError):
        LearnableFakeQuantize(8).cuda()(dev(x))          # uncalibrated
    with pytest.raises(RuntimeError):
        LearnableFakeQuantize(8)(torch.zeros(4, 4))      # CPU tensor: no fallback


# --------------------------------------------------------------------------- quantise
QUANT_CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "quant_*.npz"))
                     if "allzero" not in p)


@pytest.mark.parametrize("name", QUANT_CASES)
def test_quantize_against_golden(lib, name):
    """Reference-generated vectors: calibrate on the GPU, then quantise with the reference's own
    calibrated parameters; codes must match the reference bit for bit."""
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, quantize_codes
    g = np.load(os.path.join(GOLDEN, f"quant_{name}.npz"))
    bits, cd, sym, pc, is_in, qt = [int(v) for v in g["meta"]]
    qtype = "minmax" if qt == 0 else "log"
    q = LearnableFakeQuantize(bits, channel_dim=None if cd == -99 else cd, quantizer_type=qtype,
                              symmetric=bool(sym), per_channel=bool(pc), is_input=bool(is_in)).cuda()
    q.start_calibration()
    for xb in g["x_calib"]:
        q(dev(xb))
    q.finish_calibration()
    for key in ("running_min", "running_max", "scale", "zero_point"):
        got = getattr(q, key).cpu().numpy()
        assert got.shape == g[key].shape, key
        if qtype == "minmax":
            assert np.array_equal(got, g[key]), key
        else:   # torch-CPU's SLEEF log2 is 1 ulp off the correctly rounded value on 0.013 % of inputs
            d = np.abs(got.view(np.int32).astype(np.int64) - g[key].view(np.int32).astype(np.int64))
            assert d.max() <= (4 if key == "scale" else 1), (key, d.max())
    out, codes, sign = quantize_codes(dev(g["x_test"]), dev(g["scale"]), dev(g["zero_point"]), bits, bool(sym), qtype)
    assert np.array_equal(codes.cpu().numpy(), g["codes"]), f"{(codes.cpu().numpy() != g['codes']).sum()} code mismatches"
    out = out.cpu().numpy()
    if qtype == "minmax":
        assert np.array_equal(out, g["out"])
    else:
        assert np.array_equal(out == 0, g["out"] == 0)
        nz = g["out"] != 0
        assert np.max(np.abs(out[nz] - g["out"][nz]) / np.abs(g["out"][nz])) <= 2e-5     # value path uses FMA + ex2.approx; only the level index is bit-exact
        assert np.array_equal(sign.cpu().numpy(), np.sign(g["out"]).astype(np.int8))


@pytest.mark.parametrize("qtype,bits,symmetric", [("minmax", 4, True), ("minmax", 8, True), ("minmax", 8, False),
                                                  ("minmax", 3, True), ("log", 8, True), ("log", 4, True),
                                                  ("log", 8, False), ("log", 11, True)])
@pytest.mark.parametrize("layout", ["per_col", "per_row", "per_tensor"])
def test_quantize_codes_bit_exact_large(lib, qtype, bits, symmetric, layout):
    """4 Mi elements per case against the oracle: every code identical."""
    from llm_qat_on_gpt2_b200 import quantize_codes
    rows, cols = 4096, 1024
    x = heavy_tailed((rows, cols), 100 + bits)
    cd, pc = {"per_col": (-1, True), "per_row": (0, True), "per_tensor": (0, False)}[layout]
    o = QuantizerState(bits, channel_dim=cd, quantizer_type=qtype, symmetric=symmetric, per_channel=pc)
    o.start_calibration(); collect_statistics(o, x[: rows // 2]); finish_calibration(o)   # half: the rest clips
    if layout == "per_row":
        o.scale = np.resize(o.scale, (rows, 1)).astype(np.float32); o.zero_point = np.resize(o.zero_point, (rows, 1)).astype(np.float32)
    out, codes, sign = quantize_codes(dev(x), dev(o.scale), dev(o.zero_point), bits, symmetric, qtype)
    if qtype == "minmax":
        ref_out, ref_codes = minmax_quantize(x, o.scale, o.zero_point, bits, symmetric)
        assert np.array_equal(codes.cpu().numpy(), ref_codes)
        assert np.array_equal(out.cpu().numpy(), ref_out)
    else:
        ref_out, ref_level, ref_sign, ref_zero = log_quantize(x, o.zero_point, o.scale, bits, symmetric)
        got = codes.cpu().numpy()
        assert np.array_equal(got, ref_level), f"{(got != ref_level).sum()} level mismatches of {got.size}"
        assert np.array_equal(sign.cpu().numpy() == 0, ref_zero | (x == 0))
        o_ = out.cpu().numpy(); nz = ref_out != 0
        assert np.max(np.abs(o_[nz] - ref_out[nz]) / np.abs(ref_out[nz])) <= 2e-5     # value path uses FMA + ex2.approx; only the level index is bit-exact


def test_quantize_act_operands(lib):
    """Fused activation kernel: code operand exact, raw operand = x * 2^e[k] up to fp16 rounding;
    row-scaled variant (uncalibrated paths): x = x16 * row_scale with power-of-two row scales."""
    M, K = 777, 768
    x = heavy_tailed((M, K), 5)
    raw_mul = (2.0 ** np.random.default_rng(0).integers(-3, 4, K)).astype(np.float32)
    for qtype, bits in (("minmax", 4), ("minmax", 8), ("log", 8)):
        o = QuantizerState(bits, channel_dim=-1, quantizer_type=qtype, is_input=True)
        o.start_calibration(); collect_statistics(o, x); finish_calibration(o)
        a_q = torch.empty((M, K), dtype=torch.float16, device="cuda")
        a_raw = torch.empty((M, K), dtype=torch.float16, device="cuda")
        lib.quantize_act(dev(x), dev(o.scale.reshape(-1)), dev(o.zero_point.reshape(-1)), lib.PER_COL,
                         lib.QTYPE[qtype], bits, True, lib.OPERAND_CODE, None, 1.0, a_q, a_raw, dev(raw_mul))
        if qtype == "minmax":
            _, codes = minmax_quantize(x, o.scale.reshape(1, -1), o.zero_point.reshape(1, -1), bits, True)
        else:
            _, codes, _, _ = log_quantize(x, o.zero_point.reshape(1, -1), o.scale.reshape(1, -1), bits, True)
        assert np.array_equal(a_q.float().cpu().numpy(), codes.astype(np.float32))
        want = np.clip(x * raw_mul[None, :], -65504, 65504)
        got = a_raw.float().cpu().numpy()
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 2.0 ** -10
    x16 = torch.empty((M, K), dtype=torch.float16, device="cuda")
    rs = torch.empty(M, dtype=torch.float32, device="cuda")
    lib.rowscale_f16(dev(x), x16, rs)
    back = x16.float().cpu().numpy() * rs.cpu().numpy()[:, None]
    amax = np.abs(x).max(axis=1, keepdims=True)
    assert np.max(np.abs(back - x) / amax) <= 2.0 ** -10
    r = rs.cpu().numpy()
    assert np.array_equal(np.log2(r), np.round(np.log2(r)))          # powers of two
    # an all-zero row reports the smallest scale (so `scale / max over rows` ignores it), and stays zero
    xz = x.copy(); xz[3] = 0.0
    lib.rowscale_f16(dev(xz), x16, rs)
    assert float(rs[3]) == 2.0 ** -108 and float(x16[3].float().abs().max()) == 0.0
    assert float(rs.max()) == float(np.delete(r, 3).max())
    # odd width / padded leading dimension (LM-head gradients)
    xo = heavy_tailed((33, 211), 9)
    buf = lib.empty_f16_padded(33, 211, "cuda")
    rs2 = torch.empty(33, dtype=torch.float32, device="cuda")
    lib.rowscale_f16(dev(xo), buf, rs2)
    assert buf.stride(0) == 216
    back = buf.float().cpu().numpy() * rs2.cpu().numpy()[:, None]
    assert np.max(np.abs(back - xo) / np.abs(xo).max(axis=1, keepdims=True)) <= 2.0 ** -10


# --------------------------------------------------------------------------- GEMM
def _ref_gemm(A, B, A2=None, B2=None):
    y = A.double() @ B.double().t()
    if A2 is not None:
        y = y + A2.double() @ B2.double().t()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 64, 64), (256, 512, 192), (300, 200, 72), (1, 8, 8),
                                   (2048, 2304, 768), (2048, 768, 3072), (4096, 4800, 1600), (129, 257, 3080)])
def test_qgemm_matches_fp64(lib, M, N, K):
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda")).half()
    B = (torch.randn(N, K, device="cuda") * 0.1).half()
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.qgemm(A, B, M, N, K, out)
    assert lib.debug_status() == 0, "GEMM pipeline watchdog fired"
    ref = _ref_gemm(A, B)
    err = (out.double() - ref).norm() / ref.norm()
    assert err <= 1e-5, f"rel err {err:.3e}"      # fp32 accumulation: ~sqrt(K) * 2^-24


def test_qgemm_epilogue_and_second_segment(lib):
    torch.manual_seed(1)
    M, N, K, K2 = 515, 776, 320, 64
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    A2 = torch.randn(M, K2, device="cuda").half(); B2 = (torch.randn(N, K2, device="cuda") * 0.1).half()
    rs = torch.rand(M, device="cuda") + 0.5; cs = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda"); C = torch.randn(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    lib.qgemm(A, B, M, N, K, out, A2=A2, B2=B2, K2=K2, alpha=0.5, row_scale=rs, col_scale=cs, bias=bias, clamp_abs=3.0, C=C)
    ref = (_ref_gemm(A, B, A2, B2) * 0.5 * rs.double()[:, None] * cs.double()[None, :]).clamp(-3.0, 3.0) + bias.double() + C.double()
    assert ((out.double() - ref).norm() / ref.norm()) <= 1e-5
    # fp16 output, K2 = 16 (CPT rank), strided operands
    K2 = 16
    A2 = torch.randn(M, 32, device="cuda").half()[:, :K2]; B2 = (torch.randn(N, 32, device="cuda") * 0.1).half()[:, :K2]
    outh = torch.empty(M, N, device="cuda", dtype=torch.float16)
    lib.qgemm(A, B, M, N, K, outh, A2=A2, B2=B2, K2=K2)
    ref = _ref_gemm(A, B, A2, B2)
    assert ((outh.double() - ref).norm() / ref.norm()) <= 1e-3     # fp16 storage rounding
    assert lib.debug_status() == 0


@pytest.mark.parametrize("Mred,I,J,transposed", [(64, 128, 64, False), (1000, 768, 64, False), (4096, 2304, 64, True),
                                                 (333, 200, 136, False), (2048, 768, 768, False)])
def test_gemm_tn_matches_fp64(lib, Mred, I, J, transposed):
    torch.manual_seed(Mred + I)
    P = torch.randn(Mred, I, device="cuda").half(); Q = (torch.randn(Mred, J, device="cuda") * 0.1).half()
    isc = torch.rand(I, device="cuda") + 0.5; jsc = torch.rand(J, device="cuda") + 0.5
    adev = torch.tensor([0.25], device="cuda")
    out = torch.empty((J, I) if transposed else (I, J), device="cuda")
    lib.gemm_tn(P, Q, out, alpha=2.0, alpha_dev=adev, i_scale=isc, j_scale=jsc, transposed_out=transposed)
    assert lib.debug_status() == 0
    ref = (P.double().t() @ Q.double()) * 0.5 * isc.double()[:, None] * jsc.double()[None, :]
    if transposed:
        ref = ref.t()
    assert ((out.double() - ref).norm() / ref.norm()) <= 1e-5
    # deterministic (fixed-order fold of the split planes, no atomics) and the fused STE clamp
    out2 = torch.empty_like(out)
    lib.gemm_tn(P, Q, out2, alpha=2.0, alpha_dev=adev, i_scale=isc, j_scale=jsc, transposed_out=transposed)
    assert torch.equal(out, out2)
    c = float(ref.abs().median())
    lib.gemm_tn(P, Q, out2, alpha=2.0, alpha_dev=adev, i_scale=isc, j_scale=jsc, transposed_out=transposed, clamp_abs=c)
    assert torch.equal(out2, out.clamp(-c, c))


def test_rowscale_max_and_lora_bwd_prep(lib):
    torch.manual_seed(3)
    M, N, r = 777, 2304, 64
    g = torch.randn(M, N, device="cuda") * torch.exp(2 * torch.randn(M, 1, device="cuda")) * 1e-4
    g[5] = 0.0; g[M - 1] = 0.0
    g16 = lib.empty_f16_padded(M, N, "cuda"); eg = torch.empty(M, device="cuda"); gmax = torch.empty(1, device="cuda")
    lib.rowscale_f16_max(g, g16, eg, gmax)
    g16b = lib.empty_f16_padded(M, N, "cuda"); egb = torch.empty(M, device="cuda")
    lib.rowscale_f16(g, g16b, egb)
    assert torch.equal(g16, g16b) and torch.equal(eg, egb) and float(gmax) == float(eg.max())
    assert float(eg[5]) == 2.0 ** -108 and float(gmax) > 2.0 ** -60
    dtn = torch.randn(M, r, device="cuda") * 37.0
    t16 = (torch.randn(M, r, device="cuda") * 5.0).half()
    dt_mul = 2.0 ** -5
    dt16, dt2, t2 = lib.lora_bwd_prep(dtn, t16, eg, gmax, dt_mul)
    e = (eg / gmax).unsqueeze(1)
    assert torch.equal(dt16, (dtn * dt_mul).half())
    assert torch.equal(dt2, (dtn * dt_mul * e).half())
    assert torch.equal(t2, (t16.float() * e).half())
    assert float(dt2[5].abs().max()) == 0.0 and float(t2[5].abs().max()) == 0.0
    a, b, c = lib.lora_bwd_prep(None, t16, eg, gmax, dt_mul)
    assert a is None and b is None and torch.equal(c, t2)


@pytest.mark.parametrize("kind", ["kl", "ce"])
@pytest.mark.parametrize("M,V,T", [(96, 50257, 32), (40, 1000, 8), (33, 211, 11)])
def test_softmax_loss_grad16(lib, kind, M, V, T):
    """Loss value and the fp16 gradient operand (g16 * row_scale) against torch float64."""
    torch.manual_seed(V + M)
    ld = (V + 31) // 32 * 32
    sbuf = torch.randn(M, ld, device="cuda") * 3; tbuf = sbuf + 0.3 * torch.randn(M, ld, device="cuda")
    s, t = sbuf[:, :V], tbuf[:, :V]
    sd = s.double().clone().requires_grad_(True)
    if kind == "kl":
        Tm = 3.0
        row_loss, valid, g16, rs, mx = lib.softmax_loss_grad16(s, "kl", t2d=t, temperature=Tm, seq_len=T)
        keep = (torch.arange(M, device="cuda") % T) != T - 1
        ls = torch.log_softmax(sd / Tm, -1); lt = torch.log_softmax(t.double() / Tm, -1)
        ref_rows = (lt.exp() * (lt - ls)).sum(-1) * keep
        ref_rows.sum().backward()
        ref_grad = sd.grad * Tm                                    # kernel emits d = ps - pt; d(row)/ds = d / T
    else:
        tg = torch.randint(0, V, (M,), device="cuda"); tg[::7] = -100
        row_loss, valid, g16, rs, mx = lib.softmax_loss_grad16(s, "ce", targets=tg)
        keep = tg >= 0
        ref_rows = torch.nn.functional.cross_entropy(sd, tg.clamp_min(0), reduction="none") * keep
        ref_rows.sum().backward()
        ref_grad = sd.grad
        assert torch.equal(valid, keep.float())
    assert ((row_loss.double() - ref_rows).abs().max() <= 1e-5 * max(1.0, float(ref_rows.abs().max())))
    got = g16.double() * rs.double()[:, None]
    assert (got - ref_grad).norm() / ref_grad.norm() <= 6e-4             # fp16 rounding of a row-normalised operand
    assert float(got[~keep].abs().max()) == 0.0 and bool((rs[~keep] == 2.0 ** -108).all())
    assert float(mx) == float(rs.max()) and float(g16.float().abs().max()) <= 256.0
    assert g16.stride(0) % 8 == 0


# --------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,C", [(51, 96), (4096, 768), (300, 1024), (64, 1600), (7, 3072)])
def test_layernorm_against_oracle(lib, rows, C):
    from llm_qat_on_gpt2_b200 import SwitchableLayerNorm
    rng = np.random.default_rng(C)
    x = heavy_tailed((rows, C), 1, zeros=False)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32); b = (0.2 * rng.standard_normal(C)).astype(np.float32)
    gy = rng.standard_normal((rows, C)).astype(np.float32)
    ln = SwitchableLayerNorm(C, precision_levels=[4, 8, 32]).cuda()
    ln.set_precision(8)
    with torch.no_grad():
        ln.weights["8"].copy_(dev(w)); ln.biases["8"].copy_(dev(b))
    xd = dev(x).requires_grad_(True)
    y = ln(xd)
    y.backward(dev(gy))
    y_ref, mean, rstd = switchable_layernorm_forward(x, w, b, 1e-5)
    dx, dw, db = switchable_layernorm_backward(gy, x, w, mean, rstd)
    assert rel_fro(y.detach().cpu().numpy(), y_ref) <= 2e-6
    assert rel_fro(xd.grad.cpu().numpy(), dx) <= 1e-5
    assert rel_fro(ln.weights["8"].grad.cpu().numpy(), dw) <= 1e-5
    assert rel_fro(ln.biases["8"].grad.cpu().numpy(), db) <= 1e-5
    assert ln.weights["4"].grad is None
    with pytest.raises(ValueError):
        ln.set_precision(5)


@pytest.mark.parametrize("rows,C", [(51, 96), (4097, 768), (300, 1024), (65, 1600), (33, 2048)])
@pytest.mark.parametrize("qtype,bits", [("log", 8), ("minmax", 4)])
def test_ln_quantize_act_fused(lib, rows, C, qtype, bits):
    """spq_ln_quantize_act (SwitchableLayerNorm folded into the activation quantiser, p1/models_sp.py:139-147 +
    p1/lora.py:141-149): the normalised rows it can emit agree with the oracle's LayerNorm to fp32 rounding, and the
    operands are BIT-identical to spq_quantize_act applied to those rows -- so the codes are the reference quantiser's
    codes of the kernel's own LayerNorm output."""
    rng = np.random.default_rng(rows + C)
    x = dev(heavy_tailed((rows, C), 2, zeros=False))
    w = dev((1 + 0.2 * rng.standard_normal(C)).astype(np.float32)); b = dev((0.2 * rng.standard_normal(C)).astype(np.float32))
    y_ref, _, _ = switchable_layernorm_forward(x.cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), 1e-5)
    # calibrated parameters of an 8-bit log / 4-bit min-max per-column quantiser on the normalised rows
    st = QuantizerState(num_bits=bits, quantizer_type=qtype, channel_dim=-1, per_channel=True, symmetric=True)
    st.start_calibration(); collect_statistics(st, y_ref); finish_calibration(st)
    sc = dev(np.asarray(st.scale, dtype=np.float32).reshape(-1)); zp = dev(np.broadcast_to(np.asarray(st.zero_point, dtype=np.float32).reshape(-1), sc.shape).copy())
    QT = lib.QTYPE[qtype]
    kind = lib.OPERAND_DEQUANT if qtype == "log" else lib.OPERAND_CODE
    col_mul = dev(np.exp2(rng.integers(-3, 4, C)).astype(np.float32)) if qtype == "log" else None
    raw_mul = dev(np.exp2(rng.integers(-2, 3, C)).astype(np.float32))
    y = torch.empty_like(x)
    a_q = torch.empty((rows, C), dtype=torch.float16, device="cuda"); a_raw = torch.empty_like(a_q)
    lib.ln_quantize_act(x, w, b, 1e-5, sc, zp, lib.PER_COL, QT, bits, True, kind, col_mul, 1.0, a_q, a_raw, raw_mul, y_out=y)
    assert rel_fro(y.cpu().numpy(), y_ref) <= 2e-6
    want_q = torch.empty_like(a_q); want_raw = torch.empty_like(a_raw)
    lib.quantize_act(y, sc, zp, lib.PER_COL, QT, bits, True, kind, col_mul, 1.0, want_q, want_raw, raw_mul)
    assert torch.equal(a_q, want_q) and torch.equal(a_raw, want_raw)
    # without the optional float32 copy, and without the raw (LoRA) operand
    a_q2 = torch.empty_like(a_q)
    lib.ln_quantize_act(x, w, b, 1e-5, sc, zp, lib.PER_COL, QT, bits, True, kind, col_mul, 1.0, a_q2, None, None)
    assert torch.equal(a_q2, want_q)


@pytest.mark.parametrize("rows,C", [(51, 96), (4097, 768), (300, 1024), (65, 1600)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_ln_rowscale_stats_fused(lib, rows, C, mode):
    """spq_ln_rowscale_stats (calibration pass / 32-bit path behind a LayerNorm): operand and row scales BIT-identical to
    spq_rowscale_f16, per-column statistics (mode 1: min-max quantisers, mode 2: log quantisers) BIT-identical to
    spq_minmax_stats, both applied to the normalised rows the kernel can emit; first batch and accumulation."""
    rng = np.random.default_rng(rows * 3 + C)
    w = dev((1 + 0.2 * rng.standard_normal(C)).astype(np.float32)); b = dev((0.2 * rng.standard_normal(C)).astype(np.float32))
    smin = torch.empty(C, device="cuda"); smax = torch.empty(C, device="cuda"); state = torch.zeros(1, dtype=torch.int32, device="cuda")
    wmin = torch.empty(C, device="cuda"); wmax = torch.empty(C, device="cuda"); wstate = torch.zeros(1, dtype=torch.int32, device="cuda")
    for batch in range(2):
        x = dev(heavy_tailed((rows, C), 5 + batch, zeros=False))
        if batch == 1:
            x[3].zero_()                                    # a constant row: LayerNorm output = bias
        y = torch.empty_like(x)
        x16 = torch.empty((rows, C), dtype=torch.float16, device="cuda"); rs = torch.empty(rows, device="cuda")
        lib.ln_rowscale_stats(x, w, b, 1e-5, x16, rs, stats_mode=mode, stat_eps=1e-5, stat_min=smin if mode else None,
                              stat_max=smax if mode else None, accumulate=batch > 0, state=state if mode else None, y_out=y)
        y_ref, _, _ = switchable_layernorm_forward(x.cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), 1e-5)
        assert rel_fro(y.cpu().numpy(), y_ref) <= 2e-6
        want16 = torch.empty_like(x16); wrs = torch.empty_like(rs)
        lib.rowscale_f16(y, want16, wrs)
        assert torch.equal(x16, want16) and torch.equal(rs, wrs)
        if mode:
            lib.minmax_stats(y, lib.PER_COL, mode == 2, 1e-5, wmin, wmax, accumulate=batch > 0, state=wstate)
            assert torch.equal(smin, wmin) and torch.equal(smax, wmax) and torch.equal(state, wstate)


def test_layernorm_golden(lib):
    from llm_qat_on_gpt2_b200 import SwitchableLayerNorm
    g = np.load(os.path.join(GOLDEN, "layernorm.npz"))
    ln = SwitchableLayerNorm(96, precision_levels=[4, 8, 32]).cuda()
    for p in (4, 8, 32):
        with torch.no_grad():
            ln.weights[str(p)].copy_(dev(g[f"w{p}"])); ln.biases[str(p)].copy_(dev(g[f"b{p}"]))
    for p in (4, 8, 32):
        ln.set_precision(p); ln.zero_grad()
        x = dev(g["x"]).requires_grad_(True)
        y = ln(x); y.backward(dev(g["grad_y"]))
        assert rel_fro(y.detach().cpu().numpy(), g[f"y{p}"]) <= 2e-6
        assert rel_fro(x.grad.cpu().numpy(), g[f"gx{p}"]) <= 1e-5
        assert rel_fro(ln.weights[str(p)].grad.cpu().numpy(), g[f"gw{p}"]) <= 1e-5
        assert rel_fro(ln.biases[str(p)].grad.cpu().numpy(), g[f"gb{p}"]) <= 1e-5


def test_ste_backward(lib):
    from llm_qat_on_gpt2_b200 import apply_log_quantization, apply_minmax_quantization
    x = dev(heavy_tailed((33, 40), 2)).requires_grad_(True)
    g = torch.randn(33, 40, device="cuda") * 8
    s = torch.full((1, 40), 0.1, device="cuda"); z = torch.zeros(1, 40, device="cuda")
    apply_minmax_quantization(x, s, z, 4, True).backward(g)
    assert torch.equal(x.grad, g)
    x.grad = None
    apply_log_quantization(x, torch.full((1, 40), -10.0, device="cuda"), torch.full((1, 40), 15.0, device="cuda"), 8, True).backward(g)
    assert torch.equal(x.grad, g.clamp(-10, 10))


def test_qgemm_gelu_epilogue(lib):
    torch.manual_seed(5)
    M, N, K = 300, 520, 192
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda"); cs = torch.rand(N, device="cuda") + 0.5
    out = torch.empty(M, N, device="cuda")
    lib.qgemm(A, B, M, N, K, out, col_scale=cs, bias=bias, activation=1)
    ref = torch.nn.functional.gelu(((A.double() @ B.double().t()) * cs.double() + bias.double()).float())
    assert ((out.double() - ref.double()).norm() / ref.double().norm()) <= 1e-5
    # padded output rows (odd N): view of a wider buffer, TMA store clips at N
    N2 = 211
    B2 = (torch.randn(N2, K, device="cuda") * 0.1).half()
    buf = torch.full((M, 212), 7.0, device="cuda")
    lib.qgemm(A, B2, M, N2, K, buf[:, :N2])
    ref = (A.double() @ B2.double().t())
    assert ((buf[:, :N2].double() - ref).norm() / ref.norm()) <= 1e-5
    # documented contract: with N % 4 != 0 and padded rows, the padding floats may be overwritten
    dense = torch.full((M, N2), 7.0, device="cuda")      # no padding available -> masked store path
    lib.qgemm(A, B2, M, N2, K, dense)
    assert ((dense.double() - ref).norm() / ref.norm()) <= 1e-5
    assert lib.debug_status() == 0


@pytest.mark.parametrize("M,V,ld", [(64, 211, 212), (37, 50257, 50260), (128, 1000, 1000), (5, 33, 33)])
def test_cross_entropy_fwd(lib, M, V, ld):
    torch.manual_seed(V)
    buf = torch.randn(M, ld, device="cuda") * 5
    logits = buf[:, :V]
    tg = torch.randint(0, V, (M,), device="cuda")
    tg[::7] = -100
    got = lib.cross_entropy_fwd(logits, tg)
    want = torch.nn.functional.cross_entropy(logits.double(), tg, ignore_index=-100)
    assert abs(got.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4096, 768), (515, 3072), (33, 64), (7, 12)])
def test_half_inputs_equal_widened_inputs(shape):
    """float16 inputs to the activation-side kernels (stats, row scale, quantise) give the bits of x.float()."""
    from llm_qat_on_gpt2_b200 import _lib
    M, K = shape
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(M + K)
    xh = (torch.randn(M, K, generator=g) * 3).half().to(dev)
    xf = xh.float()
    for log_mode in (False, True):
        for bcast, n in ((_lib.PER_COL, K), (_lib.PER_TENSOR, 1)):
            outs = []
            for x in (xf, xh):
                mn = torch.empty(n, device=dev); mx = torch.empty(n, device=dev)
                st = torch.zeros(1, dtype=torch.int32, device=dev)
                _lib.minmax_stats(x, bcast, log_mode, 1e-5, mn, mx, accumulate=False, state=st)
                outs.append((mn, mx))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    if K % 4 == 0:
        res = []
        for x in (xf, xh):
            o = torch.empty(M, K, dtype=torch.float16, device=dev); rs = torch.empty(M, device=dev)
            _lib.rowscale_f16(x, o, rs)
            res.append((o, rs))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        for qtype, bits, kind in ((_lib.MINMAX, 4, _lib.OPERAND_CODE), (_lib.LOG, 8, _lib.OPERAND_DEQUANT)):
            if qtype == _lib.MINMAX:
                sc = (xf.abs().amax(0) / 7).clamp_min(1e-5); zp = torch.zeros(K, device=dev)
            else:
                lg = torch.log2(xf.abs().clamp_min(1e-5))
                zp = lg.amin(0); sc = (lg.amax(0) - zp).clamp_min(1e-5)
            cm = torch.full((K,), 0.25, device=dev); rm = torch.full((K,), 0.5, device=dev)
            res = []
            for x in (xf, xh):
                a_q = torch.empty(M, K, dtype=torch.float16, device=dev); a_raw = torch.empty(M, K, dtype=torch.float16, device=dev)
                _lib.quantize_act(x, sc, zp, _lib.PER_COL, qtype, bits, True, kind, cm, 1.0, a_q, a_raw, rm)
                res.append((a_q, a_raw))
            assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(515, 768, 320), (1024, 2304, 768), (300, 200, 64), (130, 72, 128), (4096, 3072, 256)])
def test_qgemm_tma_store_half_and_residual(M, N, K):
    """TMA-store epilogue variants: float16 D (64B-swizzled staging) and a prefetched float32 residual C,
    including the in-place form D = C (residual stream update) and ragged edge tiles."""
    from llm_qat_on_gpt2_b200 import _lib as lib
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    cs = torch.rand(N, device="cuda") + 0.5; bias = torch.randn(N, device="cuda"); rs = torch.rand(M, device="cuda") + 0.5
    base = (A.double() @ B.double().t()) * rs.double()[:, None] * cs.double()[None, :] + bias.double()
    # float16 output
    outh = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    lib.qgemm(A, B, M, N, K, outh, row_scale=rs, col_scale=cs, bias=bias)
    assert ((outh.double() - base).norm() / base.norm()) <= 1e-3
    assert (outh.float() - base.float()).abs().max() <= 2e-3 * base.abs().max()
    # float16 output + GELU
    lib.qgemm(A, B, M, N, K, outh, row_scale=rs, col_scale=cs, bias=bias, activation=1)
    refg = torch.nn.functional.gelu(base.float()).double()
    assert ((outh.double() - refg).norm() / refg.norm()) <= 1e-3
    # residual, separate output
    C = torch.randn(M, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.qgemm(A, B, M, N, K, out, row_scale=rs, col_scale=cs, bias=bias, C=C)
    ref = base + C.double()
    assert ((out.double() - ref).norm() / ref.norm()) <= 1e-5
    # residual, in place (D aliases C)
    D = C.clone()
    lib.qgemm(A, B, M, N, K, D, row_scale=rs, col_scale=cs, bias=bias, C=D)
    assert torch.equal(D, out)
    # residual + float16 output, row-padded residual view (ldc != N)
    Cp = torch.randn(M, N + 4, device="cuda")
    lib.qgemm(A, B, M, N, K, outh, row_scale=rs, col_scale=cs, bias=bias, C=Cp[:, :N])
    ref = base + Cp[:, :N].double()
    assert ((outh.double() - ref).norm() / ref.norm()) <= 1e-3
    assert lib.debug_status() == 0


@pytest.mark.gpu
def test_qgemm_general_epilogue_still_covers_unaligned():
    """Shapes the TMA store cannot take (N % 4 != 0 with a residual, odd leading dimensions) use the masked path."""
    from llm_qat_on_gpt2_b200 import _lib as lib
    torch.manual_seed(9)
    M, N, K = 333, 203, 128
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    C = torch.randn(M, N, device="cuda")
    ref = A.double() @ B.double().t() + C.double()
    out = torch.empty(M, N, device="cuda")
    lib.qgemm(A, B, M, N, K, out, C=C)
    assert ((out.double() - ref).norm() / ref.norm()) <= 1e-5
    outh = torch.empty(M, N, device="cuda", dtype=torch.float16)
    lib.qgemm(A, B, M, N, K, outh, C=C)
    assert ((outh.double() - ref).norm() / ref.norm()) <= 1e-3
    assert lib.debug_status() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(2693, 8200, 64), (4100, 12001, 128)])
def test_qgemm_grouped_tile_order(M, N, K):
    """>= 32 column tiles switch the persistent CTAs to the grouped (16 row-blocks) sweep; the last group is short."""
    from llm_qat_on_gpt2_b200 import _lib as lib
    torch.manual_seed(N)
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") * 0.1).half()
    ld = (N + 31) // 32 * 32
    buf = torch.full((M, ld), float("nan"), device="cuda")
    lib.qgemm(A, B, M, N, K, buf[:, :N] if ld != N else buf)
    ref = A.double() @ B.double().t()
    out = buf[:, :N].double()
    assert not torch.isnan(out).any()
    assert ((out - ref).norm() / ref.norm()) <= 1e-5
    assert lib.debug_status() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("M,V,K", [(300, 211, 64), (1024, 50257, 128), (130, 1000, 64), (4096, 2304, 64)])
def test_qgemm_lse_and_cross_entropy_from_parts(M, V, K):
    """LM head with the log-sum-exp folded into the epilogue: same logits as spq_qgemm bit for bit, loss equal to
    float64 cross-entropy of those logits (ignore_index rows excluded, ragged last column tile masked)."""
    from llm_qat_on_gpt2_b200 import _lib as lib
    torch.manual_seed(V)
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(V, K, device="cuda") * 0.3).half()
    rs = torch.rand(M, device="cuda") + 0.5; cs = torch.rand(V, device="cuda") + 0.5
    ld = V if V % 4 == 0 else (V + 31) // 32 * 32
    buf = torch.full((M, ld), 1e30, device="cuda"); ref_buf = torch.full((M, ld), 1e30, device="cuda")
    logits = buf[:, :V] if ld != V else buf
    ref_logits = ref_buf[:, :V] if ld != V else ref_buf
    parts = lib.qgemm_lse(A, B, M, V, K, logits, row_scale=rs, col_scale=cs)
    lib.qgemm(A, B, M, V, K, ref_logits, row_scale=rs, col_scale=cs)
    assert torch.equal(logits, ref_logits)
    tg = torch.randint(0, V, (M,), device="cuda")
    tg[::5] = -100
    got = lib.cross_entropy_from_parts(parts, logits, tg)
    want = torch.nn.functional.cross_entropy(logits.double(), tg, ignore_index=-100)
    assert abs(got.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
    # per-row log-sum-exp from the partials
    m = parts[..., 0]; s_ = parts[..., 1]
    mm = m.max(dim=1).values
    lse = mm + torch.log((s_ * torch.exp(m - mm[:, None])).sum(dim=1))
    assert float((lse.double() - torch.logsumexp(logits.double(), dim=1)).abs().max()) <= 1e-4
    assert lib.debug_status() == 0


@pytest.mark.parametrize("M,N", [(257, 3072), (64, 96), (1000, 768), (33, 4096)])
def test_rowscale_dgelu_fused(lib, M, N):
    """spq_rowscale_dgelu_f16_max: the fp16 gradient operand of g * gelu'(y) (exact erf GELU) in one pass, against torch's
    gelu_backward followed by spq_rowscale_f16_max: row scales identical (a row's absmax can only move across a power of
    two by the last ulp of gelu'), values within fp16 rounding of the product."""
    torch.manual_seed(M + N)
    y = (torch.randn(M, N, device="cuda") * 2).requires_grad_(True)
    g = torch.randn(M, N, device="cuda") * 1e-3
    g[5].zero_()                                             # a row without gradient
    (dy,) = torch.autograd.grad(torch.nn.functional.gelu(y), y, g)
    want16 = torch.empty((M, N), dtype=torch.float16, device="cuda"); wrs = torch.empty(M, device="cuda"); wmx = torch.empty(1, device="cuda")
    lib.rowscale_f16_max(dy.contiguous(), want16, wrs, wmx)
    got16 = torch.empty_like(want16); rs = torch.empty_like(wrs); mx = torch.empty_like(wmx)
    lib.rowscale_dgelu_f16_max(g, y.detach(), got16, rs, mx)
    same = rs == wrs
    assert float(same.float().mean()) >= 0.99 and torch.equal(mx, wmx) or float((mx / wmx).item()) in (0.5, 1.0, 2.0)
    a = got16.float() * rs[:, None]; b = want16.float() * wrs[:, None]
    assert float((a - b).norm() / b.norm()) <= 1e-3
    assert float((a - dy).abs().max()) <= 2e-3 * float(dy.abs().max())
    assert torch.equal(got16[5], torch.zeros_like(got16[5]))


def test_mse_select_matches_torch(lib):
    """spq_mse_select (feature term of the distillation loss, p1/distillation_manager.py:82-116): the pair is chosen by an
    index read on the device; every index, an odd element count and out-of-range indices (clamped); run-to-run identical."""
    torch.manual_seed(0)
    for numel in (8 * 32 * 64, 4099):
        a = [torch.randn(numel, device="cuda") for _ in range(13)]
        b = [torch.randn(numel, device="cuda") * 0.5 for _ in range(13)]
        if numel % 4:
            a = [torch.randn(numel + 4, device="cuda")[:numel] for _ in range(13)]     # (still 16-byte aligned views)
        sel = torch.zeros(1, dtype=torch.int32, device="cuda")
        for l in list(range(13)) + [-3, 99]:
            sel.fill_(l)
            got = lib.mse_select(a, b, sel)
            ll = min(max(l, 0), 12)
            want = torch.nn.functional.mse_loss(a[ll].double(), b[ll].double())
            assert abs(got.item() - want.item()) <= 1e-6 * want.item(), (numel, l)
            again = lib.mse_select(a, b, sel)
            assert torch.equal(got, again)


# --------------------------------------------------------------------------- fp8 (e4m3) integer-code GEMM
@pytest.mark.parametrize("M,N,K,K2", [(300, 200, 256, 0), (4096, 2304, 768, 64), (1000, 768, 3072, 64), (129, 4800, 1600, 0)])
def test_qgemm_f8_integer_codes_exact(lib, M, N, K, K2):
    """tcgen05.mma.kind::f8f6f4 on e4m3 integer codes in [-7, 7]: the fp32 accumulator holds the integer dot product
    exactly (K * 49 < 2^24), so the result equals the int64 matmul bit for bit; with the fp16 LoRA segment riding in
    the same accumulator (kind::f16) it matches float64 to fp32 rounding."""
    torch.manual_seed(M + K)
    ca = torch.randint(-7, 8, (M, K), device="cuda")
    cb = torch.randint(-7, 8, (N, K), device="cuda")
    A8 = ca.float().to(torch.float8_e4m3fn).view(torch.uint8)
    B8 = cb.float().to(torch.float8_e4m3fn).view(torch.uint8)
    exact = (ca.double() @ cb.double().t())
    out = torch.empty(M, N, device="cuda")
    lib.qgemm_f8(A8, B8, M, N, K, out)
    assert lib.debug_status() == 0
    assert torch.equal(out.double(), exact)
    cs = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    if K2:
        A2 = torch.randn(M, K2, device="cuda").half(); B2 = (torch.randn(N, K2, device="cuda") * 3).half()
        lib.qgemm_f8(A8, B8, M, N, K, out, A2=A2, B2=B2, K2=K2, col_scale=cs, bias=bias)
        ref = (exact + A2.double() @ B2.double().t()) * cs.double()[None, :] + bias.double()[None, :]
    else:
        lib.qgemm_f8(A8, B8, M, N, K, out, col_scale=cs, bias=bias)
        ref = exact * cs.double()[None, :] + bias.double()[None, :]
    assert ((out.double() - ref).norm() / ref.norm()) <= 2e-7
    outh = torch.empty(M, N, device="cuda", dtype=torch.float16)
    lib.qgemm_f8(A8, B8, M, N, K, outh, col_scale=cs * 1e-2)
    assert ((outh.double() - exact * cs.double()[None, :] * 1e-2).norm() / (exact * 1e-2).norm()) <= 1e-3
    assert lib.debug_status() == 0
