"""Full-size parity (BASELINE.json configs[1]: 32 x 1024 tokens, GPT-2 small widths) through properties that do
not need the CPU oracle at that size: exact order statistics, idempotence, monotonicity, code range, tile
independence and linearity of the GEMM, and a float64 spot check of sampled rows.  Everything goes through
the C ABI.  Run on the B200 box: pytest -m gpu."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOKENS = 32 * 1024          # per-GPU batch of the benchmark workload


@pytest.fixture(scope="module")
def lib():
    from llm_qat_on_gpt2_b200 import _lib
    _lib.load_library()
    return _lib


def _acts(rows, cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(rows, cols, device="cuda", generator=g)
    x *= torch.exp(0.7 * torch.randn(1, cols, device="cuda", generator=g))      # per-channel spread
    x[:, 5] *= 30.0                                                              # an outlier channel (GPT-2 has them)
    x[::97, 11] = 0.0
    return x


@pytest.mark.parametrize("cols", [768, 3072])
@pytest.mark.parametrize("log_mode", [False, True])
def test_fullsize_statistics_are_exact_order_statistics(lib, cols, log_mode):
    """min/max are order independent: the kernel must return exactly torch's amin/amax (of x, or of
    log2(clamp(|x|, eps)) -- log2 is monotone, so the extremes are those of |x|)."""
    x = _acts(TOKENS, cols, 1)
    mn = torch.empty(cols, device="cuda"); mx = torch.empty(cols, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.minmax_stats(x, lib.PER_COL, log_mode, 1e-5, mn, mx, accumulate=False, state=st)
    if not log_mode:
        assert torch.equal(mn, x.amin(0)) and torch.equal(mx, x.amax(0))
    else:
        a = x.abs().clamp_min(1e-5)
        lo, hi = a.amin(0), a.amax(0)
        # correctly rounded float32 log2 of the extreme values (float64 log2, rounded once)
        assert torch.equal(mn, torch.log2(lo.double()).float()) and torch.equal(mx, torch.log2(hi.double()).float())
    # two half batches accumulate to the same statistics as the whole batch
    mn2 = torch.empty(cols, device="cuda"); mx2 = torch.empty(cols, device="cuda")
    lib.minmax_stats(x[: TOKENS // 2], lib.PER_COL, log_mode, 1e-5, mn2, mx2, accumulate=False, state=st)
    lib.minmax_stats(x[TOKENS // 2:], lib.PER_COL, log_mode, 1e-5, mn2, mx2, accumulate=True, state=st)
    assert torch.equal(mn, mn2) and torch.equal(mx, mx2)


@pytest.mark.parametrize("cols", [768, 3072])
@pytest.mark.parametrize("bits,symmetric", [(8, True), (4, True), (8, False)])
def test_fullsize_minmax_quantiser_properties(lib, cols, bits, symmetric):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, quantize_codes
    x = _acts(TOKENS, cols, 2)
    q = LearnableFakeQuantize(bits, channel_dim=-1, quantizer_type="minmax", symmetric=symmetric).cuda()
    q.start_calibration(); q(x[: TOKENS // 2]); q.finish_calibration()          # second half clips
    out, codes, _ = quantize_codes(x, q.scale, q.zero_point, bits, symmetric, "minmax")
    qmin, qmax = (-(2 ** (bits - 1)) + 1, 2 ** (bits - 1) - 1) if symmetric else (0, 2 ** bits - 1)
    assert int(codes.min()) >= qmin and int(codes.max()) <= qmax
    assert int(codes.max()) == qmax                                             # the outlier channel saturates
    # dequantised value is exactly (code - zero_point) * scale in float32
    zp = q.zero_point.reshape(1, -1) if not symmetric else 0.0
    assert torch.equal(out, (codes.float() - zp) * q.scale.reshape(1, -1))
    # idempotence: quantising a dequantised tensor changes nothing
    out2, codes2, _ = quantize_codes(out, q.scale, q.zero_point, bits, symmetric, "minmax")
    assert torch.equal(codes2, codes) and torch.equal(out2, out)
    # monotone per channel: sorted inputs give non-decreasing codes
    xs, _ = torch.sort(x[:8192], dim=0)
    _, cs, _ = quantize_codes(xs.contiguous(), q.scale, q.zero_point, bits, symmetric, "minmax")
    assert bool((cs[1:] >= cs[:-1]).all())
    # module forward == the codes path
    assert torch.equal(q(x), out)


@pytest.mark.parametrize("cols", [768, 3072])
def test_fullsize_log_quantiser_properties(lib, cols):
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, quantize_codes
    bits = 8
    x = _acts(TOKENS, cols, 3)
    q = LearnableFakeQuantize(bits, channel_dim=-1, quantizer_type="log", symmetric=True).cuda()
    q.start_calibration(); q(x); q.finish_calibration()
    out, level, sign = quantize_codes(x, q.scale, q.zero_point, bits, True, "log")
    n_levels = 2 ** (bits - 1) - 1                  # symmetric: centred levels -n..n over the calibrated log range
    assert int(level.min()) >= -n_levels and int(level.max()) <= n_levels
    assert int(level.min()) == -n_levels and int(level.max()) == n_levels       # calibrated on x itself
    assert torch.equal(sign == 0, x.abs() < 1e-5)                               # zero mask == |x| < eps
    assert torch.equal(torch.sign(out), torch.sign(x) * (sign != 0))
    # a level step is log_range / (2 n) in log2: every non-zero value is within half a step of its input
    step = (q.scale.reshape(1, -1) / (2 * n_levels))
    nz = sign != 0
    err = (torch.log2(out.abs().clamp_min(1e-30)) - torch.log2(x.abs().clamp_min(1e-5))).abs()
    assert float((err[nz] - 0.5 * step.expand_as(err)[nz]).max()) <= 1e-4
    # monotone in |x| per channel
    xs, _ = torch.sort(x[:8192].abs(), dim=0)
    _, ls, ss = quantize_codes(xs.contiguous(), q.scale, q.zero_point, bits, True, "log")
    lv = torch.where(ss != 0, ls, torch.full_like(ls, -n_levels - 1))
    assert bool((lv[1:] >= lv[:-1]).all())
    # idempotence on the level index: re-quantising the dequantised values gives the same levels
    _, level2, sign2 = quantize_codes(out, q.scale, q.zero_point, bits, True, "log")
    same = (level2 == level) | (sign == 0)
    # a value exactly on a level sits on a rounding tie only at the two ends; allow a vanishing fraction
    assert float((~same).float().mean()) <= 1e-6
    assert torch.equal(q(x), out)


def test_fullsize_gemm_tiles_are_independent_and_linear(lib):
    """32768 x 2304 x 768 (c_attn): column blocks computed alone equal the same columns of the full product bit
    for bit; the product is additive over a K split within fp32 rounding; sampled rows agree with float64."""
    M, N, K = TOKENS, 2304, 768
    g = torch.Generator(device="cuda").manual_seed(4)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    cs = torch.rand(N, device="cuda", generator=g) + 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    full = torch.empty(M, N, device="cuda")
    lib.qgemm(A, B, M, N, K, full, col_scale=cs, bias=bias)
    part = torch.empty(M, 768, device="cuda")
    for j in range(3):
        sl = slice(768 * j, 768 * (j + 1))
        lib.qgemm(A, B[sl].contiguous(), M, 768, K, part, col_scale=cs[sl].contiguous(), bias=bias[sl].contiguous())
        assert torch.equal(part, full[:, sl])
    rows = torch.arange(0, M, 257, device="cuda")
    ref = (A[rows].double() @ B.double().t()) * cs.double() + bias.double()
    assert float((full[rows].double() - ref).norm() / ref.norm()) <= 1e-5
    # K split through the second operand segment: same accumulator, so only the summation order differs
    two = torch.empty(M, N, device="cuda")
    lib.qgemm(A[:, :512], B[:, :512], M, N, 512, two, A2=A[:, 512:], B2=B[:, 512:], K2=256, col_scale=cs, bias=bias)
    assert float((two - full).abs().max()) <= 1e-5 * float(full.abs().max())
    # fp16 output is the correctly rounded (saturating) fp16 of the fp32 output
    outh = torch.empty(M, N, device="cuda", dtype=torch.float16)
    lib.qgemm(A, B, M, N, K, outh, col_scale=cs, bias=bias)
    assert torch.equal(outh, full.clamp(-65504, 65504).half())
    # the residual epilogue is the separate add
    C = torch.randn(M, N, device="cuda", generator=g)
    withc = torch.empty(M, N, device="cuda")
    lib.qgemm(A, B, M, N, K, withc, col_scale=cs, bias=bias, C=C)
    assert torch.equal(withc, full + C)
    assert lib.debug_status() == 0


def test_fullsize_activation_operands_match_codes(lib):
    """spq_quantize_act at 32768 x 3072 (c_proj input): the fp16 GEMM operand holds exactly the integer codes
    (min-max) and the raw operand is x times its per-channel power of two, saturating."""
    from llm_qat_on_gpt2_b200 import LearnableFakeQuantize, quantize_codes
    M, K = TOKENS, 3072
    x = _acts(M, K, 5)
    q = LearnableFakeQuantize(8, channel_dim=-1, quantizer_type="minmax", symmetric=True).cuda()
    q.start_calibration(); q(x); q.finish_calibration()
    _, codes, _ = quantize_codes(x, q.scale, q.zero_point, 8, True, "minmax")
    sc = q.scale.reshape(-1).contiguous(); zp = q.zero_point.reshape(-1).contiguous()
    raw_mul = torch.exp2(torch.floor(3 - torch.log2(x.abs().amax(0))))          # maps the bound into (4, 8]
    a_q = torch.empty(M, K, device="cuda", dtype=torch.float16); a_raw = torch.empty_like(a_q)
    lib.quantize_act(x, sc, zp, lib.PER_COL, lib.MINMAX, 8, True, lib.OPERAND_CODE, None, 1.0, a_q, a_raw, raw_mul)
    assert torch.equal(a_q.int(), codes)
    assert torch.equal(a_raw, (x * raw_mul).half())
