// Quantise kernels (HBM-bound): fake-quant dequant / integer codes / fp16 GEMM operands.
//
// Replaces MinMaxQuantizationFunction.forward and LogQuantizationFunction.forward of the
// reference (p1/quantization_methods.py:8-22, 33-79; ~4 resp. ~25 eager kernels with full-size
// temporaries) by one pass: 4 B read per element, 2 B (operand) and/or 4 B (dequant) written.
//
// Bit-exactness: the arithmetic below is the reference's, operation by operation, with explicit
// round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an FMA or replace the
// division by a reciprocal multiply.  log2 on the fast path is CUDA's 1-ulp log2f; whenever the
// pre-rounding level lies within the error band of a rounding tie the element is re-evaluated
// with the correctly rounded logarithm (log2_cr, via double), which is the oracle's definition,
// so the level index is exact everywhere while the double-precision unit is touched by < 0.1 %
// of the elements at 8 bits.
#include "spq_common.cuh"

namespace spq {
namespace quant {

constexpr float LOG_EPS = 1e-5f;   // p1/quantization_methods.py:35 (hard-coded)

struct QParams {
    int bits;
    int symmetric;
    float n_sym;    // 2^(b-1) - 1
    float full;     // 2^b - 1
};

// ---------------------------------------------------------------- element arithmetic
struct MinMaxOut { float dq; float code; float centered; };

__device__ __forceinline__ MinMaxOut minmax_elem(float x, float s, float zp, const QParams& qp) {
    MinMaxOut o;
    if (qp.symmetric) {                                   // :13-16
        float q = rintf(__fdiv_rn(x, s));
        q = fminf(fmaxf(q, -qp.n_sym), qp.n_sym);
        o.code = q;
        o.centered = q;
        o.dq = __fmul_rn(q, s);
    } else {                                              // :17-20
        float q = rintf(__fadd_rn(__fdiv_rn(x, s), zp));
        q = fminf(fmaxf(q, 0.f), qp.full);
        o.code = q;
        o.centered = __fsub_rn(q, zp);
        o.dq = __fmul_rn(o.centered, s);
    }
    return o;
}

struct LogOut { float dq; float level; float sign; };

// v(l): the value the reference rounds (p1/quantization_methods.py:43-60), as a function of l = log2|x|
__device__ __forceinline__ float log_prelevel(float l, float log_min, float range_c, const QParams& qp) {
    float ln = __fdiv_rn(__fsub_rn(l, log_min), range_c);
    ln = fminf(fmaxf(ln, 0.f), 1.f);
    if (qp.symmetric) return __fmul_rn(__fmul_rn(__fsub_rn(ln, 0.5f), 2.f), qp.n_sym);
    return __fmul_rn(ln, qp.full);
}

// Per-channel constants of the log quantiser that do not depend on the element.
struct LogCh {
    float log_min, log_range, range_c, inv_range;
};
__device__ __forceinline__ LogCh make_logch(float log_min, float log_range) {
    LogCh c;
    c.log_min = log_min;
    c.log_range = log_range;
    c.range_c = (log_range < LOG_EPS) ? LOG_EPS : log_range;          // :43 clamp(min=eps)
    c.inv_range = __frcp_rn(c.range_c);
    return c;
}

// Level index: bit-exact w.r.t. the reference arithmetic (correctly rounded log2, IEEE division).
// Fast path: lg2.approx + reciprocal multiply + one FMA; its result can differ from the exact
// pre-rounding value by at most `band`, so whenever it lies that close to a rounding tie the
// element is re-evaluated with the exact sequence.  ~3e-4 of the elements at 8 bits.
__device__ __forceinline__ LogOut log_elem(float x, const LogCh& ch, const QParams& qp) {
    LogOut o;
    const float ax = fabsf(x);
    const bool zero = ax < LOG_EPS;                                      // :36
    o.sign = zero ? 0.f : ((x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f));
    const float axc = (ax < LOG_EPS) ? LOG_EPS : ax;                     // :40 clamp(min=eps), >= 1e-5: never denormal
    const float nl = qp.symmetric ? qp.n_sym : qp.full;
    const float lev_mul = qp.symmetric ? 2.f * nl : nl;

    float l = __log2f(axc);                                              // |error| < 2^-20 on this range
    float ln = fminf(fmaxf((l - ch.log_min) * ch.inv_range, 0.f), 1.f);
    float v = qp.symmetric ? fmaf(ln, lev_mul, -nl) : ln * lev_mul;
    float r = rintf(v);
    const float tie_dist = fabsf(fabsf(v - r) - 0.5f);
    const float band = lev_mul * (fmaf(fabsf(l), 4.76837158203125e-07f, 1.9073486328125e-06f) * ch.inv_range + 9.5367431640625e-07f) +
                       fabsf(v) * 4.76837158203125e-07f + 1e-9f;
    if (tie_dist <= band) {                                              // rare: the reference's exact sequence
        l = log2_cr(axc);
        v = log_prelevel(l, ch.log_min, ch.range_c, qp);
        r = rintf(v);
    }
    float qn;
    if (qp.symmetric) {                                                  // :50-56, :63 (value only: tolerance-level)
        r = fminf(fmaxf(r, -nl), nl);
        qn = fmaf(r, __frcp_rn(lev_mul), 0.5f);
    } else {                                                             // :57-60, :65
        r = fminf(fmaxf(r, 0.f), nl);
        qn = r * __frcp_rn(nl);
    }
    o.level = r;
    const float mag = exp2f(fmaf(qn, ch.log_range, ch.log_min));         // :67-69
    o.dq = zero ? 0.f : mag * o.sign;                                    // :71-74
    return o;
}

__device__ __forceinline__ float bparam(const float* p, int bcast, long long row, long long col) {
    return bcast == SPQ_PER_TENSOR ? __ldg(p) : (bcast == SPQ_PER_ROW ? __ldg(p + row) : __ldg(p + col));
}

// ---------------------------------------------------------------- general elementwise kernel
struct FqArgs {
    const float* x;
    long long rows, cols;
    const float* scale;
    const float* zp;
    int bcast;
    QParams qp;
    float* dequant;
    int32_t* codes;
    int8_t* sign;
    unsigned short* operand;
    int operand_kind;
    const float* row_mul;
    const float* col_mul;
    float mul;
    int transposed;
    long long op_ld;      // leading dimension of the operand output (elements)
};

// block (32, 8): x -> 4 consecutive columns (VEC) or 1; y -> rows; grid-stride over row tiles.
template <int QTYPE, int VEC>
__global__ void __launch_bounds__(256)
fake_quantize_kernel(FqArgs a) {
    const long long c0 = (static_cast<long long>(blockIdx.x) * 32 + threadIdx.x) * VEC;
    if (c0 >= a.cols) return;
    float s[VEC], z[VEC], cm[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const long long c = (c0 + j < a.cols) ? c0 + j : c0;
        s[j] = (a.bcast == SPQ_PER_ROW) ? 0.f : bparam(a.scale, a.bcast, 0, c);
        z[j] = (a.bcast == SPQ_PER_ROW) ? 0.f : bparam(a.zp, a.bcast, 0, c);
        cm[j] = a.col_mul ? __ldg(a.col_mul + c) : 1.f;
    }
    for (long long r = static_cast<long long>(blockIdx.y) * 8 + threadIdx.y; r < a.rows; r += static_cast<long long>(gridDim.y) * 8) {
        float xv[VEC];
        const float* px = a.x + r * a.cols + c0;
        if constexpr (VEC == 4) {
            const float4 t = ld_stream_f4(px);
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
        } else {
            xv[0] = __ldg(px);
        }
        float sr = 0.f, zr = 0.f;
        if (a.bcast == SPQ_PER_ROW) { sr = __ldg(a.scale + r); zr = __ldg(a.zp + r); }
        const float rm = (a.row_mul ? __ldg(a.row_mul + r) : 1.f) * a.mul;
        float dq[VEC], code[VEC], opv[VEC], sg[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float sj = (a.bcast == SPQ_PER_ROW) ? sr : s[j];
            const float zj = (a.bcast == SPQ_PER_ROW) ? zr : z[j];
            float centered;
            if constexpr (QTYPE == SPQ_MINMAX) {
                const MinMaxOut o = minmax_elem(xv[j], sj, zj, a.qp);
                dq[j] = o.dq; code[j] = o.code; centered = o.centered; sg[j] = 0.f;
            } else {
                const LogOut o = log_elem(xv[j], make_logch(zj, sj), a.qp);
                dq[j] = o.dq; code[j] = o.level; centered = o.level; sg[j] = o.sign;
            }
            const float base = (a.operand_kind == SPQ_OPERAND_CODE) ? centered : (a.operand_kind == SPQ_OPERAND_DEQUANT ? dq[j] : xv[j]);
            opv[j] = base * rm * cm[j];
        }
        const long long off = r * a.cols + c0;
        if constexpr (VEC == 4) {
            if (a.dequant) *reinterpret_cast<float4*>(a.dequant + off) = make_float4(dq[0], dq[1], dq[2], dq[3]);
            if (a.codes) *reinterpret_cast<int4*>(a.codes + off) = make_int4((int)code[0], (int)code[1], (int)code[2], (int)code[3]);
            if (a.sign) *reinterpret_cast<char4*>(a.sign + off) = make_char4((signed char)sg[0], (signed char)sg[1], (signed char)sg[2], (signed char)sg[3]);
            if (a.operand) {
                if (!a.transposed) {
                    *reinterpret_cast<uint2*>(a.operand + r * a.op_ld + c0) = make_uint2(pack_h2(opv[0], opv[1]), pack_h2(opv[2], opv[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) a.operand[(c0 + j) * a.op_ld + r] = f2h_sat(opv[j]);
                }
            }
        } else {
            if (a.dequant) a.dequant[off] = dq[0];
            if (a.codes) a.codes[off] = (int)code[0];
            if (a.sign) a.sign[off] = (signed char)sg[0];
            if (a.operand) a.operand[a.transposed ? (c0 * a.op_ld + r) : (r * a.op_ld + c0)] = f2h_sat(opv[0]);
        }
    }
}

// ---------------------------------------------------------------- fused activation-side kernel
// One CTA of G threads owns a row at a time (grid-stride over rows): the row stays in registers,
// its absmax gives the power-of-two scale of the raw fp16 operand, and the same registers are
// quantised into the code / dequant operand.  Per-column parameters live in registers across rows.
struct ActArgs {
    const float* x;
    long long M, K;
    const float* scale;
    const float* zp;
    int bcast;            // PER_COL or PER_TENSOR
    QParams qp;
    int operand_kind;
    const float* col_mul;
    float mul;
    unsigned short* a_q;
    unsigned short* a_raw;
    float* raw_row_scale;
};

template <int QTYPE, int NV>   // QTYPE: -1 none, 0 minmax, 1 log; NV float4 chunks per thread
__global__ void __launch_bounds__(256)
quantize_act_kernel(ActArgs a) {
    const int G = blockDim.x;
    const int tid = threadIdx.x;
    __shared__ float s_red[8];
    // per-column parameters of this thread's columns
    // min-max: s = scale, z = zero point.  log: s = log_range, z = log_min, ir = 1 / max(log_range, eps).
    float s[NV][4], z[NV][4], cm[NV][4], ir[NV][4];
    if constexpr (QTYPE >= 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = c + j < a.K;
                s[i][j] = ok ? bparam(a.scale, a.bcast, 0, c + j) : 1.f;
                z[i][j] = ok ? bparam(a.zp, a.bcast, 0, c + j) : 0.f;
                cm[i][j] = ((ok && a.col_mul) ? __ldg(a.col_mul + c + j) : 1.f) * a.mul;
                ir[i][j] = (QTYPE == SPQ_LOG) ? make_logch(z[i][j], s[i][j]).inv_range : 0.f;
            }
        }
    }
    for (long long row = blockIdx.x; row < a.M; row += gridDim.x) {
        const float* px = a.x + row * a.K;
        float4 v[NV];
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            v[i] = (c < a.K) ? ld_stream_f4(px + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
        }
        if (a.a_raw) {
            amax = warp_fmax(amax);
            if (G > 32) {
                __syncthreads();                      // s_red reuse across rows
                if ((tid & 31) == 0) s_red[tid >> 5] = amax;
                __syncthreads();
                amax = s_red[0];
                for (int w = 1; w < (G >> 5); ++w) amax = fmaxf(amax, s_red[w]);
            }
            // amax in [2^(E-1), 2^E)  ->  raw = x * 2^(8-E) in (-256, 256); inf/0 rows: scale 1
            int E = 0;
            if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = 8;
            E = E < -100 ? -100 : E;
            const float down = exp2f(static_cast<float>(8 - E));
            if (tid == 0 && a.raw_row_scale) a.raw_row_scale[row] = exp2f(static_cast<float>(E - 8));
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const long long c = (static_cast<long long>(i) * G + tid) * 4;
                if (c < a.K)
                    *reinterpret_cast<uint2*>(a.a_raw + row * a.K + c) =
                        make_uint2(pack_h2(v[i].x * down, v[i].y * down), pack_h2(v[i].z * down, v[i].w * down));
            }
        }
        if constexpr (QTYPE >= 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const long long c = (static_cast<long long>(i) * G + tid) * 4;
                if (c < a.K) {
                    const float xv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float base;
                        if constexpr (QTYPE == SPQ_MINMAX) {
                            const MinMaxOut r = minmax_elem(xv[j], s[i][j], z[i][j], a.qp);
                            base = (a.operand_kind == SPQ_OPERAND_CODE) ? r.centered : r.dq;
                        } else {
                            LogCh ch;
                            ch.log_min = z[i][j]; ch.log_range = s[i][j]; ch.inv_range = ir[i][j];
                            ch.range_c = (s[i][j] < LOG_EPS) ? LOG_EPS : s[i][j];
                            const LogOut r = log_elem(xv[j], ch, a.qp);
                            base = (a.operand_kind == SPQ_OPERAND_CODE) ? r.level : r.dq;
                        }
                        o[j] = base * cm[i][j];
                    }
                    *reinterpret_cast<uint2*>(a.a_q + row * a.K + c) = make_uint2(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]));
                }
            }
        }
    }
}

// Rows wider than the register-resident limit, or not 16-byte aligned (LM-head gradients, N = 50257):
// one CTA per row, two passes (the second pass hits L2), output rows `ld_out` apart.
__global__ void __launch_bounds__(256)
rowscale_wide_kernel(const float* __restrict__ x, long long M, long long K, unsigned short* __restrict__ out,
                     long long ld_out, float* __restrict__ row_scale) {
    __shared__ float s_red[8];
    const int tid = threadIdx.x;
    for (long long row = blockIdx.x; row < M; row += gridDim.x) {
        const float* p = x + row * K;
        float amax = 0.f;
        for (long long c = tid; c < K; c += 256) amax = fmaxf(amax, fabsf(__ldg(p + c)));
        amax = warp_fmax(amax);
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = amax;
        __syncthreads();
        amax = s_red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) amax = fmaxf(amax, s_red[w]);
        int E = 0;
        if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = 8;
        E = E < -100 ? -100 : E;
        const float down = exp2f(static_cast<float>(8 - E));
        if (tid == 0 && row_scale) row_scale[row] = exp2f(static_cast<float>(E - 8));
        unsigned short* o = out + row * ld_out;
        for (long long c = 2 * tid; c < K; c += 512) {
            const float a = __ldg(p + c) * down;
            if (c + 1 < K) *reinterpret_cast<unsigned int*>(o + c) = pack_h2(a, __ldg(p + c + 1) * down);
            else o[c] = f2h_sat(a);
        }
    }
}

__global__ void __launch_bounds__(256) ste_backward_kernel(const float* __restrict__ g, long long n, int clampit, float* __restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
    for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n && aligned16_dev(g + i) && aligned16_dev(out + i)) {
            float4 v = *reinterpret_cast<const float4*>(g + i);
            if (clampit) {
                v.x = fminf(fmaxf(v.x, -10.f), 10.f); v.y = fminf(fmaxf(v.y, -10.f), 10.f);
                v.z = fminf(fmaxf(v.z, -10.f), 10.f); v.w = fminf(fmaxf(v.w, -10.f), 10.f);
            }
            *reinterpret_cast<float4*>(out + i) = v;
        } else {
            for (long long j = i; j < n && j < i + 4; ++j) {
                float v = g[j];
                if (clampit) v = fminf(fmaxf(v, -10.f), 10.f);
                out[j] = v;
            }
        }
    }
}

static QParams make_qparams(int bits, int symmetric) {
    QParams q;
    q.bits = bits;
    q.symmetric = symmetric;
    q.n_sym = static_cast<float>(static_cast<double>(1ull << (bits - 1)) - 1.0);
    q.full = static_cast<float>(static_cast<double>(1ull << bits) - 1.0);
    return q;
}

template <int QTYPE>
static int launch_act(const ActArgs& a, cudaStream_t st) {
    const long long nvec = (a.K + 3) / 4;
    int G, NV;
    if (nvec <= 32) { G = 32; NV = 1; }
    else if (nvec <= 64) { G = 32; NV = 2; }
    else if (nvec <= 128) { G = 32; NV = 4; }
    else if (nvec <= 256) { G = 64; NV = 4; }
    else if (nvec <= 512) { G = 128; NV = 4; }
    else if (nvec <= 1024) { G = 256; NV = 4; }
    else if (nvec <= 2048) { G = 256; NV = 8; }
    else {
        set_error("spq_quantize_act: K = %lld > 8192 is not supported by the row-resident kernel", a.K);
        return SPQ_ERR_UNSUPPORTED;
    }
    long long ctas = static_cast<long long>(sm_count()) * (2048 / G > 32 ? 32 : 2048 / G);
    if (NV == 8) ctas = static_cast<long long>(sm_count()) * 4;
    if (ctas > a.M) ctas = a.M;
    const unsigned grid = static_cast<unsigned>(ctas);
    switch (NV) {
        case 1: quantize_act_kernel<QTYPE, 1><<<grid, G, 0, st>>>(a); break;
        case 2: quantize_act_kernel<QTYPE, 2><<<grid, G, 0, st>>>(a); break;
        case 4: quantize_act_kernel<QTYPE, 4><<<grid, G, 0, st>>>(a); break;
        default: quantize_act_kernel<QTYPE, 8><<<grid, G, 0, st>>>(a); break;
    }
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

}  // namespace quant
}  // namespace spq

using namespace spq;
using namespace spq::quant;

extern "C" int spq_fake_quantize(const float* x, int64_t rows, int64_t cols, const float* scale, const float* zero_point,
                                 int bcast, int qtype, int bits, int symmetric, float* dequant, int32_t* codes, int8_t* sign,
                                 spq_half_t* operand, int operand_kind, const float* row_mul, const float* col_mul, float mul,
                                 int operand_transposed, int64_t operand_ld, spq_stream_t stream) {
    SPQ_REQUIRE(x && scale && zero_point, "spq_fake_quantize: null pointer");
    SPQ_REQUIRE(rows > 0 && cols > 0, "spq_fake_quantize: empty tensor");
    SPQ_REQUIRE(bits >= 1 && bits < 32, "spq_fake_quantize: bits %d outside [1, 31]", bits);
    SPQ_REQUIRE(qtype == SPQ_MINMAX || qtype == SPQ_LOG, "spq_fake_quantize: unknown quantizer type %d", qtype);
    SPQ_REQUIRE(bcast >= 0 && bcast <= 2, "spq_fake_quantize: bad bcast %d", bcast);
    FqArgs a;
    a.x = x; a.rows = rows; a.cols = cols; a.scale = scale; a.zp = zero_point; a.bcast = bcast;
    a.qp = make_qparams(bits, symmetric);
    a.dequant = dequant; a.codes = codes; a.sign = sign; a.operand = operand; a.operand_kind = operand_kind;
    a.row_mul = row_mul; a.col_mul = col_mul; a.mul = mul; a.transposed = operand_transposed;
    a.op_ld = operand_ld > 0 ? operand_ld : (operand_transposed ? rows : cols);
    SPQ_REQUIRE(!operand || a.op_ld >= (operand_transposed ? rows : cols), "spq_fake_quantize: operand_ld too small");
    const bool vec = (cols % 4 == 0) && aligned16(x) && (!dequant || aligned16(dequant)) && (!codes || aligned16(codes)) &&
                     (!sign || (reinterpret_cast<uintptr_t>(sign) & 3u) == 0) &&
                     (!operand || ((reinterpret_cast<uintptr_t>(operand) & 7u) == 0 && (operand_transposed || (a.op_ld & 3) == 0)));
    const long long col_threads = vec ? cols / 4 : cols;
    const unsigned gx = static_cast<unsigned>((col_threads + 31) / 32);
    long long gy = (static_cast<long long>(sm_count()) * 8 + gx - 1) / gx;
    const long long max_gy = (rows + 7) / 8;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    dim3 grid(gx, static_cast<unsigned>(gy)), block(32, 8);
    cudaStream_t st = as_stream(stream);
    if (qtype == SPQ_MINMAX) {
        if (vec) fake_quantize_kernel<SPQ_MINMAX, 4><<<grid, block, 0, st>>>(a);
        else fake_quantize_kernel<SPQ_MINMAX, 1><<<grid, block, 0, st>>>(a);
    } else {
        if (vec) fake_quantize_kernel<SPQ_LOG, 4><<<grid, block, 0, st>>>(a);
        else fake_quantize_kernel<SPQ_LOG, 1><<<grid, block, 0, st>>>(a);
    }
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_quantize_act(const float* x, int64_t M, int64_t K, const float* scale, const float* zero_point, int bcast,
                                int qtype, int bits, int symmetric, int operand_kind, const float* col_mul, float mul,
                                spq_half_t* a_q, spq_half_t* a_raw, float* raw_row_scale, spq_stream_t stream) {
    SPQ_REQUIRE(x && M > 0 && K > 0, "spq_quantize_act: bad input");
    SPQ_REQUIRE((K % 4) == 0 && aligned16(x), "spq_quantize_act: K must be a multiple of 4 and x 16-byte aligned");
    SPQ_REQUIRE(qtype < 0 || (scale && zero_point && a_q), "spq_quantize_act: quantised output requested without parameters");
    SPQ_REQUIRE(qtype < 0 || bcast == SPQ_PER_COL || bcast == SPQ_PER_TENSOR, "spq_quantize_act: per-row scales are not an activation layout");
    SPQ_REQUIRE(qtype >= 0 || a_raw, "spq_quantize_act: nothing to do");
    SPQ_REQUIRE(qtype < 0 || (bits >= 1 && bits < 32), "spq_quantize_act: bits %d", bits);
    ActArgs a;
    a.x = x; a.M = M; a.K = K; a.scale = scale; a.zp = zero_point; a.bcast = bcast;
    a.qp = make_qparams(qtype < 0 ? 8 : bits, symmetric);
    a.operand_kind = operand_kind; a.col_mul = col_mul; a.mul = mul;
    a.a_q = a_q; a.a_raw = a_raw; a.raw_row_scale = raw_row_scale;
    cudaStream_t st = as_stream(stream);
    if (qtype < 0) return launch_act<-1>(a, st);
    if (qtype == SPQ_MINMAX) return launch_act<SPQ_MINMAX>(a, st);
    if (qtype == SPQ_LOG) return launch_act<SPQ_LOG>(a, st);
    set_error("spq_quantize_act: unknown quantizer type %d", qtype);
    return SPQ_ERR_INVALID;
}

extern "C" int spq_rowscale_f16(const float* g, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out, float* row_scale,
                                spq_stream_t stream) {
    SPQ_REQUIRE(g && out && M > 0 && N > 0, "spq_rowscale_f16: bad arguments");
    if (ld_out <= 0) ld_out = N;
    SPQ_REQUIRE(ld_out >= N, "spq_rowscale_f16: ld_out < N");
    if (ld_out == N && (N % 4) == 0 && N <= 8192 && aligned16(g))
        return spq_quantize_act(g, M, N, nullptr, nullptr, SPQ_PER_TENSOR, -1, 8, 1, SPQ_OPERAND_RAW, nullptr, 1.0f, nullptr,
                                out, row_scale, stream);
    SPQ_REQUIRE((ld_out % 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0, "spq_rowscale_f16: ld_out must be even");
    long long ctas = static_cast<long long>(sm_count()) * 8;
    if (ctas > M) ctas = M;
    rowscale_wide_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(g, M, N, out, ld_out, row_scale);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_ste_backward(const float* grad, int64_t n, int qtype, float* out, spq_stream_t stream) {
    SPQ_REQUIRE(grad && out && n > 0, "spq_ste_backward: bad arguments");
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    ste_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(grad, n, qtype == SPQ_LOG ? 1 : 0, out);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}
