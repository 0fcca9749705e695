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
#include <stdlib.h>
#include <cuda_fp8.h>

#include "spq_common.cuh"

namespace spq {
namespace quant {

constexpr float LOG_EPS = 1e-5f;   // p1/quantization_methods.py:35 (hard-coded)

// four floats -> four e4m3 bytes (saturating).  Integer codes |c| <= 16 are exact in e4m3 (3 mantissa bits).
__device__ __forceinline__ unsigned int pack_e4m3x4(float a, float b, float c, float d) {
    const unsigned int lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const unsigned int hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
__device__ __forceinline__ unsigned char to_e4m3(float a) {
    return static_cast<unsigned char>(__nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3));
}

struct QParams {
    int bits;
    int symmetric;
    float n_sym;    // 2^(b-1) - 1
    float full;     // 2^b - 1
    int debug;      // profiling experiments only: bit0 = never take the exact slow path
};

// ---------------------------------------------------------------- element arithmetic
// Every quantiser below has a FAST path (reciprocal multiplies, lg2.approx, FMA contraction) whose
// pre-rounding value provably differs from the reference arithmetic by less than a per-channel
// `band`; whenever the fast value lies within that band of a rounding tie, the element is redone
// with the EXACT sequence (IEEE division, correctly rounded log2, separate roundings).  The integer
// code / level index is therefore bit-identical to the reference everywhere, at ~20 instructions
// per element instead of ~100.  Dequantised VALUES are tolerance-level quantities (rel 1e-3 in
// BASELINE.json; ~1e-6 here) and always use the fast formulas.
struct MinMaxOut { float dq; float code; float centered; };
__device__ __noinline__ float minmax_code_exact(float x, float s, float zp, int symmetric);
__device__ __noinline__ float log_level_exact(float axc, float log_min, float range_c, int symmetric, float n_sym, float full);

struct MmCol {            // per-channel constants, min-max
    float s, inv_s, zp;
};
__device__ __forceinline__ MmCol make_mmcol(float s, float zp) {
    MmCol c;
    c.s = s; c.zp = zp;
    c.inv_s = __frcp_rn(s);
    return c;
}

__device__ __forceinline__ MinMaxOut minmax_elem(float x, const MmCol& c, const QParams& qp) {
    MinMaxOut o;
    // reference: q = round(x / s [+ zp]) (p1/quantization_methods.py:14 / :18)
    float t = qp.symmetric ? x * c.inv_s : fmaf(x, c.inv_s, c.zp);
    float q = rintf(t);
    const float tie_dist = fabsf(fabsf(t - q) - 0.5f);
    // |t - exact| <= |t| * 2^-21 (reciprocal, product and, if asymmetric, sum roundings); beyond the clamp
    // range the rounding cannot matter
    const float tmag = fminf(fabsf(t), 4.0f * qp.full) + (qp.symmetric ? 0.f : qp.full);   // |x/s| <= |t| + zp
    if (tie_dist <= tmag * 4.76837158203125e-07f && !(qp.debug & 1)) {
        q = minmax_code_exact(x, c.s, c.zp, qp.symmetric);
    }
    if (qp.symmetric) {                                   // :15-16
        q = fminf(fmaxf(q, -qp.n_sym), qp.n_sym);
        o.centered = q;
    } else {                                              // :19-20
        q = fminf(fmaxf(q, 0.f), qp.full);
        o.centered = __fsub_rn(q, c.zp);
    }
    o.code = q;
    o.dq = __fmul_rn(o.centered, c.s);
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

// The exact sequences, OUT OF LINE: they are taken by < 0.1 % of the elements, and inlined next to every
// unrolled fast path they push the hot loop past the instruction caches (64 KB of code, stall_no_inst).
__device__ __noinline__ float log_level_exact(float axc, float log_min, float range_c, int symmetric, float n_sym, float full) {
    QParams qp;
    qp.bits = 0; qp.symmetric = symmetric; qp.n_sym = n_sym; qp.full = full; qp.debug = 0;
    float r = rintf(log_prelevel(log2_cr(axc), log_min, range_c, qp));
    return symmetric ? fminf(fmaxf(r, -n_sym), n_sym) : fminf(fmaxf(r, 0.f), full);
}
__device__ __noinline__ float minmax_code_exact(float x, float s, float zp, int symmetric) {
    return rintf(symmetric ? __fdiv_rn(x, s) : __fadd_rn(__fdiv_rn(x, s), zp));
}

struct LogCol {           // per-channel constants, log
    float log_min, log_range, range_c;
    float inv, c0;        // ln = sat(l * inv + c0),  inv = 1 / max(range, eps), c0 = -log_min * inv
    float band;           // |v_fast - v_exact| bound (see make_logcol)
    float out_add;        // added to the dequantised exponent: log2 of a power-of-two output multiplier
    float e0;             // log_min + out_add
};
__device__ __forceinline__ LogCol make_logcol(float log_min, float log_range, const QParams& qp, float out_add = 0.f) {
    LogCol c;
    c.log_min = log_min; c.log_range = log_range; c.out_add = out_add; c.e0 = log_min + out_add;
    c.range_c = (log_range < LOG_EPS) ? LOG_EPS : log_range;          // :43 clamp(min=eps)
    c.inv = __frcp_rn(c.range_c);
    c.c0 = -log_min * c.inv;
    const float nl = qp.symmetric ? qp.n_sym : qp.full;
    const float lev_mul = qp.symmetric ? 2.f * nl : nl;
    // l ranges over [log_min, log_min + range] where it matters (outside, ln saturates and v = +-n exactly).
    // Error budget of the fast pre-rounding value:  lg2.approx 2^-20, roundings of l*inv, c0 and the FMA
    // 3 * Lmax * inv * 2^-23 + 2^-23, then v = ln * lev_mul - n: n * 2^-22.  Doubled for margin.
    const float lmax = fmaxf(fabsf(log_min), fabsf(log_min + log_range));
    c.band = 2.0f * (lev_mul * (c.inv * (9.5367431640625e-07f + 3.0f * lmax * 1.1920928955078125e-07f) + 1.1920928955078125e-07f) +
                     nl * 2.384185791015625e-07f) + 1e-9f;
    return c;
}

// lg2.approx with flush-to-zero: the argument is max(|x|, 1e-5), never subnormal, so the result is the one of
// __log2f() without the three instructions it spends on rescaling subnormal inputs
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ LogOut log_elem(float x, const LogCol& ch, const QParams& qp) {
    LogOut o;
    const float ax = fabsf(x);
    const bool zero = ax < LOG_EPS;                                      // :36
    const float axc = fmaxf(ax, LOG_EPS);                                // :40 clamp(min=eps): >= 1e-5, never denormal
    const float nl = qp.symmetric ? qp.n_sym : qp.full;
    const float lev_mul = qp.symmetric ? 2.f * nl : nl;

    float v = __saturatef(fmaf(lg2_approx(axc), ch.inv, ch.c0));            // ln in [0, 1]
    v = qp.symmetric ? fmaf(v, lev_mul, -nl) : v * lev_mul;
    float r = rintf(v);
    if (fabsf(fabsf(v - r) - 0.5f) <= ch.band && !(qp.debug & 1)) {      // near a tie: the reference's exact sequence
        r = log_level_exact(axc, ch.log_min, ch.range_c, qp.symmetric, qp.n_sym, qp.full);
    }
    o.level = r;
    // value (:50-74): qn = L/(2n) + 0.5 (symmetric) or L/n; 2^(qn * range + log_min) * sign, 0 under the zero mask
    const float qn = qp.symmetric ? fmaf(r, __frcp_rn(lev_mul), 0.5f) : r * __frcp_rn(nl);
    const float mag = ex2_approx(fmaf(qn, ch.log_range, ch.e0));
    o.sign = zero ? 0.f : copysignf(1.0f, x);
    o.dq = zero ? 0.f : copysignf(mag, x);
    return o;
}

// ---- branch-free split of the two quantisers, for kernels that process 4 elements at a time:
//   *_fast  : level / code from the fast arithmetic + "needs the exact path" flag, no branch
//   *_value : operand value from the final level / code
// so that the four independent dependency chains of a float4 interleave and at most ONE (rarely
// taken) fix-up branch is executed per float4 instead of one per element.
__device__ __forceinline__ float log_level_fast(float x, const LogCol& ch, const QParams& qp, bool& tie) {
    const float nl = qp.symmetric ? qp.n_sym : qp.full;
    const float lev_mul = qp.symmetric ? 2.f * nl : nl;
    float v = __saturatef(fmaf(lg2_approx(fmaxf(fabsf(x), LOG_EPS)), ch.inv, ch.c0));
    v = qp.symmetric ? fmaf(v, lev_mul, -nl) : v * lev_mul;
    const float r = rintf(v);
    tie = fabsf(fabsf(v - r) - 0.5f) <= ch.band;
    return r;
}
__device__ __forceinline__ float log_value(float x, float r, const LogCol& ch, const QParams& qp, float inv_lev) {
    const float qn = qp.symmetric ? fmaf(r, inv_lev, 0.5f) : r * inv_lev;
    const float mag = ex2_approx(fmaf(qn, ch.log_range, ch.e0));
    return (fabsf(x) < LOG_EPS) ? 0.f : copysignf(mag, x);
}
__device__ __forceinline__ float minmax_code_fast(float x, const MmCol& c, const QParams& qp, bool& tie) {
    const float t = qp.symmetric ? x * c.inv_s : fmaf(x, c.inv_s, c.zp);
    const float q = rintf(t);
    const float tmag = fminf(fabsf(t), 4.0f * qp.full) + (qp.symmetric ? 0.f : qp.full);
    tie = fabsf(fabsf(t - q) - 0.5f) <= tmag * 4.76837158203125e-07f;
    return q;
}
__device__ __forceinline__ float minmax_centered(float q, const MmCol& c, const QParams& qp) {
    if (qp.symmetric) return fminf(fmaxf(q, -qp.n_sym), qp.n_sym);
    return __fsub_rn(fminf(fmaxf(q, 0.f), qp.full), c.zp);
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
                const MinMaxOut o = minmax_elem(xv[j], make_mmcol(sj, zj), a.qp);
                dq[j] = o.dq; code[j] = o.code; centered = o.centered; sg[j] = 0.f;
            } else {
                const LogOut o = log_elem(xv[j], make_logcol(zj, sj, a.qp), a.qp);
                dq[j] = o.dq; code[j] = o.level; centered = o.level; sg[j] = o.sign;
            }
            const float base = (a.operand_kind == SPQ_OPERAND_CODE || a.operand_kind == SPQ_OPERAND_CODE_E4M3)
                                   ? centered : (a.operand_kind == SPQ_OPERAND_DEQUANT ? dq[j] : xv[j]);
            opv[j] = base * rm * cm[j];
        }
        const long long off = r * a.cols + c0;
        if constexpr (VEC == 4) {
            if (a.dequant) *reinterpret_cast<float4*>(a.dequant + off) = make_float4(dq[0], dq[1], dq[2], dq[3]);
            if (a.codes) *reinterpret_cast<int4*>(a.codes + off) = make_int4((int)code[0], (int)code[1], (int)code[2], (int)code[3]);
            if (a.sign) *reinterpret_cast<char4*>(a.sign + off) = make_char4((signed char)sg[0], (signed char)sg[1], (signed char)sg[2], (signed char)sg[3]);
            if (a.operand && a.operand_kind == SPQ_OPERAND_CODE_E4M3) {
                unsigned char* o8 = reinterpret_cast<unsigned char*>(a.operand);      // op_ld in bytes
                if (!a.transposed) {
                    *reinterpret_cast<unsigned int*>(o8 + r * a.op_ld + c0) = pack_e4m3x4(opv[0], opv[1], opv[2], opv[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) o8[(c0 + j) * a.op_ld + r] = to_e4m3(opv[j]);
                }
            } else if (a.operand) {
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
            if (a.operand && a.operand_kind == SPQ_OPERAND_CODE_E4M3)
                reinterpret_cast<unsigned char*>(a.operand)[a.transposed ? (c0 * a.op_ld + r) : (r * a.op_ld + c0)] = to_e4m3(opv[0]);
            else if (a.operand) a.operand[a.transposed ? (c0 * a.op_ld + r) : (r * a.op_ld + c0)] = f2h_sat(opv[0]);
        }
    }
}

// ---------------------------------------------------------------- fused activation-side kernel
// One CTA of G threads owns a row at a time (grid-stride over rows): the row stays in registers,
// its absmax gives the power-of-two scale of the raw fp16 operand, and the same registers are
// quantised into the code / dequant operand.  Per-column parameters live in registers across rows.
struct ActArgs {
    const void* x;        // float32, or float16 when x_half
    int x_half;
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
    float* raw_row_scale;       // row-scaled variant only
    const float* raw_col_mul;   // elementwise variant: per-column multiplier of the raw operand
    float* max_scale;           // row-scaled variant: optional device scalar, max over rows of the row scale
    const float* dgelu_y;       // row-scaled DGELU variant: pre-activation y; the row that is scaled is x * gelu'(y)
};

// Row-scaled raw operand (no quantiser): one CTA of G threads owns a row at a time, the row stays in
// registers (NV float4 per thread), its absmax gives the power-of-two scale.  Used where no calibrated
// bound exists: calibration pass, 32-bit path, LM head, gradients.
// d/dy of the exact (erf) GELU, as torch's gelu_backward: Phi(y) + y phi(y)
__device__ __forceinline__ float dgelu_erf(float y) {
    const float cdf = 0.5f * (1.0f + erff(y * 0.70710678118654752f));
    const float pdf = __expf(-0.5f * y * y) * 0.39894228040143268f;
    return fmaf(y, pdf, cdf);
}

// DGELU: the row is x[m,:] * gelu'(y[m,:]) -- the gradient entering a linear whose output went through GELU, straight from
// the gradient of the GELU output: torch's gelu_backward pass (12 B / element) and its float32 result never exist
template <int NV, typename XT, bool DGELU = false>
__global__ void __launch_bounds__(256)
rowscale_kernel(ActArgs a) {
    const int G = blockDim.x;
    const int tid = threadIdx.x;
    __shared__ float s_red[8];
    float cta_max_scale = 0.f;
    for (long long row = blockIdx.x; row < a.M; row += gridDim.x) {
        const XT* px = static_cast<const XT*>(a.x) + row * a.K;
        float4 v[NV];
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            v[i] = (c < a.K) ? ld_stream_x4<XT>(px + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (DGELU) {
                if (c < a.K) {
                    const float4 y = ld_stream_f4(a.dgelu_y + row * a.K + c);
                    v[i].x *= dgelu_erf(y.x); v[i].y *= dgelu_erf(y.y); v[i].z *= dgelu_erf(y.z); v[i].w *= dgelu_erf(y.w);
                }
            }
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
        }
        amax = warp_fmax(amax);
        if (G > 32) {
            __syncthreads();                      // s_red reuse across rows
            if ((tid & 31) == 0) s_red[tid >> 5] = amax;
            __syncthreads();
            amax = s_red[0];
            for (int w = 1; w < (G >> 5); ++w) amax = fmaxf(amax, s_red[w]);
        }
        // amax in [2^(E-1), 2^E)  ->  raw = x * 2^(8-E) in (-256, 256); inf / NaN rows: scale 1.  An all-zero row
        // (the last position of every sequence in a next-token loss has no gradient) gets the SMALLEST scale, 2^-108:
        // callers fold `scale / max over rows` into fp16 operands of token reductions, and a zero row reporting
        // scale 1 against real rows at ~2^-20 pushed those operands into the fp16 subnormal range
        int E = 0;
        if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = (amax == 0.f) ? -100 : 8;
        E = E < -100 ? -100 : E;
        const float down = exp2f(static_cast<float>(8 - E));
        const float up = exp2f(static_cast<float>(E - 8));
        if (tid == 0 && a.raw_row_scale) a.raw_row_scale[row] = up;
        cta_max_scale = fmaxf(cta_max_scale, up);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            if (c < a.K)
                *reinterpret_cast<uint2*>(a.a_raw + row * a.K + c) =
                    make_uint2(pack_h2(v[i].x * down, v[i].y * down), pack_h2(v[i].z * down, v[i].w * down));
        }
    }
    // scales are positive floats: their bit patterns order like integers
    if (tid == 0 && a.max_scale) atomicMax(reinterpret_cast<int*>(a.max_scale), __float_as_int(cta_max_scale));
}

// Fused activation-side kernel (calibrated quantiser): purely elementwise.  Thread = 4 consecutive
// columns (their constants live in registers), block (32, 8) strides over rows with four 16-byte loads
// in flight per thread; writes the quantised operand and, for the LoRA branch, the raw operand scaled
// per COLUMN by a power of two derived from the calibrated bound (saturating conversion).
// SYM / KIND >= 0: compile-time copies of qp.symmetric / operand_kind for the configurations the model uses
// (symmetric quantisers; codes for min-max, dequantised values for log) -- the per-float4 uniform branches and
// the duplicated code paths of the generic kernel cost ~10 % of its issue slots.
template <int QTYPE, typename XT, int SYM = -1, int KIND = -1>
__global__ void __launch_bounds__(256, 3)
quantize_act_kernel(ActArgs a) {
    const long long c0 = (static_cast<long long>(blockIdx.x) * 32 + threadIdx.x) * 4;
    if (c0 >= a.K) return;
    QParams qp = a.qp;                                 // (a copy: kernel parameters are read-only)
    if constexpr (SYM >= 0) qp.symmetric = SYM;
    const int kind = (KIND >= 0) ? KIND : a.operand_kind;
    MmCol mm[4];
    LogCol lg[4];
    float cm[4], rm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float sc = bparam(a.scale, a.bcast, 0, c0 + j);
        const float zp = bparam(a.zp, a.bcast, 0, c0 + j);
        cm[j] = (a.col_mul ? __ldg(a.col_mul + c0 + j) : 1.f) * a.mul;
        rm[j] = a.raw_col_mul ? __ldg(a.raw_col_mul + c0 + j) : 1.f;
        if constexpr (QTYPE == SPQ_MINMAX) mm[j] = make_mmcol(sc, zp);
        else lg[j] = make_logcol(zp, sc, qp);
    }
    const long long rstep = static_cast<long long>(gridDim.y) * 8;
    const float nl_ = qp.symmetric ? qp.n_sym : qp.full;
    const float inv_lev = __frcp_rn(qp.symmetric ? 2.f * nl_ : nl_);
    const bool no_exact = (qp.debug & 1) != 0;
    auto one = [&](long long r, const float4& v) {
        const float xv[4] = {v.x, v.y, v.z, v.w};
        float q[4];
        bool tie[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if constexpr (QTYPE == SPQ_MINMAX) q[j] = minmax_code_fast(xv[j], mm[j], qp, tie[j]);
            else q[j] = log_level_fast(xv[j], lg[j], qp, tie[j]);
        }
        if ((tie[0] | tie[1] | tie[2] | tie[3]) && !no_exact) {           // rare: redo the flagged elements exactly
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (tie[j]) {
                    if constexpr (QTYPE == SPQ_MINMAX) q[j] = minmax_code_exact(xv[j], mm[j].s, mm[j].zp, qp.symmetric);
                    else q[j] = log_level_exact(fmaxf(fabsf(xv[j]), LOG_EPS), lg[j].log_min, lg[j].range_c, qp.symmetric,
                                                qp.n_sym, qp.full);
                }
            }
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float base;
            if constexpr (QTYPE == SPQ_MINMAX) {
                const float cen = minmax_centered(q[j], mm[j], qp);
                base = (kind == SPQ_OPERAND_CODE || kind == SPQ_OPERAND_CODE_E4M3) ? cen : __fmul_rn(cen, mm[j].s);
            } else {
                // q is already in range: the fast path rounds v = sat(.) * 2n - n (or sat(.) * n), the exact path clamps
                const float lvl = q[j];
                base = (kind == SPQ_OPERAND_CODE) ? lvl : log_value(xv[j], lvl, lg[j], qp, inv_lev);
            }
            o[j] = base * cm[j];
        }
        if (kind == SPQ_OPERAND_CODE_E4M3)           // one byte per code (a_q is an e4m3 buffer, rows K bytes apart)
            *reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(a.a_q) + r * a.K + c0) = pack_e4m3x4(o[0], o[1], o[2], o[3]);
        else
            *reinterpret_cast<uint2*>(a.a_q + r * a.K + c0) = make_uint2(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]));
        if (a.a_raw)
            *reinterpret_cast<uint2*>(a.a_raw + r * a.K + c0) =
                make_uint2(pack_h2(xv[0] * rm[0], xv[1] * rm[1]), pack_h2(xv[2] * rm[2], xv[3] * rm[3]));
    };
    long long r = static_cast<long long>(blockIdx.y) * 8 + threadIdx.y;
    for (; r + 3 * rstep < a.M; r += 4 * rstep) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream_x4<XT>(static_cast<const XT*>(a.x) + (r + u * rstep) * a.K + c0);
#pragma unroll
        for (int u = 0; u < 4; ++u) one(r + u * rstep, v[u]);
    }
    for (; r < a.M; r += rstep) one(r, ld_stream_x4<XT>(static_cast<const XT*>(a.x) + r * a.K + c0));
}

// Rows wider than the register-resident limit, or not 16-byte aligned (LM-head gradients, N = 50257):
// one CTA per row, two passes (the second pass hits L2), output rows `ld_out` apart.
__global__ void __launch_bounds__(256)
rowscale_wide_kernel(const float* __restrict__ x, long long M, long long K, unsigned short* __restrict__ out,
                     long long ld_out, float* __restrict__ row_scale, float* __restrict__ max_scale) {
    __shared__ float s_red[8];
    const int tid = threadIdx.x;
    float cta_max_scale = 0.f;
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
        if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = (amax == 0.f) ? -100 : 8;   // see rowscale_kernel
        E = E < -100 ? -100 : E;
        const float down = exp2f(static_cast<float>(8 - E));
        if (tid == 0 && row_scale) row_scale[row] = exp2f(static_cast<float>(E - 8));
        cta_max_scale = fmaxf(cta_max_scale, exp2f(static_cast<float>(E - 8)));
        unsigned short* o = out + row * ld_out;
        for (long long c = 2 * tid; c < K; c += 512) {
            const float a = __ldg(p + c) * down;
            if (c + 1 < K) *reinterpret_cast<unsigned int*>(o + c) = pack_h2(a, __ldg(p + c + 1) * down);
            else o[c] = f2h_sat(a);
        }
    }
    if (tid == 0 && max_scale) atomicMax(reinterpret_cast<int*>(max_scale), __float_as_int(cta_max_scale));
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
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("SPQ_QUANT_DEBUG"); dbg = e ? atoi(e) : 0; }
        q.debug = dbg;
    }
    return q;
}

static void rowscale_cfg(long long K, int& G, int& NV) {
    const long long nvec = (K + 3) / 4;
    if (nvec <= 32) { G = 32; NV = 1; }
    else if (nvec <= 64) { G = 32; NV = 2; }
    else if (nvec <= 128) { G = 32; NV = 4; }
    else if (nvec <= 256) { G = 64; NV = 4; }
    else if (nvec <= 512) { G = 128; NV = 4; }
    else if (nvec <= 1024) { G = 256; NV = 4; }
    else { G = 256; NV = 8; }
}
static int launch_rowscale(const ActArgs& a, cudaStream_t st) {
    int G, NV;
    rowscale_cfg(a.K, G, NV);
    long long ctas = static_cast<long long>(sm_count()) * (2048 / G > 32 ? 32 : 2048 / G);
    if (NV == 8) ctas = static_cast<long long>(sm_count()) * 4;
    if (ctas > a.M) ctas = a.M;
    const unsigned grid = static_cast<unsigned>(ctas);
    if (a.dgelu_y) {
        switch (NV) {
            case 1: rowscale_kernel<1, float, true><<<grid, G, 0, st>>>(a); break;
            case 2: rowscale_kernel<2, float, true><<<grid, G, 0, st>>>(a); break;
            case 4: rowscale_kernel<4, float, true><<<grid, G, 0, st>>>(a); break;
            default: rowscale_kernel<8, float, true><<<grid, G, 0, st>>>(a); break;
        }
    } else if (a.x_half) {
        switch (NV) {
            case 1: rowscale_kernel<1, __half><<<grid, G, 0, st>>>(a); break;
            case 2: rowscale_kernel<2, __half><<<grid, G, 0, st>>>(a); break;
            case 4: rowscale_kernel<4, __half><<<grid, G, 0, st>>>(a); break;
            default: rowscale_kernel<8, __half><<<grid, G, 0, st>>>(a); break;
        }
    } else {
        switch (NV) {
            case 1: rowscale_kernel<1, float><<<grid, G, 0, st>>>(a); break;
            case 2: rowscale_kernel<2, float><<<grid, G, 0, st>>>(a); break;
            case 4: rowscale_kernel<4, float><<<grid, G, 0, st>>>(a); break;
            default: rowscale_kernel<8, float><<<grid, G, 0, st>>>(a); break;
        }
    }
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

// ---------------------------------------------------------------- LayerNorm-fused activation side
// SwitchableLayerNorm feeds exactly one consumer on this path -- ln_1 -> c_attn, ln_2 -> c_fc, ln_f -> LM head
// (p1/models_sp.py:139-147, 316-319) -- so its float32 output only exists to be read back by the consumer's
// activation-side kernel.  The two kernels below normalise a row in registers (the arithmetic of
// spq_layernorm.cu: mean, biased variance, w * ((x - mean) * rstd) + b) and hand the normalised values straight to
//   (a) the calibrated quantiser: a_q / a_raw as spq_quantize_act would have written them   (quantised forward), or
//   (b) the row-scaled fp16 operand of spq_rowscale_f16 plus the per-column statistics of spq_minmax_stats
//       (calibration pass, 32-bit path),
// 4 B read + 2..4 B written per element instead of 4 + 4 (LayerNorm) + 4 + 2..4 (consumer) [+ 4 (statistics)].
// One CTA of G threads owns RPI rows at a time; a thread owns the same 4 * NV columns of every row, so the
// per-column constants stay in registers.  The codes are those of spq_quantize_act on the SAME normalised values
// (identical element arithmetic, including the exact tie path).
struct LnArgs {
    const float* x;
    long long M, K;
    const float* ln_w;
    const float* ln_b;
    float ln_eps;
    float* y;                   // optional float32 copy of the normalised rows (null: not written)
    // (a) quantiser
    const float* scale;
    const float* zp;
    int bcast;
    QParams qp;
    int operand_kind;
    const float* col_mul;
    float mul;
    unsigned short* a_q;
    unsigned short* a_raw;
    const float* raw_col_mul;
    // (b) row-scaled operand + statistics
    unsigned short* x16;
    float* row_scale;
    int stats_mode;             // 0 none, 1 min / max of y, 2 min / max of |y| + any(|y| > eps)   (log quantisers)
    float stat_eps;
    float* pmin;                // [gridDim.x, K] per-CTA partials
    float* pmax;
    int32_t* flags;
};

// Layout: ONE WARP PER ROW (8 rows per CTA in flight, no block barrier in the row loop): lane l owns the float4 column
// groups (32 i + l), i < NV, so every load / store instruction of a warp covers 512 contiguous bytes and a thread has NV
// independent 16-byte loads in flight.  (A first version with one CTA of K / 4 threads per row pair and block-wide
// reductions ran at 1.5-2.1 TB/s -- slower than the kernels it replaced; profiles/r02b_ln_fused_*.)  Per-column
// constants (LayerNorm weight / bias, quantiser constants, operand multipliers) sit in shared memory, filled once per
// persistent CTA: a lane owns 4 NV columns, too many to keep their ~9 constants each in registers.
constexpr int LN_WARPS = 8;

template <int NV>
__device__ __forceinline__ void ln_warp_normalise(const LnArgs& a, long long row, int lane, const float* s_w, const float* s_b,
                                                  float4 (&v)[NV]) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[i] = (c < a.K) ? ld_stream_f4(a.x + row * a.K + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float inv_c = 1.0f / static_cast<float>(a.K);
    const float mean = warp_sum(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < a.K) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * inv_c + a.ln_eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 w4 = *reinterpret_cast<const float4*>(s_w + c);
        const float4 b4 = *reinterpret_cast<const float4*>(s_b + c);
        float4 o;
        o.x = w4.x * ((v[i].x - mean) * rstd) + b4.x;
        o.y = w4.y * ((v[i].y - mean) * rstd) + b4.y;
        o.z = w4.z * ((v[i].z - mean) * rstd) + b4.z;
        o.w = w4.w * ((v[i].w - mean) * rstd) + b4.w;
        v[i] = o;
        if (a.y && c < a.K) *reinterpret_cast<float4*>(a.y + row * a.K + c) = o;
    }
}

// (a) LayerNorm -> calibrated quantiser.  Shared table: 9 arrays of Kp = 128 NV floats:
//   w, b, cm, rm, then min-max: s, inv_s, zp (2 unused) / log: inv, c0, band, log_range, log_min
template <int QTYPE, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_quantize_act_kernel(LnArgs a) {
    extern __shared__ __align__(16) float s_tab[];
    constexpr int Kp = NV * 128;
    float* s_w = s_tab;
    float* s_b = s_tab + Kp;
    float* s_cm = s_tab + 2 * Kp;
    float* s_rm = s_tab + 3 * Kp;
    float* s_q = s_tab + 4 * Kp;                 // 5 quantiser arrays
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    QParams qp = a.qp;
    const int kind = a.operand_kind;
    for (int c = tid; c < Kp; c += LN_WARPS * 32) {
        const bool in = c < a.K;
        const int cc = in ? c : 0;
        s_w[c] = in ? __ldg(a.ln_w + cc) : 0.f;
        s_b[c] = in ? __ldg(a.ln_b + cc) : 0.f;
        s_cm[c] = (a.col_mul ? __ldg(a.col_mul + cc) : 1.f) * a.mul;
        s_rm[c] = a.raw_col_mul ? __ldg(a.raw_col_mul + cc) : 1.f;
        const float sc = bparam(a.scale, a.bcast, 0, cc);
        const float zp = bparam(a.zp, a.bcast, 0, cc);
        if constexpr (QTYPE == SPQ_MINMAX) {
            const MmCol m = make_mmcol(sc, zp);
            s_q[c] = m.s; s_q[Kp + c] = m.inv_s; s_q[2 * Kp + c] = m.zp;
        } else {
            const LogCol l = make_logcol(zp, sc, qp);
            s_q[c] = l.inv; s_q[Kp + c] = l.c0; s_q[2 * Kp + c] = l.band; s_q[3 * Kp + c] = l.log_range; s_q[4 * Kp + c] = l.log_min;
        }
    }
    __syncthreads();
    const float nl_ = qp.symmetric ? qp.n_sym : qp.full;
    const float inv_lev = __frcp_rn(qp.symmetric ? 2.f * nl_ : nl_);
    const bool no_exact = (qp.debug & 1) != 0;
    for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.M; row += static_cast<long long>(gridDim.x) * LN_WARPS) {
        float4 v[NV];
        ln_warp_normalise<NV>(a, row, lane, s_w, s_b, v);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c0 = (i * 32 + lane) * 4;
            if (c0 >= a.K) continue;
            const float xv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            MmCol mm[4];
            LogCol lg[4];
            {
                const float4 q0 = *reinterpret_cast<const float4*>(s_q + c0);
                const float4 q1 = *reinterpret_cast<const float4*>(s_q + Kp + c0);
                const float4 q2 = *reinterpret_cast<const float4*>(s_q + 2 * Kp + c0);
                const float f0[4] = {q0.x, q0.y, q0.z, q0.w}, f1[4] = {q1.x, q1.y, q1.z, q1.w}, f2[4] = {q2.x, q2.y, q2.z, q2.w};
                if constexpr (QTYPE == SPQ_MINMAX) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { mm[j].s = f0[j]; mm[j].inv_s = f1[j]; mm[j].zp = f2[j]; }
                } else {
                    const float4 q3 = *reinterpret_cast<const float4*>(s_q + 3 * Kp + c0);
                    const float4 q4 = *reinterpret_cast<const float4*>(s_q + 4 * Kp + c0);
                    const float f3[4] = {q3.x, q3.y, q3.z, q3.w}, f4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        lg[j].inv = f0[j]; lg[j].c0 = f1[j]; lg[j].band = f2[j]; lg[j].log_range = f3[j]; lg[j].log_min = f4[j];
                        lg[j].range_c = (f3[j] < LOG_EPS) ? LOG_EPS : f3[j];
                        lg[j].out_add = 0.f; lg[j].e0 = f4[j];             // make_logcol: e0 = log_min + 0
                    }
                }
            }
            float q[4];
            bool tie[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (QTYPE == SPQ_MINMAX) q[j] = minmax_code_fast(xv[j], mm[j], qp, tie[j]);
                else q[j] = log_level_fast(xv[j], lg[j], qp, tie[j]);
            }
            if ((tie[0] | tie[1] | tie[2] | tie[3]) && !no_exact) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (tie[j]) {
                        if constexpr (QTYPE == SPQ_MINMAX) q[j] = minmax_code_exact(xv[j], mm[j].s, mm[j].zp, qp.symmetric);
                        else q[j] = log_level_exact(fmaxf(fabsf(xv[j]), LOG_EPS), lg[j].log_min, lg[j].range_c, qp.symmetric,
                                                    qp.n_sym, qp.full);
                    }
                }
            }
            const float4 cm4 = *reinterpret_cast<const float4*>(s_cm + c0);
            const float cmv[4] = {cm4.x, cm4.y, cm4.z, cm4.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float base;
                if constexpr (QTYPE == SPQ_MINMAX) {
                    const float cen = minmax_centered(q[j], mm[j], qp);
                    base = (kind == SPQ_OPERAND_CODE || kind == SPQ_OPERAND_CODE_E4M3) ? cen : __fmul_rn(cen, mm[j].s);
                } else {
                    base = (kind == SPQ_OPERAND_CODE) ? q[j] : log_value(xv[j], q[j], lg[j], qp, inv_lev);
                }
                o[j] = base * cmv[j];
            }
            const long long off = row * a.K + c0;
            if (kind == SPQ_OPERAND_CODE_E4M3)
                *reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(a.a_q) + off) = pack_e4m3x4(o[0], o[1], o[2], o[3]);
            else
                *reinterpret_cast<uint2*>(a.a_q + off) = make_uint2(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]));
            if (a.a_raw) {
                const float4 rm4 = *reinterpret_cast<const float4*>(s_rm + c0);
                *reinterpret_cast<uint2*>(a.a_raw + off) =
                    make_uint2(pack_h2(xv[0] * rm4.x, xv[1] * rm4.y), pack_h2(xv[2] * rm4.z, xv[3] * rm4.w));
            }
        }
    }
}

// (b) LayerNorm -> row-scaled fp16 operand (+ per-column statistics partials).  Shared: w, b [Kp]; at the end the
// eight warps fold their column statistics through 2 x [LN_WARPS][Kp] floats of the same buffer.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_rowscale_stats_kernel(LnArgs a) {
    extern __shared__ __align__(16) float s_tab[];
    constexpr int Kp = NV * 128;
    float* s_w = s_tab;
    float* s_b = s_tab + Kp;
    float* s_mn = s_tab + 2 * Kp;                // [LN_WARPS][Kp], statistics only
    float* s_mx = s_mn + LN_WARPS * Kp;
    __shared__ int s_any;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < Kp; c += LN_WARPS * 32) {
        s_w[c] = (c < a.K) ? __ldg(a.ln_w + c) : 0.f;
        s_b[c] = (c < a.K) ? __ldg(a.ln_b + c) : 0.f;
    }
    if (tid == 0) s_any = 0;
    __syncthreads();
    float mn[NV][4], mx[NV][4];
    unsigned long long nan_cols = 0ull;   // bit (4 i + j): this lane's column saw a NaN (fminf / fmaxf drop it, torch keeps it)
    bool any = false;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { mn[i][j] = INFINITY; mx[i][j] = -INFINITY; }
    }
    const bool log_stats = a.stats_mode == 2;
    for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.M; row += static_cast<long long>(gridDim.x) * LN_WARPS) {
        float4 v[NV];
        ln_warp_normalise<NV>(a, row, lane, s_w, s_b, v);
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c0 = (i * 32 + lane) * 4;
            if (c0 >= a.K) continue;
            const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(e[0]), fabsf(e[1])), fmaxf(fabsf(e[2]), fabsf(e[3]))));
            if (a.stats_mode) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float s = log_stats ? fabsf(e[j]) : e[j];
                    nan_cols |= (s != s) ? (1ull << (4 * i + j)) : 0ull;
                    if (log_stats) any |= (s > a.stat_eps);
                    mn[i][j] = fminf(mn[i][j], s);
                    mx[i][j] = fmaxf(mx[i][j], s);
                }
            }
        }
        amax = warp_fmax(amax);
        // the scale rule of rowscale_kernel: amax in [2^(E-1), 2^E) -> x * 2^(8-E); zero rows: smallest scale; inf / NaN: 1
        int E = 0;
        if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = (amax == 0.f) ? -100 : 8;
        E = E < -100 ? -100 : E;
        const float down = exp2f(static_cast<float>(8 - E));
        if (lane == 0 && a.row_scale) a.row_scale[row] = exp2f(static_cast<float>(E - 8));
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c0 = (i * 32 + lane) * 4;
            if (c0 < a.K)
                *reinterpret_cast<uint2*>(a.x16 + row * a.K + c0) =
                    make_uint2(pack_h2(v[i].x * down, v[i].y * down), pack_h2(v[i].z * down, v[i].w * down));
        }
    }
    if (a.stats_mode) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c0 = (i * 32 + lane) * 4;
            float lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool bad = (nan_cols >> (4 * i + j)) & 1ull;
                lo[j] = bad ? NAN : mn[i][j];
                hi[j] = bad ? NAN : mx[i][j];
            }
            *reinterpret_cast<float4*>(s_mn + warp * Kp + c0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(s_mx + warp * Kp + c0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
        if (log_stats && any) s_any = 1;
        __syncthreads();
        for (int c = tid; c < a.K; c += LN_WARPS * 32) {
            float lo = s_mn[c], hi = s_mx[c];
#pragma unroll
            for (int w = 1; w < LN_WARPS; ++w) { lo = nan_min(lo, s_mn[w * Kp + c]); hi = nan_max(hi, s_mx[w * Kp + c]); }
            a.pmin[static_cast<long long>(blockIdx.x) * a.K + c] = lo;
            a.pmax[static_cast<long long>(blockIdx.x) * a.K + c] = hi;
        }
        if (log_stats && tid == 0 && s_any) atomicOr(a.flags, 1);
    }
}

// float4 column groups per lane: the smallest instantiated NV with 128 NV >= K
static int ln_fused_nv(long long K) {
    if ((K % 4) != 0 || K < 4) return 0;
    const int opts[] = {2, 4, 6, 8, 13, 16};
    for (int nv : opts)
        if (128ll * nv >= K) return nv;
    return 0;
}
static size_t ln_quant_smem(int NV) { return static_cast<size_t>(9) * NV * 128 * sizeof(float); }
static size_t ln_stats_smem(int NV, bool stats) { return static_cast<size_t>(2 + (stats ? 2 * LN_WARPS : 0)) * NV * 128 * sizeof(float); }
constexpr int LN_MAX_CTAS_PER_SM = 8;
// persistent grid: as many CTAs as are co-resident (registers and the shared table decide), never more than the rows need
template <typename Kern>
static int ln_launch(Kern kern, size_t smem, LnArgs& a, cudaStream_t st, unsigned* grid_out) {
    if (smem > 48 * 1024) SPQ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    SPQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LN_WARPS * 32, smem));
    static int cap_env = -1;                    // SPQ_LN_CTAS_PER_SM: tuning switch
    if (cap_env < 0) { const char* e = getenv("SPQ_LN_CTAS_PER_SM"); cap_env = e ? atoi(e) : 0; }
    if (cap_env > 0 && per_sm > cap_env) per_sm = cap_env;
    if (per_sm > LN_MAX_CTAS_PER_SM) per_sm = LN_MAX_CTAS_PER_SM;
    if (per_sm < 1) per_sm = 1;
    long long ctas = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * per_sm;
    const long long need = (a.M + LN_WARPS - 1) / LN_WARPS;
    if (ctas > need) ctas = need;
    const unsigned grid = static_cast<unsigned>(ctas < 1 ? 1 : ctas);
    if (a.pmin) a.pmax = a.pmin + static_cast<size_t>(grid) * a.K;       // partials are [grid, K]
    kern<<<grid, LN_WARPS * 32, smem, st>>>(a);
    SPQ_LAUNCH_OK();
    if (grid_out) *grid_out = grid;
    return SPQ_OK;
}

template <int QTYPE>
static int launch_act(const ActArgs& a, cudaStream_t st) {
    const unsigned gx = static_cast<unsigned>((a.K / 4 + 31) / 32);
    long long gy = (static_cast<long long>(sm_count()) * 8 + gx - 1) / gx;
    const long long max_gy = (a.M + 31) / 32;            // at least 4 rows per thread
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const dim3 grid(gx, static_cast<unsigned>(gy)), block(32, 8);
    constexpr int HOT_KIND = (QTYPE == SPQ_MINMAX) ? SPQ_OPERAND_CODE : SPQ_OPERAND_DEQUANT;
    // (a variant with the per-column constants in shared memory -- 64 registers, 8 warps per sub-partition instead of 6 --
    // measured the same 50-51 us at 32768 x 768: the kernel is not occupancy-bound; profiles/r02b_quantize_act_smem_table.txt)
    if (a.qp.symmetric && a.operand_kind == HOT_KIND) {
        if (a.x_half) quantize_act_kernel<QTYPE, __half, 1, HOT_KIND><<<grid, block, 0, st>>>(a);
        else quantize_act_kernel<QTYPE, float, 1, HOT_KIND><<<grid, block, 0, st>>>(a);
    } else {
        if (a.x_half) quantize_act_kernel<QTYPE, __half><<<grid, block, 0, st>>>(a);
        else quantize_act_kernel<QTYPE, float><<<grid, block, 0, st>>>(a);
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

extern "C" int spq_quantize_act(const void* x, int x_is_half, int64_t M, int64_t K, const float* scale, const float* zero_point,
                                int bcast, int qtype, int bits, int symmetric, int operand_kind, const float* col_mul, float mul,
                                spq_half_t* a_q, spq_half_t* a_raw, const float* raw_col_mul, spq_stream_t stream) {
    SPQ_REQUIRE(x && M > 0 && K > 0, "spq_quantize_act: bad input");
    SPQ_REQUIRE((K % 4) == 0 && aligned16(x), "spq_quantize_act: K must be a multiple of 4 and x 16-byte aligned");
    SPQ_REQUIRE(qtype == SPQ_MINMAX || qtype == SPQ_LOG, "spq_quantize_act: unknown quantizer type %d", qtype);
    SPQ_REQUIRE(scale && zero_point && a_q, "spq_quantize_act: missing parameters or output");
    SPQ_REQUIRE(bcast == SPQ_PER_COL || bcast == SPQ_PER_TENSOR, "spq_quantize_act: per-row scales are not an activation layout");
    SPQ_REQUIRE(bits >= 1 && bits < 32, "spq_quantize_act: bits %d", bits);
    ActArgs a;
    a.dgelu_y = nullptr;
    a.x = x; a.x_half = x_is_half ? 1 : 0; a.M = M; a.K = K; a.scale = scale; a.zp = zero_point; a.bcast = bcast;
    a.qp = make_qparams(bits, symmetric);
    a.operand_kind = operand_kind; a.col_mul = col_mul; a.mul = mul;
    a.a_q = a_q; a.a_raw = a_raw; a.raw_row_scale = nullptr; a.raw_col_mul = raw_col_mul; a.max_scale = nullptr;
    cudaStream_t st = as_stream(stream);
    if (qtype == SPQ_MINMAX) return launch_act<SPQ_MINMAX>(a, st);
    return launch_act<SPQ_LOG>(a, st);
}

static int rowscale_impl(const void* g, int g_is_half, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out,
                         float* row_scale, float* max_scale, spq_stream_t stream) {
    SPQ_REQUIRE(g && out && M > 0 && N > 0, "spq_rowscale_f16: bad arguments");
    if (ld_out <= 0) ld_out = N;
    SPQ_REQUIRE(ld_out >= N, "spq_rowscale_f16: ld_out < N");
    if (max_scale) SPQ_CUDA_OK(cudaMemsetAsync(max_scale, 0, sizeof(float), as_stream(stream)));
    if (ld_out == N && (N % 4) == 0 && N <= 8192 && aligned16(g)) {
        ActArgs a;
    a.dgelu_y = nullptr;
        a.x = g; a.x_half = g_is_half ? 1 : 0; a.M = M; a.K = N; a.scale = nullptr; a.zp = nullptr; a.bcast = SPQ_PER_TENSOR;
        a.qp = make_qparams(8, 1);
        a.operand_kind = SPQ_OPERAND_RAW; a.col_mul = nullptr; a.mul = 1.0f;
        a.a_q = nullptr; a.a_raw = out; a.raw_row_scale = row_scale; a.raw_col_mul = nullptr; a.max_scale = max_scale;
        return launch_rowscale(a, as_stream(stream));
    }
    SPQ_REQUIRE(!g_is_half, "spq_rowscale_f16: float16 input needs dense rows (ld_out == N), N %% 4 == 0, N <= 8192, 16-byte alignment");
    SPQ_REQUIRE((ld_out % 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0, "spq_rowscale_f16: ld_out must be even");
    long long ctas = static_cast<long long>(sm_count()) * 8;
    if (ctas > M) ctas = M;
    rowscale_wide_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(static_cast<const float*>(g), M, N, out, ld_out,
                                                                                    row_scale, max_scale);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_rowscale_f16(const void* g, int g_is_half, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out,
                                float* row_scale, spq_stream_t stream) {
    return rowscale_impl(g, g_is_half, M, N, out, ld_out, row_scale, nullptr, stream);
}

extern "C" int spq_rowscale_f16_max(const void* g, int g_is_half, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out,
                                    float* row_scale, float* max_scale, spq_stream_t stream) {
    SPQ_REQUIRE(max_scale, "spq_rowscale_f16_max: null max_scale");
    return rowscale_impl(g, g_is_half, M, N, out, ld_out, row_scale, max_scale, stream);
}

// Operands of the LoRA gradient GEMMs from the down-projected gradient (SURVEY section 8 a13 backward):
//   dt16[m,j] = fp16(dtn[m,j] * dt_mul)                     -> dX += dt q(A)^T      (token scale eg[m] in the epilogue)
//   dt2 [m,j] = fp16(dtn[m,j] * dt_mul * eg[m] / gmax)      -> dA  = x^T dt         (token scale folded: reduction over m)
//   t2  [m,j] = fp16(t16[m,j] * eg[m] / gmax)               -> dB  = t^T dY
// eg[m] are the power-of-two row scales of the fp16 gradient operand and gmax their maximum (both from
// spq_rowscale_f16_max), so eg / gmax <= 1 is a power of two and the products are exact up to fp16 underflow.
__global__ void __launch_bounds__(256)
lora_bwd_prep_kernel(const float* __restrict__ dtn, const unsigned short* __restrict__ t16, long long ld_t,
                     const float* __restrict__ eg, const float* __restrict__ gmax, long long M, int r, float dt_mul,
                     unsigned short* __restrict__ dt16, unsigned short* __restrict__ dt2, unsigned short* __restrict__ t2) {
    const float inv_gmax = 1.0f / __ldg(gmax);
    const int per_row = r >> 2;                                  // float4 groups per row
    const long long total = M * per_row;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long m = i / per_row;
        const int c = static_cast<int>(i - m * per_row) * 4;
        const float e = __ldg(eg + m) * inv_gmax;
        if (dtn) {
            const float4 v = ld_stream_f4(dtn + m * r + c);
            const float a0 = v.x * dt_mul, a1 = v.y * dt_mul, a2 = v.z * dt_mul, a3 = v.w * dt_mul;
            if (dt16) *reinterpret_cast<uint2*>(dt16 + m * r + c) = make_uint2(pack_h2(a0, a1), pack_h2(a2, a3));
            if (dt2) *reinterpret_cast<uint2*>(dt2 + m * r + c) = make_uint2(pack_h2(a0 * e, a1 * e), pack_h2(a2 * e, a3 * e));
        }
        if (t16 && t2) {
            const uint2 h = *reinterpret_cast<const uint2*>(t16 + m * ld_t + c);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            *reinterpret_cast<uint2*>(t2 + m * r + c) = make_uint2(pack_h2(lo.x * e, lo.y * e), pack_h2(hi.x * e, hi.y * e));
        }
    }
}

extern "C" int spq_lora_bwd_prep(const float* dtn, const spq_half_t* t16, int64_t ld_t16, const float* row_scale,
                                 const float* max_scale, int64_t M, int64_t r, float dt_mul, spq_half_t* dt16, spq_half_t* dt2,
                                 spq_half_t* t2, spq_stream_t stream) {
    SPQ_REQUIRE(row_scale && max_scale && M > 0 && r > 0 && (r % 4) == 0, "spq_lora_bwd_prep: bad arguments (rank must be a multiple of 4)");
    SPQ_REQUIRE((!dt16 || (reinterpret_cast<uintptr_t>(dt16) & 7u) == 0) && (!dt2 || (reinterpret_cast<uintptr_t>(dt2) & 7u) == 0) && (!t2 || (reinterpret_cast<uintptr_t>(t2) & 7u) == 0) && (!t16 || (reinterpret_cast<uintptr_t>(t16) & 7u) == 0), "spq_lora_bwd_prep: operands must be 8-byte aligned");
    SPQ_REQUIRE((dtn && (dt16 || dt2)) || (t16 && t2), "spq_lora_bwd_prep: nothing to do");
    SPQ_REQUIRE(!dtn || aligned16(dtn), "spq_lora_bwd_prep: dtn must be 16-byte aligned");
    SPQ_REQUIRE(!t16 || ((ld_t16 % 4) == 0 && ld_t16 >= r), "spq_lora_bwd_prep: bad t16 leading dimension");
    const long long total = M * (r / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    lora_bwd_prep_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(dtn, t16, ld_t16, row_scale, max_scale, M,
                                                                                     static_cast<int>(r), dt_mul, dt16, dt2, t2);
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

extern "C" int spq_ln_quantize_act(const float* x, int64_t M, int64_t K, const float* ln_weight, const float* ln_bias, float ln_eps,
                                   const float* scale, const float* zero_point, int bcast, int qtype, int bits, int symmetric,
                                   int operand_kind, const float* col_mul, float mul, spq_half_t* a_q, spq_half_t* a_raw,
                                   const float* raw_col_mul, float* y_out, spq_stream_t stream) {
    SPQ_REQUIRE(x && ln_weight && ln_bias && M > 0 && K > 0, "spq_ln_quantize_act: bad input");
    SPQ_REQUIRE(aligned16(x) && (!y_out || aligned16(y_out)), "spq_ln_quantize_act: x and y must be 16-byte aligned");
    SPQ_REQUIRE(qtype == SPQ_MINMAX || qtype == SPQ_LOG, "spq_ln_quantize_act: unknown quantizer type %d", qtype);
    SPQ_REQUIRE(scale && zero_point && a_q, "spq_ln_quantize_act: missing parameters or output");
    SPQ_REQUIRE(bcast == SPQ_PER_COL || bcast == SPQ_PER_TENSOR, "spq_ln_quantize_act: per-row scales are not an activation layout");
    SPQ_REQUIRE(bits >= 1 && bits < 32, "spq_ln_quantize_act: bits %d", bits);
    SPQ_REQUIRE((reinterpret_cast<uintptr_t>(a_q) & 7u) == 0 && (!a_raw || (reinterpret_cast<uintptr_t>(a_raw) & 7u) == 0),
                "spq_ln_quantize_act: operands must be 8-byte aligned");
    const int NV = ln_fused_nv(K);
    if (!NV) {
        set_error("spq_ln_quantize_act: normalized dim %lld unsupported (multiple of 4, <= 2048)", (long long)K);
        return SPQ_ERR_UNSUPPORTED;
    }
    LnArgs a = {};
    a.x = x; a.M = M; a.K = K; a.ln_w = ln_weight; a.ln_b = ln_bias; a.ln_eps = ln_eps; a.y = y_out;
    a.scale = scale; a.zp = zero_point; a.bcast = bcast; a.qp = make_qparams(bits, symmetric);
    a.operand_kind = operand_kind; a.col_mul = col_mul; a.mul = mul; a.a_q = a_q; a.a_raw = a_raw; a.raw_col_mul = raw_col_mul;
    const size_t smem = ln_quant_smem(NV);
    cudaStream_t st = as_stream(stream);
#define SPQ_LNQ(QT, N) return ln_launch(ln_quantize_act_kernel<QT, N>, smem, a, st, nullptr)
    if (qtype == SPQ_MINMAX) {
        switch (NV) { case 2: SPQ_LNQ(SPQ_MINMAX, 2); case 4: SPQ_LNQ(SPQ_MINMAX, 4); case 6: SPQ_LNQ(SPQ_MINMAX, 6);
                      case 8: SPQ_LNQ(SPQ_MINMAX, 8); case 13: SPQ_LNQ(SPQ_MINMAX, 13); default: SPQ_LNQ(SPQ_MINMAX, 16); }
    }
    switch (NV) { case 2: SPQ_LNQ(SPQ_LOG, 2); case 4: SPQ_LNQ(SPQ_LOG, 4); case 6: SPQ_LNQ(SPQ_LOG, 6);
                  case 8: SPQ_LNQ(SPQ_LOG, 8); case 13: SPQ_LNQ(SPQ_LOG, 13); default: SPQ_LNQ(SPQ_LOG, 16); }
#undef SPQ_LNQ
}

extern "C" size_t spq_ln_rowscale_stats_workspace_bytes(int64_t M, int64_t K) {
    const int NV = ln_fused_nv(K);
    if (M <= 0 || K <= 0 || !NV) return 256;
    // [grid, K] minima and maxima; the grid is decided at launch (occupancy), bounded by LN_MAX_CTAS_PER_SM CTAs per SM
    long long ctas = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * LN_MAX_CTAS_PER_SM;
    const long long need = (M + LN_WARPS - 1) / LN_WARPS;
    if (ctas > need) ctas = need;
    return 256 + 2 * static_cast<size_t>(ctas) * static_cast<size_t>(K) * sizeof(float);
}

extern "C" int spq_ln_rowscale_stats(const float* x, int64_t M, int64_t K, const float* ln_weight, const float* ln_bias, float ln_eps,
                                     spq_half_t* out, float* row_scale, int stats_mode, float stat_eps, float* stat_min,
                                     float* stat_max, int accumulate, int32_t* state, float* y_out, void* workspace,
                                     size_t workspace_bytes, spq_stream_t stream) {
    SPQ_REQUIRE(x && ln_weight && ln_bias && out && row_scale && M > 0 && K > 0, "spq_ln_rowscale_stats: bad input");
    SPQ_REQUIRE(aligned16(x) && (!y_out || aligned16(y_out)) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
                "spq_ln_rowscale_stats: alignment");
    SPQ_REQUIRE(stats_mode >= 0 && stats_mode <= 2, "spq_ln_rowscale_stats: stats_mode %d", stats_mode);
    SPQ_REQUIRE(!stats_mode || (stat_min && stat_max && workspace && aligned16(workspace) &&
                                workspace_bytes >= spq_ln_rowscale_stats_workspace_bytes(M, K)),
                "spq_ln_rowscale_stats: statistics need stat_min / stat_max and a workspace");
    const int NV = ln_fused_nv(K);
    if (!NV) {
        set_error("spq_ln_rowscale_stats: normalized dim %lld unsupported (multiple of 4, <= 2048)", (long long)K);
        return SPQ_ERR_UNSUPPORTED;
    }
    LnArgs a = {};
    a.x = x; a.M = M; a.K = K; a.ln_w = ln_weight; a.ln_b = ln_bias; a.ln_eps = ln_eps; a.y = y_out;
    a.x16 = out; a.row_scale = row_scale; a.stats_mode = stats_mode; a.stat_eps = stat_eps;
    const size_t smem = ln_stats_smem(NV, stats_mode != 0);
    cudaStream_t st = as_stream(stream);
    if (stats_mode) {
        a.flags = reinterpret_cast<int32_t*>(workspace);
        a.pmin = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);      // pmax follows (set at launch)
        SPQ_CUDA_OK(cudaMemsetAsync(a.flags, 0, 16, st));
    }
    unsigned grid = 0;
    int rc;
    switch (NV) {
        case 2: rc = ln_launch(ln_rowscale_stats_kernel<2>, smem, a, st, &grid); break;
        case 4: rc = ln_launch(ln_rowscale_stats_kernel<4>, smem, a, st, &grid); break;
        case 6: rc = ln_launch(ln_rowscale_stats_kernel<6>, smem, a, st, &grid); break;
        case 8: rc = ln_launch(ln_rowscale_stats_kernel<8>, smem, a, st, &grid); break;
        case 13: rc = ln_launch(ln_rowscale_stats_kernel<13>, smem, a, st, &grid); break;
        default: rc = ln_launch(ln_rowscale_stats_kernel<16>, smem, a, st, &grid); break;
    }
    if (rc != SPQ_OK) return rc;
    if (stats_mode)
        return spq::stats::finalize_partials(a.pmin, a.pmax, K, static_cast<int>(grid), stats_mode == 2, stat_eps, accumulate, a.flags,
                                             stat_min, stat_max, state, st);
    return SPQ_OK;
}

extern "C" int spq_rowscale_dgelu_f16_max(const float* g, const float* y, int64_t M, int64_t N, spq_half_t* out, float* row_scale,
                                          float* max_scale, spq_stream_t stream) {
    SPQ_REQUIRE(g && y && out && row_scale && max_scale && M > 0 && N > 0, "spq_rowscale_dgelu_f16_max: bad arguments");
    SPQ_REQUIRE((N % 4) == 0 && N <= 8192 && aligned16(g) && aligned16(y) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
                "spq_rowscale_dgelu_f16_max: dense float32 rows, N %% 4 == 0, N <= 8192, 16-byte aligned inputs");
    SPQ_CUDA_OK(cudaMemsetAsync(max_scale, 0, sizeof(float), as_stream(stream)));
    ActArgs a;
    a.dgelu_y = y;
    a.x = g; a.x_half = 0; a.M = M; a.K = N; a.scale = nullptr; a.zp = nullptr; a.bcast = SPQ_PER_TENSOR;
    a.qp = make_qparams(8, 1);
    a.operand_kind = SPQ_OPERAND_RAW; a.col_mul = nullptr; a.mul = 1.0f;
    a.a_q = nullptr; a.a_raw = out; a.raw_row_scale = row_scale; a.raw_col_mul = nullptr; a.max_scale = max_scale;
    return launch_rowscale(a, as_stream(stream));
}

