// One-launch preparation of every scale vector a fused SPLinearWithLoRA forward needs after its
// input quantiser has been (re)calibrated.  Replaces ~100 tiny elementwise launches per layer
// (exp2 / ceil / where / amax ... on K- and N-sized vectors) that made the calibrate+forward step
// launch-bound.  Single CTA: the vectors are at most a few thousand elements long.
//
// Given the activation quantiser's (scale, zero_point) [K or 1] and the per-row absmax of the
// dequantised weight, it produces
//   absorb[k]   what the weight operand must be multiplied by per input channel
//               (min-max: scale[k] / code_mul;  log: 2^(ceil(log_max[k]) - 8))
//   act_mul[k]  what the activation operand is multiplied by (min-max: unused = 1;  log: 1 / absorb[k])
//   raw_mul[k], inv_raw_mul[k]  power of two mapping the calibrated bound of input channel k to (8, 16]:
//               multiplier of the raw (unquantised) fp16 operand the LoRA branch reads, and its inverse
//   pw[n], inv_pw[n]   power-of-two row normaliser of the weight operand and its reciprocal
//   lora[0:r] = tau, [r:2r] = 1/tau, [2r:3r] = scaling/tau, [3r:4r] = pa, [4r:5r] = 1/pa   with tau the
//               static pre-scale of t = x q(A), chosen from the bound  max_j sum_k xbound[k] |q(A)[k,j]|,
//               and pa[j] the power-of-two normaliser of column j of q(A)[k,j] / raw_mul[k]
#include "spq_common.cuh"

namespace spq {
namespace prep {

__device__ __forceinline__ float pow2_ceil(float x) {       // smallest power of two >= x (x > 0, finite)
    const int b = __float_as_int(x);
    const int e = b & 0x7f800000;
    if (e != 0 && e != 0x7f000000) {                         // normal, and the result stays finite: bump the exponent
        return __int_as_float((b & 0x007fffff) ? e + 0x00800000 : e);      // unless x already is a power of two
    }
    int ex;                                                   // denormal / top binade: the library route
    const float f = frexpf(x, &ex);                           // x = f * 2^ex, f in [0.5, 1)
    return ldexpf(1.0f, f == 0.5f ? ex - 1 : ex);
}

__device__ __forceinline__ float block_max(float v, float* s_red) {
    v = warp_fmax(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = s_red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
    return m;
}

struct PrepArgs {
    const float* in_scale; const float* in_zp; int in_n;     // 1 or K
    int qtype, bits, symmetric;
    long long K, N, r;
    const float* w_rowmax;                                   // [N]
    const float* aq_abs;                                     // [K, r] q(A) (sign ignored) or null
    const float* bq;                                         // [r, N] q(B) or null -> lora[5r:6r] = pb, [6r:7r] = 1/pb
    float lora_scaling;
    float* absorb; float* act_mul; float* raw_mul; float* inv_raw_mul; float* pw; float* inv_pw; float* lora;
};

__device__ __forceinline__ void chan(const PrepArgs& a, long long k, float& absorb, float& act_mul, float& xbound) {
    const float sc = __ldg(a.in_scale + (a.in_n == 1 ? 0 : k));
    const float zp = __ldg(a.in_zp + (a.in_n == 1 ? 0 : k));
    if (a.qtype == SPQ_MINMAX) {
        const int shift = a.bits > 11 ? a.bits - 11 : 0;     // codes beyond 11 bits are pre-scaled to stay in fp16 range
        absorb = ldexpf(sc, shift);
        act_mul = 1.0f;
        if (a.symmetric) {
            xbound = sc * static_cast<float>((1ll << (a.bits - 1)) - 1);
        } else {
            const float full = static_cast<float>((1ll << a.bits) - 1);
            xbound = fmaxf(fabsf(zp), fabsf(full - zp)) * sc;
        }
    } else {
        float lmax = zp + fmaxf(sc, 0.f);                    // log_min + log_range
        lmax = fminf(fmaxf(lmax, -100.f), 100.f);
        const float e = ceilf(lmax);
        absorb = exp2f(e - 8.0f);
        act_mul = exp2f(8.0f - e);
        xbound = exp2f(lmax);
    }
}

__global__ void __launch_bounds__(1024) prep_linear_scales_kernel(PrepArgs a) {
    __shared__ float s_red[32];
    __shared__ float s_t[1024];
    __shared__ float s_xb[4096];        // calibrated bound per input channel (K <= 4096: else recomputed)
    const int tid = threadIdx.x;
    float amax = 0.f;
    for (long long k = tid; k < a.K; k += blockDim.x) {
        float ab, am, xb;
        chan(a, k, ab, am, xb);
        a.absorb[k] = ab;
        a.act_mul[k] = am;
        const float rmul = (xb > 0.f && xb < INFINITY) ? 16.0f / pow2_ceil(xb) : 1.0f;
        a.raw_mul[k] = rmul;
        a.inv_raw_mul[k] = 1.0f / rmul;
        if (k < 4096) s_xb[k] = xb;
        amax = fmaxf(amax, ab);
    }
    const float absorb_max = block_max(amax, s_red);
    for (long long n = tid; n < a.N; n += blockDim.x) {
        const float wm = __ldg(a.w_rowmax + n) * absorb_max;
        const float p = (wm > 0.f && wm < INFINITY) ? pow2_ceil(wm) * 0.00390625f : 1.0f;   // row max -> (128, 256]
        a.pw[n] = p;
        a.inv_pw[n] = 1.0f / p;
    }
    if (a.aq_abs != nullptr && a.r > 0) {
        // per LoRA column j:  tsum[j] = sum_k xbound[k] |Aq[k,j]|,  amx[j] = max_k |Aq[k,j]| / raw_mul[k]
        const long long r = a.r;
        __shared__ float s_m[1024];
        float tmax = 0.f;
        if ((r % 4) == 0 && r <= 4096 && (1024 % (r / 4)) == 0 && aligned16_dev(a.aq_abs)) {
            // thread -> 4 consecutive LoRA columns (one float4 per k), k strided; 4 loads in flight
            const long long g = r / 4, jg = (tid % g) * 4, k0 = tid / g, kstep = 1024 / g;
            float acc[4] = {0.f, 0.f, 0.f, 0.f}, mx[4] = {0.f, 0.f, 0.f, 0.f};
            auto fold = [&](long long k, const float4& av) {
                float xb;
                if (k < 4096) xb = s_xb[k];
                else { float ab, am; chan(a, k, ab, am, xb); }
                const float inv_rmul = (xb > 0.f && xb < INFINITY) ? pow2_ceil(xb) * 0.0625f : 1.0f;
                acc[0] += xb * av.x; acc[1] += xb * av.y; acc[2] += xb * av.z; acc[3] += xb * av.w;
                mx[0] = fmaxf(mx[0], av.x * inv_rmul); mx[1] = fmaxf(mx[1], av.y * inv_rmul);
                mx[2] = fmaxf(mx[2], av.z * inv_rmul); mx[3] = fmaxf(mx[3], av.w * inv_rmul);
            };
            long long k = k0;
            for (; k + 3 * kstep < a.K; k += 4 * kstep) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(a.aq_abs + (k + u * kstep) * r + jg);
#pragma unroll
                for (int u = 0; u < 4; ++u) fold(k + u * kstep, make_float4(fabsf(v[u].x), fabsf(v[u].y), fabsf(v[u].z), fabsf(v[u].w)));
            }
            for (; k < a.K; k += kstep) {
                const float4 q = *reinterpret_cast<const float4*>(a.aq_abs + k * r + jg);
                fold(k, make_float4(fabsf(q.x), fabsf(q.y), fabsf(q.z), fabsf(q.w)));
            }
            // fold the kstep partial results of each column: column j lives in threads with tid % g == j / 4
            __shared__ float s_acc[4][1024];
#pragma unroll
            for (int u = 0; u < 4; ++u) { s_acc[u][tid] = acc[u]; }
            __syncthreads();
            float colsum = 0.f;
            if (tid < r) {
                const long long gj = tid / 4, u = tid % 4;
                for (long long q = gj; q < 1024; q += g) colsum += s_acc[u][q];
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) { s_acc[u][tid] = mx[u]; }
            __syncthreads();
            if (tid < r) {
                const long long gj = tid / 4, u = tid % 4;
                float m = 0.f;
                for (long long q = gj; q < 1024; q += g) m = fmaxf(m, s_acc[u][q]);
                tmax = colsum;
                const float pa = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
                a.lora[3 * r + tid] = pa;
                a.lora[4 * r + tid] = 1.0f / pa;
            }
        } else if (r <= 1024 && (1024 % r) == 0) {
            const long long j = tid % r, k0 = tid / r, kstep = 1024 / r;
            float acc = 0.f, mx = 0.f;
            for (long long k = k0; k < a.K; k += kstep) {
                float xb;
                if (k < 4096) xb = s_xb[k];
                else { float ab, am; chan(a, k, ab, am, xb); }
                const float av = fabsf(__ldg(a.aq_abs + k * r + j));
                const float inv_rmul = (xb > 0.f && xb < INFINITY) ? pow2_ceil(xb) * 0.0625f : 1.0f;
                acc += xb * av;
                mx = fmaxf(mx, av * inv_rmul);
            }
            s_t[tid] = acc;
            s_m[tid] = mx;
            __syncthreads();
            if (tid < r) {
                float t = 0.f, m = 0.f;
                for (long long q = tid; q < 1024; q += r) { t += s_t[q]; m = fmaxf(m, s_m[q]); }
                tmax = t;
                const float pa = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
                a.lora[3 * r + tid] = pa;
                a.lora[4 * r + tid] = 1.0f / pa;
            }
        } else {
            for (long long j = tid; j < r; j += blockDim.x) {
                float acc = 0.f, m = 0.f;
                for (long long k = 0; k < a.K; ++k) {
                    float ab, am, xb;
                    chan(a, k, ab, am, xb);
                    const float av = fabsf(__ldg(a.aq_abs + k * r + j));
                    const float rmul = (xb > 0.f && xb < INFINITY) ? 16.0f / pow2_ceil(xb) : 1.0f;
                    acc += xb * av;
                    m = fmaxf(m, av / rmul);
                }
                tmax = fmaxf(tmax, acc);
                const float pa = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
                a.lora[3 * r + j] = pa;
                a.lora[4 * r + j] = 1.0f / pa;
            }
        }
        if (a.bq != nullptr) {
            // pb[j] = power-of-two normaliser of row j of scaling * q(B)  (operand of dT = dY q(B)^T)
            const int warp = tid >> 5, lane = tid & 31;
            for (long long j = warp; j < r; j += 32) {
                float m = 0.f;
                for (long long n = lane; n < a.N; n += 32) m = fmaxf(m, fabsf(__ldg(a.bq + j * a.N + n)));
                m = warp_fmax(m) * fabsf(a.lora_scaling);
                if (lane == 0) {
                    const float pb = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
                    a.lora[5 * r + j] = pb;
                    a.lora[6 * r + j] = 1.0f / pb;
                }
            }
        }
        tmax = block_max(tmax, s_red);
        const float tau = (tmax > 0.f && tmax < INFINITY) ? 16384.0f / pow2_ceil(tmax) : 1.0f;
        for (long long j = tid; j < r; j += blockDim.x) {
            a.lora[j] = tau;
            a.lora[r + j] = 1.0f / tau;
            a.lora[2 * r + j] = a.lora_scaling / tau;
            a.lora[7 * r + j] = a.lora[3 * r + j] * tau;      // pa * tau (block_max's barrier ordered the pa writes)
        }
    }
}


// CPT variant (p2/cpt_model.py:92-114): every scale vector of the shared-LoRA level of one CPTLinear in ONE launch.  The
// adapter reads the QUANTISED input (its per-K factor `absorb` goes into the A operand) and B is stored [N, r].
//   out[0r:1r] pa[j]      power-of-two normaliser of column j of q(A)[k,j] * absorb[k]   (operand values in (1/2, 1])
//   out[1r:2r] 1 / pa
//   out[2r:3r] tau        2^14 / pow2ceil(max_j sum_k xbound[k] |q(A)[k,j]|): multiplier of the fp16 down-projection
//   out[3r:4r] 1 / tau
//   out[4r:5r] pa * tau   epilogue scale of the down-projection GEMM
//   out[5r:6r] scaling / tau   column multiplier of the up-projection operand
//   out[6r:7r] pb[j]      power-of-two normaliser of column j of scaling * q(B)[n,j]     (operand of dT = dY q(B))
//   out[7r:8r] 1 / pb
// The shared adapter changes at every optimizer step and the width at every step (configs[3]): this replaces ~85
// eager launches per linear per step (abs / amax / where / log2 / ceil / exp2 / ... on [K, r] and [N, r]).
__global__ void __launch_bounds__(1024)
cpt_lora_scales_kernel(const float* __restrict__ aq, const float* __restrict__ bq, const float* __restrict__ absorb,
                       const float* __restrict__ xbound, long long K, long long N, int r, float scaling, float* __restrict__ out) {
    __shared__ float s_a[1024], s_b[1024], s_red[32];
    const int tid = threadIdx.x;
    const int j = tid % r, lane0 = tid / r, lanes = 1024 / r;      // r divides 1024
    float amx = 0.f, tsum = 0.f, bmx = 0.f;
    for (long long k = lane0; k < K; k += lanes) {
        const float av = fabsf(__ldg(aq + k * r + j));
        amx = fmaxf(amx, av * __ldg(absorb + k));
        tsum = fmaf(__ldg(xbound + k), av, tsum);
    }
    for (long long n = lane0; n < N; n += lanes) bmx = fmaxf(bmx, fabsf(__ldg(bq + n * r + j)));
    s_a[tid] = amx; s_b[tid] = tsum;
    __syncthreads();
    float pa = 1.f, tj = 0.f;
    if (tid < r) {
        float m = 0.f, t = 0.f;
        for (int q = tid; q < 1024; q += r) { m = fmaxf(m, s_a[q]); t += s_b[q]; }
        pa = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
        tj = t;
    }
    __syncthreads();
    s_a[tid] = bmx;
    __syncthreads();
    float pb = 1.f;
    if (tid < r) {
        float m = 0.f;
        for (int q = tid; q < 1024; q += r) m = fmaxf(m, s_a[q]);
        m *= fabsf(scaling);
        pb = (m > 0.f && m < INFINITY) ? pow2_ceil(m) : 1.0f;
    }
    const float tmax = block_max(tid < r ? tj : 0.f, s_red);
    const float tau = (tmax > 0.f && tmax < INFINITY) ? 16384.0f / pow2_ceil(tmax) : 1.0f;
    if (tid < r) {
        out[tid] = pa;               out[r + tid] = 1.0f / pa;
        out[2 * r + tid] = tau;      out[3 * r + tid] = 1.0f / tau;
        out[4 * r + tid] = pa * tau; out[5 * r + tid] = scaling / tau;
        out[6 * r + tid] = pb;       out[7 * r + tid] = 1.0f / pb;
    }
}

}  // namespace prep
}  // namespace spq

using namespace spq;

extern "C" int spq_prep_linear_scales(const float* in_scale, const float* in_zero_point, int64_t in_n, int qtype, int bits,
                                      int symmetric, int64_t K, const float* w_rowmax, int64_t N, const float* aq_abs, const float* bq,
                                      int64_t r, float lora_scaling, float* absorb, float* act_mul, float* raw_mul,
                                      float* inv_raw_mul, float* pw, float* inv_pw, float* lora_vec, spq_stream_t stream) {
    SPQ_REQUIRE(in_scale && in_zero_point && w_rowmax && absorb && act_mul && raw_mul && inv_raw_mul && pw && inv_pw,
                "spq_prep_linear_scales: null pointer");
    SPQ_REQUIRE(K > 0 && N > 0 && (in_n == 1 || in_n == K), "spq_prep_linear_scales: input scale must have 1 or K elements");
    SPQ_REQUIRE(bits >= 1 && bits < 32 && (qtype == SPQ_MINMAX || qtype == SPQ_LOG), "spq_prep_linear_scales: bad quantiser");
    SPQ_REQUIRE(aq_abs == nullptr || (r > 0 && lora_vec), "spq_prep_linear_scales: LoRA outputs missing");
    prep::PrepArgs a;
    a.in_scale = in_scale; a.in_zp = in_zero_point; a.in_n = static_cast<int>(in_n);
    a.qtype = qtype; a.bits = bits; a.symmetric = symmetric; a.K = K; a.N = N; a.r = aq_abs ? r : 0;
    a.w_rowmax = w_rowmax; a.aq_abs = aq_abs; a.bq = aq_abs ? bq : nullptr; a.lora_scaling = lora_scaling;
    a.absorb = absorb; a.act_mul = act_mul; a.raw_mul = raw_mul; a.inv_raw_mul = inv_raw_mul;
    a.pw = pw; a.inv_pw = inv_pw; a.lora = lora_vec;
    prep::prep_linear_scales_kernel<<<1, 1024, 0, as_stream(stream)>>>(a);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_cpt_lora_scales(const float* aq, const float* bq, const float* absorb, const float* xbound, int64_t K, int64_t N,
                                   int64_t r, float scaling, float* out, spq_stream_t stream) {
    SPQ_REQUIRE(aq && bq && absorb && xbound && out && K > 0 && N > 0, "spq_cpt_lora_scales: bad arguments");
    SPQ_REQUIRE(r > 0 && r <= 1024 && (1024 % r) == 0, "spq_cpt_lora_scales: rank must divide 1024");
    prep::cpt_lora_scales_kernel<<<1, 1024, 0, as_stream(stream)>>>(aq, bq, absorb, xbound, K, N, static_cast<int>(r), scaling, out);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}
