// Next-token cross-entropy forward over a [M, V] logits matrix (row stride `ld`), one pass.
//
// Not part of the quantised-linear hot path proper: this is the first consumer after it (SURVEY.md
// section 8 f1).  The reference computes the loss with nn.CrossEntropyLoss on shifted, re-materialised
// copies of the [B, T, 50257] logits (p1/models_sp.py:441-449): ~5 full passes over 6.6 GB at
// 32 x 1024 tokens.  Here each row is read once: online max / sum-of-exponentials per thread,
// block combine, loss_row = log(sum) + max - logit[target].  Forward only (no-grad evaluation /
// calibration); training keeps torch's differentiable implementation.
#include <stdlib.h>

#include "spq_common.cuh"

namespace spq {
namespace loss {

// exp(x) as ONE multiply + ex2.approx.ftz: `__expf` without -use_fast_math is the non-ftz ex2 (range test, two scaling
// multiplies around the MUFU: ~5 instructions), which made the softmax passes issue-bound (the same finding as the LM
// head's log-sum-exp epilogue).  Results below 2^-126 flush to zero: probabilities < 1e-38.
__device__ __forceinline__ float fexp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

struct MS { float m, s; };                                   // running max and sum of exp(x - m)
__device__ __forceinline__ MS combine(MS a, MS b) {
    if (b.m == -INFINITY) return a;
    if (a.m == -INFINITY) return b;
    MS r;
    r.m = fmaxf(a.m, b.m);
    r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
    return r;
}
__device__ __forceinline__ void push(MS& a, float x) {
    if (x > a.m) { a.s = a.s * __expf(a.m - x) + 1.0f; a.m = x; }
    else a.s += __expf(x - a.m);
}

__global__ void __launch_bounds__(256)
cross_entropy_fwd_kernel(const float* __restrict__ logits, long long M, long long V, long long ld,
                         const long long* __restrict__ targets, long long ignore_index, float* __restrict__ row_loss,
                         float* __restrict__ row_valid) {
    __shared__ float sm[8], ss[8];
    const int tid = threadIdx.x;
    for (long long row = blockIdx.x; row < M; row += gridDim.x) {
        const long long tgt = targets[row];
        if (tgt == ignore_index || tgt < 0 || tgt >= V) {       // block-uniform
            if (tid == 0) { row_loss[row] = 0.f; row_valid[row] = 0.f; }
            continue;
        }
        const float* p = logits + row * ld;
        MS acc; acc.m = -INFINITY; acc.s = 0.f;
        const long long v4 = ((ld & 3) == 0 && aligned16_dev(logits)) ? (V >> 2) : 0;
        for (long long i = tid; i < v4; i += 256) {
            const float4 x = ld_stream_f4(p + 4 * i);
            const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
            if (mx > acc.m) { acc.s *= __expf(acc.m - mx); acc.m = mx; }
            acc.s += __expf(x.x - acc.m) + __expf(x.y - acc.m) + __expf(x.z - acc.m) + __expf(x.w - acc.m);
        }
        for (long long i = 4 * v4 + tid; i < V; i += 256) push(acc, __ldg(p + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MS b; b.m = __shfl_xor_sync(0xffffffffu, acc.m, o); b.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
            acc = combine(acc, b);
        }
        __syncthreads();
        if ((tid & 31) == 0) { sm[tid >> 5] = acc.m; ss[tid >> 5] = acc.s; }
        __syncthreads();
        if (tid == 0) {
            MS t; t.m = sm[0]; t.s = ss[0];
            for (int w = 1; w < 8; ++w) { MS b; b.m = sm[w]; b.s = ss[w]; t = combine(t, b); }
            row_loss[row] = logf(t.s) + t.m - __ldg(p + tgt);
            row_valid[row] = 1.f;
        }
    }
}

// Cross-entropy from the (max, sum exp) pairs the LM-head GEMM left per row and column half-tile
// (spq_qgemm_lse): one warp per row folds the P pairs and reads the single target logit -- the 6.6 GB logits
// matrix is not read again.
__global__ void __launch_bounds__(256)
cross_entropy_from_parts_kernel(const float2* __restrict__ parts, long long P, long long part_ld,
                                const float* __restrict__ logits, long long M, long long V, long long ld,
                                const long long* __restrict__ targets, long long ignore_index,
                                float* __restrict__ row_loss, float* __restrict__ row_valid) {
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    for (long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
        const long long tgt = targets[row];
        if (tgt == ignore_index || tgt < 0 || tgt >= V) {       // warp-uniform
            if (lane == 0) { row_loss[row] = 0.f; row_valid[row] = 0.f; }
            continue;
        }
        MS acc; acc.m = -INFINITY; acc.s = 0.f;
        for (long long i = lane; i < P; i += 32) {
            const float2 v = parts[row * part_ld + i];
            MS b; b.m = v.x; b.s = v.y;
            acc = combine(acc, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MS b; b.m = __shfl_xor_sync(0xffffffffu, acc.m, o); b.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
            acc = combine(acc, b);
        }
        if (lane == 0) {
            row_loss[row] = logf(acc.s) + acc.m - __ldg(logits + row * ld + tgt);
            row_valid[row] = 1.f;
        }
    }
}

// Distillation loss of the switchable-precision training step (p1/distillation_manager.py:64-80):
//   KL( softmax(t / T) || softmax(s / T) ) per row, forward value AND the gradient w.r.t. the student logits,
// in two passes over the two [M, V] matrices instead of torch's log_softmax x2 + kl_div + their backward
// (~10 passes of 1.6 GB each at 32 x 256 x 50257).  One CTA per row: pass A = online (max, sum exp) of both
// rows, pass B = loss terms and  grad[v] = g * (ps[v] - pt[v])  (the second read mostly hits L2).
// Rows with  row % seq_len == seq_len - 1  (the last position of each sequence, excluded upstream) are skipped.
__global__ void __launch_bounds__(256)
distill_kl_kernel(const float* __restrict__ s_logits, long long ld_s, const float* __restrict__ t_logits, long long ld_t,
                  long long M, long long V, float inv_T, long long seq_len, float g, float* __restrict__ row_loss,
                  float* __restrict__ grad) {
    __shared__ float sm[4][8];
    __shared__ float s_lse[2];
    __shared__ float s_part[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long row = blockIdx.x; row < M; row += gridDim.x) {
        float* gr = grad ? grad + row * V : nullptr;
        if (seq_len > 0 && (row % seq_len) == seq_len - 1) {         // block-uniform
            if (tid == 0) row_loss[row] = 0.f;
            if (gr) for (long long i = tid; i < V; i += 256) gr[i] = 0.f;
            continue;
        }
        const float* ps_ = s_logits + row * ld_s;
        const float* pt_ = t_logits + row * ld_t;
        MS a; a.m = -INFINITY; a.s = 0.f;
        MS b; b.m = -INFINITY; b.s = 0.f;
        const bool vec = ((ld_s & 3) == 0) && ((ld_t & 3) == 0) && aligned16_dev(s_logits) && aligned16_dev(t_logits);
        const long long v4 = vec ? (V >> 2) : 0;
        for (long long i = tid; i < v4; i += 256) {
            float4 x = ld_stream_f4(ps_ + 4 * i);
            float4 y = ld_stream_f4(pt_ + 4 * i);
            x.x *= inv_T; x.y *= inv_T; x.z *= inv_T; x.w *= inv_T;
            y.x *= inv_T; y.y *= inv_T; y.z *= inv_T; y.w *= inv_T;
            float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
            if (mx > a.m) { a.s *= __expf(a.m - mx); a.m = mx; }
            a.s += __expf(x.x - a.m) + __expf(x.y - a.m) + __expf(x.z - a.m) + __expf(x.w - a.m);
            mx = fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w));
            if (mx > b.m) { b.s *= __expf(b.m - mx); b.m = mx; }
            b.s += __expf(y.x - b.m) + __expf(y.y - b.m) + __expf(y.z - b.m) + __expf(y.w - b.m);
        }
        for (long long i = 4 * v4 + tid; i < V; i += 256) { push(a, __ldg(ps_ + i) * inv_T); push(b, __ldg(pt_ + i) * inv_T); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MS c; c.m = __shfl_xor_sync(0xffffffffu, a.m, o); c.s = __shfl_xor_sync(0xffffffffu, a.s, o);
            a = combine(a, c);
            c.m = __shfl_xor_sync(0xffffffffu, b.m, o); c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
            b = combine(b, c);
        }
        __syncthreads();                                            // previous row's readers of sm / s_lse are done
        if (lane == 0) { sm[0][warp] = a.m; sm[1][warp] = a.s; sm[2][warp] = b.m; sm[3][warp] = b.s; }
        __syncthreads();
        if (tid == 0) {
            MS x; x.m = sm[0][0]; x.s = sm[1][0];
            MS y; y.m = sm[2][0]; y.s = sm[3][0];
            for (int w = 1; w < 8; ++w) {
                MS c; c.m = sm[0][w]; c.s = sm[1][w]; x = combine(x, c);
                c.m = sm[2][w]; c.s = sm[3][w]; y = combine(y, c);
            }
            s_lse[0] = x.m + logf(x.s);
            s_lse[1] = y.m + logf(y.s);
        }
        __syncthreads();
        const float lse_s = s_lse[0], lse_t = s_lse[1];
        float acc = 0.f;
        // pass B: lanes walk consecutive elements (coalesced loads and, for the dense gradient, stores)
        for (long long i = tid; i < V; i += 256) {
            const float ls = fmaf(__ldg(ps_ + i), inv_T, -lse_s);      // log softmax(s / T)
            const float lt = fmaf(__ldg(pt_ + i), inv_T, -lse_t);
            const float pt = __expf(lt);
            acc = fmaf(pt, lt - ls, acc);
            if (gr) gr[i] = g * (__expf(ls) - pt);
        }
        acc = warp_sum(acc);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.f;
            for (int w = 0; w < 8; ++w) tot += s_part[w];
            row_loss[row] = tot;
        }
    }
}


// Softmax loss over the LM-head logits with the gradient emitted AS THE fp16 OPERAND of the LM-head backward GEMM
// (SURVEY section 8 f1 / VERDICT r01 item 7: the float32 [B, T, V] dlogits matrix is never materialised):
//   KIND 0 (distillation, p1/distillation_manager.py:64-80): row_loss = KL(softmax(t/T) || softmax(s/T)),
//           d[v] = softmax(s/T)[v] - softmax(t/T)[v]        (caller scale: T^2/rows * 1/T)
//   KIND 1 (next-token cross-entropy, p1/models_sp.py:441-449): row_loss = logsumexp(s) - s[target],
//           d[v] = softmax(s)[v] - [v == target]            (caller scale: 1 / number of scored rows)
// Output: g16[m, v] = fp16(d[v] * 2^(8-E[m])), row_scale[m] = 2^(E[m]-8) with 2^E an upper bound of max_v |d| (rows that are
// not scored -- the last position of each sequence, ignored targets -- are zero with the smallest scale 2^-108, the
// convention of spq_rowscale_f16), max_scale[0] = max_m row_scale[m].
// One 1024-thread CTA per row and ONE CTA per SM: 148 rows x 2 x 201 KB stay in L2, so of the two passes (online
// max / sum-exp; loss + scaled store, the row scale from an analytic bound on |d|) only the first reads HBM.  Algorithmic bytes: 4 (8 for KIND 0) read +
// 2 written per logit.
template <int KIND, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)            // 64 registers per thread at every CTA size
softmax_loss_grad16_kernel(const float* __restrict__ s_logits, long long ld_s, const float* __restrict__ t_logits, long long ld_t,
                           const long long* __restrict__ targets, long long ignore_index, long long M, long long V, float inv_T,
                           long long seq_len, float* __restrict__ row_loss, float* __restrict__ row_valid,
                           unsigned short* __restrict__ g16, long long ld_g, float* __restrict__ row_scale,
                           float* __restrict__ max_scale, int prefetch_mode) {
    constexpr int NW = NT / 32;
    __shared__ float sm[4][32];
    __shared__ float s_b[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float cta_max_scale = 0.f;
    for (long long row = blockIdx.x; row < M; row += gridDim.x) {
        unsigned short* go = g16 + row * ld_g;
        long long tgt = -1;
        bool skip = seq_len > 0 && (row % seq_len) == seq_len - 1;
        if (KIND == 1) {
            tgt = targets[row];
            skip = skip || tgt == ignore_index || tgt < 0 || tgt >= V;
        }
        if (skip) {                                                  // block-uniform
            if (tid == 0) { row_loss[row] = 0.f; if (row_valid) row_valid[row] = 0.f; row_scale[row] = 3.0814879110195774e-33f; }   // 2^-108
            for (long long i = tid; i < (ld_g >> 1); i += NT) reinterpret_cast<unsigned int*>(go)[i] = 0u;
            cta_max_scale = fmaxf(cta_max_scale, 3.0814879110195774e-33f);
            continue;
        }
        const float* ps_ = s_logits + row * ld_s;
        const float* pt_ = KIND == 0 ? t_logits + row * ld_t : nullptr;
        const bool vec = ((ld_s & 3) == 0) && aligned16_dev(s_logits) && (KIND == 1 || (((ld_t & 3) == 0) && aligned16_dev(t_logits)));
        const long long v4 = vec ? (V >> 2) : 0;
        // ---- pass 1 (HBM): online max / sum exp
        MS a; a.m = -INFINITY; a.s = 0.f;
        MS b; b.m = -INFINITY; b.s = 0.f;
        auto fold4 = [&](MS& acc, float4 x) {
            x.x *= inv_T; x.y *= inv_T; x.z *= inv_T; x.w *= inv_T;
            const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
            if (mx > acc.m) { acc.s *= fexp(acc.m - mx); acc.m = mx; }
            acc.s += fexp(x.x - acc.m) + fexp(x.y - acc.m) + fexp(x.z - acc.m) + fexp(x.w - acc.m);
        };
        {
            // four 16-byte loads per matrix in flight per thread (the pass streams from HBM)
            long long i = tid;
            for (; i + 3 * NT < v4; i += 4 * NT) {
                float4 xs[4], ys[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xs[u] = ld_stream_f4(ps_ + 4 * (i + u * NT));
                    if (KIND == 0) ys[u] = ld_stream_f4(pt_ + 4 * (i + u * NT));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    fold4(a, xs[u]);
                    if (KIND == 0) fold4(b, ys[u]);
                }
            }
            for (; i < v4; i += NT) {
                fold4(a, ld_stream_f4(ps_ + 4 * i));
                if (KIND == 0) fold4(b, ld_stream_f4(pt_ + 4 * i));
            }
        }
        for (long long i = 4 * v4 + tid; i < V; i += NT) {
            push(a, __ldg(ps_ + i) * inv_T);
            if (KIND == 0) push(b, __ldg(pt_ + i) * inv_T);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MS c; c.m = __shfl_xor_sync(0xffffffffu, a.m, o); c.s = __shfl_xor_sync(0xffffffffu, a.s, o);
            a = combine(a, c);
            if (KIND == 0) {
                c.m = __shfl_xor_sync(0xffffffffu, b.m, o); c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
                b = combine(b, c);
            }
        }
        __syncthreads();                                             // previous row's readers of sm / s_b are done
        if (lane == 0) { sm[0][warp] = a.m; sm[1][warp] = a.s; sm[2][warp] = b.m; sm[3][warp] = b.s; }
        __syncthreads();
        if (warp == 0) {
            MS x; x.m = lane < NW ? sm[0][lane] : -INFINITY; x.s = lane < NW ? sm[1][lane] : 0.f;
            MS y; y.m = lane < NW ? sm[2][lane] : -INFINITY; y.s = lane < NW ? sm[3][lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                MS c; c.m = __shfl_xor_sync(0xffffffffu, x.m, o); c.s = __shfl_xor_sync(0xffffffffu, x.s, o);
                x = combine(x, c);
                c.m = __shfl_xor_sync(0xffffffffu, y.m, o); c.s = __shfl_xor_sync(0xffffffffu, y.s, o);
                y = combine(y, c);
            }
            if (lane == 0) { s_b[0] = x.m + logf(x.s); s_b[1] = (KIND == 0) ? y.m + logf(y.s) : 0.f; s_b[2] = x.m; s_b[3] = y.m; }
        }
        __syncthreads();
        const float lse_s = s_b[0], lse_t = s_b[1];
        // while passes 2 and 3 run out of L2, pull the NEXT row of this CTA towards L2 (its pass 1 then starts warm)
        {
            const long long nrow = row + gridDim.x;
            if (nrow < M && prefetch_mode != 0) {
                const char* ns = reinterpret_cast<const char*>(s_logits + nrow * ld_s);
                const long long bytes = V * 4;
                for (long long o = static_cast<long long>(tid) * 128; o < bytes; o += static_cast<long long>(NT) * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ns + o));
                    if (KIND == 0 && prefetch_mode == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(t_logits + nrow * ld_t) + o));
                }
            }
        }
        // The row scale comes from an analytic bound instead of a pass over d: |d[v]| <= max(max_v p_s, max_v p_t) (KL) resp.
        // max(1 - p_s[target], max_v p_s) (CE), and the row maxima are known from pass 1.  The bound is within a factor 2 of
        // max |d| for CE; for KL it can be far above it when student and teacher agree, which only moves the operand down
        // inside fp16's normal range (elements more than 2^-22 below the bound are negligible either way).
        float amax;
        if (KIND == 0) amax = fmaxf(fexp(s_b[2] - lse_s), fexp(s_b[3] - lse_t));
        else amax = fmaxf(1.0f - fexp(fmaf(__ldg(ps_ + tgt), inv_T, -lse_s)), fexp(s_b[2] - lse_s));
        int E = 0;
        if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &E); else E = (amax == 0.f) ? -100 : 8;
        E = E < -100 ? -100 : E;
        const float down = exp2f(static_cast<float>(8 - E));
        const float up = exp2f(static_cast<float>(E - 8));
        cta_max_scale = fmaxf(cta_max_scale, up);
        // ---- pass 2 (L2): loss terms and the scaled fp16 operand
        float acc = 0.f;
        auto dq = [&](float sv, float tv, long long idx) -> float {
            const float ls = fmaf(sv, inv_T, -lse_s);
            float d;
            if (KIND == 0) {
                const float lt = fmaf(tv, inv_T, -lse_t);
                const float pt = fexp(lt);
                acc = fmaf(pt, lt - ls, acc);
                d = fexp(ls) - pt;
            } else {
                d = fexp(ls) - (idx == tgt ? 1.0f : 0.0f);
            }
            return d * down;
        };
        const bool gvec = (ld_g & 3) == 0 && (reinterpret_cast<uintptr_t>(g16) & 7u) == 0;
        if (gvec) {
            auto put = [&](long long i, const float4& x, const float4& y) {
                *reinterpret_cast<uint2*>(go + 4 * i) = make_uint2(pack_h2(dq(x.x, y.x, 4 * i), dq(x.y, y.y, 4 * i + 1)),
                                                                  pack_h2(dq(x.z, y.z, 4 * i + 2), dq(x.w, y.w, 4 * i + 3)));
            };
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            long long i = tid;
            for (; i + 3 * NT < v4; i += 4 * NT) {
                float4 xs[4], ys[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xs[u] = *reinterpret_cast<const float4*>(ps_ + 4 * (i + u * NT));
                    ys[u] = (KIND == 0) ? *reinterpret_cast<const float4*>(pt_ + 4 * (i + u * NT)) : z4;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) put(i + u * NT, xs[u], ys[u]);
            }
            for (; i < v4; i += NT)
                put(i, *reinterpret_cast<const float4*>(ps_ + 4 * i), (KIND == 0) ? *reinterpret_cast<const float4*>(pt_ + 4 * i) : z4);
            for (long long j = 4 * v4 + tid; j < ld_g; j += NT)
                go[j] = j < V ? f2h_sat(dq(__ldg(ps_ + j), KIND == 0 ? __ldg(pt_ + j) : 0.f, j)) : static_cast<unsigned short>(0);
        } else {
            for (long long j = tid; j < ld_g; j += NT)
                go[j] = j < V ? f2h_sat(dq(__ldg(ps_ + j), KIND == 0 ? __ldg(pt_ + j) : 0.f, j)) : static_cast<unsigned short>(0);
        }
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) sm[0][warp] = acc;
        __syncthreads();
        if (warp == 0) {
            const float t0 = warp_sum(lane < NW ? sm[0][lane] : 0.f);
            if (lane == 0) {
                row_loss[row] = (KIND == 0) ? t0 : (lse_s - __ldg(ps_ + tgt) * inv_T);
                if (row_valid) row_valid[row] = 1.f;
                row_scale[row] = up;
            }
        }
    }
    if (tid == 0 && max_scale) atomicMax(reinterpret_cast<int*>(max_scale), __float_as_int(cta_max_scale));
}

// Feature term of the distillation loss (p1/distillation_manager.py:82-116): mean squared error between ONE pair of
// hidden-state tensors, the layer drawn at random per micro-step.  A captured micro-step cannot know the layer, so the
// training driver used to evaluate all 13 pairs with torch (26 launches, 650 MB of reads per student micro-step) and pick
// one on the host; here the candidate pairs are kernel arguments and the drawn index is read from device memory.
constexpr int MSE_MAX_PAIRS = 32;
struct MsePairs {
    const float* a[MSE_MAX_PAIRS];
    const float* b[MSE_MAX_PAIRS];
};

__global__ void __launch_bounds__(256)
mse_select_partial_kernel(MsePairs p, int n_pairs, const int32_t* __restrict__ select, long long numel, float* __restrict__ partial) {
    int sel = *select;
    sel = sel < 0 ? 0 : (sel >= n_pairs ? n_pairs - 1 : sel);
    const float* a = p.a[sel];
    const float* b = p.b[sel];
    const long long n4 = numel >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    float acc0 = 0.f, acc1 = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = ld_stream_f4(a + 4 * i), y = ld_stream_f4(b + 4 * i);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        acc0 += d0 * d0 + d1 * d1;
        acc1 += d2 * d2 + d3 * d3;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = n4 << 2; i < numel; ++i) { const float d = a[i] - b[i]; acc0 += d * d; }
    }
    float acc = warp_sum(acc0 + acc1);
    __shared__ float s_red[8];
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        partial[blockIdx.x] = t;
    }
}

// fixed-order fold: the result is reproducible run to run
__global__ void __launch_bounds__(32) mse_select_final_kernel(const float* __restrict__ partial, int n, long long numel, float* __restrict__ out) {
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t / static_cast<float>(numel);
}

}  // namespace loss
}  // namespace spq

using namespace spq;

extern "C" int spq_cross_entropy_fwd(const float* logits, int64_t M, int64_t V, int64_t ld, const int64_t* targets,
                                     int64_t ignore_index, float* row_loss, float* row_valid, spq_stream_t stream) {
    SPQ_REQUIRE(logits && targets && row_loss && row_valid && M > 0 && V > 0 && ld >= V, "spq_cross_entropy_fwd: bad arguments");
    long long ctas = static_cast<long long>(sm_count()) * 8;
    if (ctas > M) ctas = M;
    loss::cross_entropy_fwd_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(
        logits, M, V, ld, reinterpret_cast<const long long*>(targets), ignore_index, row_loss, row_valid);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_cross_entropy_from_parts(const float* parts, int64_t P, int64_t part_ld, const float* logits, int64_t M,
                                            int64_t V, int64_t ld, const int64_t* targets, int64_t ignore_index,
                                            float* row_loss, float* row_valid, spq_stream_t stream) {
    SPQ_REQUIRE(parts && logits && targets && row_loss && row_valid && M > 0 && V > 0 && ld >= V && P > 0 && part_ld >= P,
                "spq_cross_entropy_from_parts: bad arguments");
    long long ctas = (M + 7) / 8;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    if (ctas > cap) ctas = cap;
    loss::cross_entropy_from_parts_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(parts), P, part_ld, logits, M, V, ld, reinterpret_cast<const long long*>(targets),
        ignore_index, row_loss, row_valid);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_distill_kl(const float* s_logits, int64_t ld_s, const float* t_logits, int64_t ld_t, int64_t M, int64_t V,
                              float temperature, int64_t seq_len, float grad_scale, float* row_loss, float* grad,
                              spq_stream_t stream) {
    SPQ_REQUIRE(s_logits && t_logits && row_loss && M > 0 && V > 0 && ld_s >= V && ld_t >= V && temperature > 0.f,
                "spq_distill_kl: bad arguments");
    long long ctas = static_cast<long long>(sm_count()) * 8;
    if (ctas > M) ctas = M;
    loss::distill_kl_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(
        s_logits, ld_s, t_logits, ld_t, M, V, 1.0f / temperature, seq_len, grad_scale / temperature, row_loss, grad);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_softmax_loss_grad16(int kind, const float* s_logits, int64_t ld_s, const float* t_logits, int64_t ld_t,
                                       const int64_t* targets, int64_t ignore_index, int64_t M, int64_t V, float temperature,
                                       int64_t seq_len, float* row_loss, float* row_valid, spq_half_t* g16, int64_t ld_g,
                                       float* row_scale, float* max_scale, spq_stream_t stream) {
    SPQ_REQUIRE(s_logits && row_loss && g16 && row_scale && max_scale && M > 0 && V > 0 && ld_s >= V && ld_g >= V && (ld_g % 2) == 0,
                "spq_softmax_loss_grad16: bad arguments");
    SPQ_REQUIRE((reinterpret_cast<uintptr_t>(g16) & 3u) == 0, "spq_softmax_loss_grad16: g16 must be 4-byte aligned");
    SPQ_REQUIRE(kind == 0 || kind == 1, "spq_softmax_loss_grad16: kind must be 0 (distillation KL) or 1 (cross-entropy)");
    SPQ_REQUIRE(kind == 1 || (t_logits && ld_t >= V && temperature > 0.f), "spq_softmax_loss_grad16: KL needs teacher logits and T > 0");
    SPQ_REQUIRE(kind == 0 || targets, "spq_softmax_loss_grad16: cross-entropy needs targets");
    cudaStream_t st = as_stream(stream);
    SPQ_CUDA_OK(cudaMemsetAsync(max_scale, 0, sizeof(float), st));
    // CTA size / rows in flight / next-row L2 prefetch.  ncu of the first layout (one 1024-thread CTA per SM, prefetch of
    // both matrices: profiles/r02c_softmax_loss_ncu_before.txt) showed DRAM at 5.3 TB/s for 2.4 TB/s of algorithmic
    // traffic and an L2 hit rate of 16 %: 148 rows x 2 matrices x 201 KB plus the prefetched next rows do not fit the
    // 126 MB L2, so pass 2 re-read everything from DRAM.  Measured at 8192 x 50257 (KL / CE, us): 1024 threads + prefetch
    // 1722 / 921; 1024 threads, none 1433 / 919; 512 threads (2 CTAs per SM) + prefetch 1485 / 783; 512 threads, none
    // 1085 / 671 <- default.  SPQ_LOSS_THREADS = 256 | 512 | 1024 and SPQ_LOSS_PREFETCH = 0 none | 1 every matrix |
    // 2 student matrix only are the A/B switches.
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("SPQ_LOSS_THREADS"); variant = e ? atoi(e) : 512; }
    const int nt = variant == 1024 ? 1024 : variant == 256 ? 256 : 512;
    static int pf_env = -2;
    if (pf_env == -2) { const char* e = getenv("SPQ_LOSS_PREFETCH"); pf_env = e ? atoi(e) : -1; }
    const int pf = pf_env >= 0 ? pf_env : 0;
    long long ctas = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * (1024 / nt);     // 1024 threads per SM
    if (ctas > M) ctas = M;
    const unsigned grid = static_cast<unsigned>(ctas);
    const long long* tg = reinterpret_cast<const long long*>(targets);
#define SPQ_SOFTMAX_LAUNCH(KIND, NT, T_PTR, LD_T, TG, INV_T)                                                                  \
    loss::softmax_loss_grad16_kernel<KIND, NT><<<grid, NT, 0, st>>>(s_logits, ld_s, T_PTR, LD_T, TG, ignore_index, M, V, INV_T, \
                                                                     seq_len, row_loss, row_valid, g16, ld_g, row_scale, max_scale, pf)
    if (kind == 0) {
        if (nt == 1024) SPQ_SOFTMAX_LAUNCH(0, 1024, t_logits, ld_t, nullptr, 1.0f / temperature);
        else if (nt == 512) SPQ_SOFTMAX_LAUNCH(0, 512, t_logits, ld_t, nullptr, 1.0f / temperature);
        else SPQ_SOFTMAX_LAUNCH(0, 256, t_logits, ld_t, nullptr, 1.0f / temperature);
    } else {
        if (nt == 1024) SPQ_SOFTMAX_LAUNCH(1, 1024, nullptr, 0, tg, 1.0f);
        else if (nt == 512) SPQ_SOFTMAX_LAUNCH(1, 512, nullptr, 0, tg, 1.0f);
        else SPQ_SOFTMAX_LAUNCH(1, 256, nullptr, 0, tg, 1.0f);
    }
#undef SPQ_SOFTMAX_LAUNCH
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" size_t spq_mse_select_workspace_bytes(void) {
    return static_cast<size_t>((sm_count() > 0 ? sm_count() : 148) * 4) * sizeof(float);
}

extern "C" int spq_mse_select(const float* const* a, const float* const* b, int n_pairs, const int32_t* select, int64_t numel,
                              float* out, void* workspace, size_t workspace_bytes, spq_stream_t stream) {
    SPQ_REQUIRE(a && b && select && out && workspace && n_pairs > 0 && n_pairs <= loss::MSE_MAX_PAIRS && numel > 0,
                "spq_mse_select: bad arguments (at most %d pairs)", loss::MSE_MAX_PAIRS);
    SPQ_REQUIRE(workspace_bytes >= spq_mse_select_workspace_bytes(), "spq_mse_select: workspace too small");
    loss::MsePairs p = {};
    for (int i = 0; i < n_pairs; ++i) {
        SPQ_REQUIRE(a[i] && b[i] && aligned16(a[i]) && aligned16(b[i]), "spq_mse_select: pair %d null or not 16-byte aligned", i);
        p.a[i] = a[i];
        p.b[i] = b[i];
    }
    long long blocks = (numel / 4 + 255) / 256;
    const long long cap = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = as_stream(stream);
    loss::mse_select_partial_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(p, n_pairs, select, numel, static_cast<float*>(workspace));
    SPQ_LAUNCH_OK();
    loss::mse_select_final_kernel<<<1, 32, 0, st>>>(static_cast<const float*>(workspace), static_cast<int>(blocks), numel, out);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

