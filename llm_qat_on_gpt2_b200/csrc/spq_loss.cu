// Next-token cross-entropy forward over a [M, V] logits matrix (row stride `ld`), one pass.
//
// Not part of the quantised-linear hot path proper: this is the first consumer after it (SURVEY.md
// section 8 f1).  The reference computes the loss with nn.CrossEntropyLoss on shifted, re-materialised
// copies of the [B, T, 50257] logits (p1/models_sp.py:441-449): ~5 full passes over 6.6 GB at
// 32 x 1024 tokens.  Here each row is read once: online max / sum-of-exponentials per thread,
// block combine, loss_row = log(sum) + max - logit[target].  Forward only (no-grad evaluation /
// calibration); training keeps torch's differentiable implementation.
#include "spq_common.cuh"

namespace spq {
namespace loss {

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
