// Flat-buffer optimiser step of the switchable-precision training step (p1/train_sp.py:390-393
// `unscale_ / clip_grad_norm_(1.0) / AdamW.step`): the trainable parameters of the path (LoRA A/B, LayerNorm pairs)
// live in ONE flat fp32 buffer whose slices are the module parameters, their gradients in a second one (the slices
// are the .grad tensors, so the data-parallel all-reduce needs no pack / scatter kernels), and the step is
//   (1) sum of squares of the whole gradient buffer  -> device scalar   (two launches, deterministic tree)
//   (2) per touched segment: clip coefficient from that scalar, AdamW update, all in one pass.
// HBM-bound: (1) reads 4 B / element, (2) reads 16 B and writes 12 B / element.
#include "spq_common.cuh"

namespace spq {
namespace optim {

constexpr int THREADS = 256;
constexpr int MAX_PARTS = 1024;

__global__ void __launch_bounds__(THREADS)
sumsq_partial_kernel(const float* __restrict__ g, long long n4, long long n, float* __restrict__ parts) {
    float acc = 0.f;
    const long long stride = static_cast<long long>(gridDim.x) * THREADS;
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < n4; i += stride) {
        const float4 v = ld_stream_f4(g + 4 * i);
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    if (blockIdx.x == 0)
        for (long long i = 4 * n4 + threadIdx.x; i < n; i += THREADS) acc = fmaf(g[i], g[i], acc);
    __shared__ float s[THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < THREADS / 32 ? s[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) parts[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(THREADS)
sumsq_final_kernel(const float* __restrict__ parts, int nparts, float scale_sq, float* __restrict__ out) {
    // fixed-order tree: the result does not depend on scheduling
    float acc = 0.f;
    for (int i = threadIdx.x; i < nparts; i += THREADS) acc += parts[i];
    __shared__ float s[THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < THREADS / 32 ? s[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * scale_sq;
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_corr1, inv_sqrt_bias_corr2, grad_scale, max_norm;
};

__global__ void __launch_bounds__(THREADS)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n4, const float* __restrict__ total_sumsq, AdamArgs a) {
    // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6)); max_norm <= 0 disables clipping
    float coef = a.grad_scale;
    if (a.max_norm > 0.f && total_sumsq) {
        const float norm = sqrtf(__ldg(total_sumsq));
        coef *= fminf(1.0f, a.max_norm / (norm + 1e-6f));
    }
    const float step = a.lr / a.bias_corr1;
    const float decay = 1.0f - a.lr * a.weight_decay;
    const long long stride = static_cast<long long>(gridDim.x) * THREADS;
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < n4; i += stride) {
        float4 pv = *reinterpret_cast<const float4*>(p + 4 * i);
        const float4 gv = ld_stream_f4(g + 4 * i);
        float4 mv = *reinterpret_cast<const float4*>(m + 4 * i);
        float4 vv = *reinterpret_cast<const float4*>(v + 4 * i);
        auto one = [&](float& pp, float gg, float& mm, float& v2) {
            gg *= coef;
            pp *= decay;                                             // decoupled weight decay (AdamW)
            mm = fmaf(a.beta1, mm, (1.0f - a.beta1) * gg);           // lerp(m, g, 1 - beta1)
            v2 = fmaf(a.beta2, v2, (1.0f - a.beta2) * gg * gg);
            const float denom = fmaf(sqrtf(v2), a.inv_sqrt_bias_corr2, a.eps);
            pp -= step * __fdiv_rn(mm, denom);
        };
        one(pv.x, gv.x, mv.x, vv.x); one(pv.y, gv.y, mv.y, vv.y); one(pv.z, gv.z, mv.z, vv.z); one(pv.w, gv.w, mv.w, vv.w);
        *reinterpret_cast<float4*>(p + 4 * i) = pv;
        *reinterpret_cast<float4*>(m + 4 * i) = mv;
        *reinterpret_cast<float4*>(v + 4 * i) = vv;
    }
}

}  // namespace optim
}  // namespace spq

using namespace spq;
using namespace spq::optim;

extern "C" size_t spq_sumsq_workspace_bytes(void) { return MAX_PARTS * sizeof(float); }

extern "C" int spq_grad_sumsq(const float* g, int64_t n, float scale, float* out_sumsq, void* workspace, size_t workspace_bytes,
                              spq_stream_t stream) {
    SPQ_REQUIRE(g && out_sumsq && workspace && n > 0, "spq_grad_sumsq: bad arguments");
    SPQ_REQUIRE(workspace_bytes >= spq_sumsq_workspace_bytes() && aligned16(g), "spq_grad_sumsq: workspace too small or unaligned buffer");
    cudaStream_t st = as_stream(stream);
    const long long n4 = n / 4;
    long long blocks = (n4 + THREADS - 1) / THREADS;
    const long long cap = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks > MAX_PARTS) blocks = MAX_PARTS;
    if (blocks < 1) blocks = 1;
    float* parts = reinterpret_cast<float*>(workspace);
    sumsq_partial_kernel<<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(g, n4, n, parts);
    SPQ_LAUNCH_OK();
    sumsq_final_kernel<<<1, THREADS, 0, st>>>(parts, static_cast<int>(blocks), scale * scale, out_sumsq);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                              const float* total_sumsq, float max_norm, spq_stream_t stream) {
    SPQ_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "spq_adamw_flat: bad arguments");
    SPQ_REQUIRE((n % 4) == 0 && aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq),
                "spq_adamw_flat: segments must be 16-byte aligned and a multiple of 4 elements");
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bias_corr1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
    a.inv_sqrt_bias_corr2 = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
    a.grad_scale = grad_scale; a.max_norm = max_norm;
    const long long n4 = n / 4;
    long long blocks = (n4 + THREADS - 1) / THREADS;
    const long long cap = static_cast<long long>(sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    adamw_flat_kernel<<<static_cast<unsigned>(blocks), THREADS, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n4,
                                                                                       total_sumsq, a);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}
