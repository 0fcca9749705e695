// SwitchableLayerNorm forward / backward (HBM-bound, row-resident).
//
// Replaces SwitchableLayerNorm.forward of the reference (p1/switchable_batchnorm.py:102-109:
// mean, biased var, sub, add, sqrt, div, mul, add = 8 eager kernels) and its autograd backward.
// The (weight, bias) pair of the active precision is selected on the host; the kernel sees plain
// pointers.  One CTA of G threads owns a row at a time, the row lives in registers: forward is
// 4 B read + 4 B written per element, backward 8 B read + 4 B written, and the dweight / dbias
// column sums are accumulated per CTA and folded by a second small kernel.
#include "spq_common.cuh"

namespace spq {
namespace ln {

template <int NV>
__device__ __forceinline__ float block_sum(float v, float* s_red, int G, int tid) {
    v = warp_sum(v);
    if (G > 32) {
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = v;
        __syncthreads();
        v = s_red[0];
        for (int w = 1; w < (G >> 5); ++w) v += s_red[w];
    }
    return v;
}

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, long long rows, long long C, const float* __restrict__ w,
                     const float* __restrict__ b, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
    const int G = blockDim.x, tid = threadIdx.x;
    __shared__ float s_red[8];
    float4 wv[NV], bv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long long c = (static_cast<long long>(i) * G + tid) * 4;
        wv[i] = (c < C) ? *reinterpret_cast<const float4*>(w + c) : make_float4(0, 0, 0, 0);
        bv[i] = (c < C) ? *reinterpret_cast<const float4*>(b + c) : make_float4(0, 0, 0, 0);
    }
    const float inv_c = 1.0f / static_cast<float>(C);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        float4 v[NV];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            v[i] = (c < C) ? ld_stream_f4(x + row * C + c) : make_float4(0, 0, 0, 0);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = block_sum<NV>(sum, s_red, G, tid) * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            if (c < C) {
                const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
        const float var = block_sum<NV>(sq, s_red, G, tid) * inv_c;
        const float rstd = 1.0f / sqrtf(var + eps);
        if (tid == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            if (c < C) {
                float4 o;
                o.x = wv[i].x * ((v[i].x - mean) * rstd) + bv[i].x;
                o.y = wv[i].y * ((v[i].y - mean) * rstd) + bv[i].y;
                o.z = wv[i].z * ((v[i].z - mean) * rstd) + bv[i].z;
                o.w = wv[i].w * ((v[i].w - mean) * rstd) + bv[i].w;
                *reinterpret_cast<float4*>(y + row * C + c) = o;
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long rows, long long C,
                     float* __restrict__ dx, float* __restrict__ part_dw, float* __restrict__ part_db) {
    const int G = blockDim.x, tid = threadIdx.x;
    __shared__ float s_red[8];
    float4 wv[NV], aw[NV], ab[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long long c = (static_cast<long long>(i) * G + tid) * 4;
        wv[i] = (c < C) ? *reinterpret_cast<const float4*>(w + c) : make_float4(0, 0, 0, 0);
        aw[i] = make_float4(0, 0, 0, 0);
        ab[i] = make_float4(0, 0, 0, 0);
    }
    const float inv_c = 1.0f / static_cast<float>(C);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
        float4 xh[NV], g[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            if (c < C) {
                const float4 xv = ld_stream_f4(x + row * C + c);
                const float4 gv = ld_stream_f4(dy + row * C + c);
                xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                aw[i].x += gv.x * xh[i].x; aw[i].y += gv.y * xh[i].y; aw[i].z += gv.z * xh[i].z; aw[i].w += gv.w * xh[i].w;
                ab[i].x += gv.x; ab[i].y += gv.y; ab[i].z += gv.z; ab[i].w += gv.w;
                g[i] = make_float4(gv.x * wv[i].x, gv.y * wv[i].y, gv.z * wv[i].z, gv.w * wv[i].w);
                s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            } else {
                xh[i] = make_float4(0, 0, 0, 0);
                g[i] = make_float4(0, 0, 0, 0);
            }
        }
        const float m1 = block_sum<NV>(s1, s_red, G, tid) * inv_c;
        const float m2 = block_sum<NV>(s2, s_red, G, tid) * inv_c;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const long long c = (static_cast<long long>(i) * G + tid) * 4;
            if (c < C) {
                float4 o;
                o.x = (g[i].x - m1 - xh[i].x * m2) * rstd;
                o.y = (g[i].y - m1 - xh[i].y * m2) * rstd;
                o.z = (g[i].z - m1 - xh[i].z * m2) * rstd;
                o.w = (g[i].w - m1 - xh[i].w * m2) * rstd;
                *reinterpret_cast<float4*>(dx + row * C + c) = o;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long long c = (static_cast<long long>(i) * G + tid) * 4;
        if (c < C) {
            *reinterpret_cast<float4*>(part_dw + static_cast<long long>(blockIdx.x) * C + c) = aw[i];
            *reinterpret_cast<float4*>(part_db + static_cast<long long>(blockIdx.x) * C + c) = ab[i];
        }
    }
}

// Fold the per-CTA partial column sums.  Block = 8 columns x 32 part-lanes: a warp row reads 8 consecutive floats
// (one 32-byte sector) of 4 different partial rows, every thread keeps 4 independent loads in flight, and the grid
// is C / 8 blocks (96 for C = 768; the first version's 24 blocks of 32 columns took as long as the backward itself).
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const float* __restrict__ part_a, const float* __restrict__ part_b, int parts, long long C,
                       float* __restrict__ out_a, float* __restrict__ out_b, int accumulate) {
    __shared__ float fa[32][9], fb[32][9];
    const int cl = threadIdx.x & 7, sl = threadIdx.x >> 3;            // column in the block, part lane 0..31
    const long long c = static_cast<long long>(blockIdx.x) * 8 + cl;
    float a = 0.f, b = 0.f;
    if (c < C) {
        int p = sl;
        for (; p + 96 < parts; p += 128) {
            const float a0 = part_a[static_cast<long long>(p) * C + c], a1 = part_a[static_cast<long long>(p + 32) * C + c];
            const float a2 = part_a[static_cast<long long>(p + 64) * C + c], a3 = part_a[static_cast<long long>(p + 96) * C + c];
            const float b0 = part_b[static_cast<long long>(p) * C + c], b1 = part_b[static_cast<long long>(p + 32) * C + c];
            const float b2 = part_b[static_cast<long long>(p + 64) * C + c], b3 = part_b[static_cast<long long>(p + 96) * C + c];
            a += (a0 + a1) + (a2 + a3);
            b += (b0 + b1) + (b2 + b3);
        }
        for (; p < parts; p += 32) {
            a += part_a[static_cast<long long>(p) * C + c];
            b += part_b[static_cast<long long>(p) * C + c];
        }
    }
    fa[sl][cl] = a; fb[sl][cl] = b;
    __syncthreads();
    if (sl == 0 && c < C) {
#pragma unroll
        for (int w = 1; w < 32; ++w) { a += fa[w][cl]; b += fb[w][cl]; }
        if (out_a) out_a[c] = accumulate ? out_a[c] + a : a;
        if (out_b) out_b[c] = accumulate ? out_b[c] + b : b;
    }
}

struct Cfg { int G, NV; unsigned grid; };

static bool pick(long long rows, long long C, Cfg* cfg, int ctas_per_sm_cap) {
    const long long nvec = (C + 3) / 4;
    if (nvec <= 32) { cfg->G = 32; cfg->NV = 1; }
    else if (nvec <= 64) { cfg->G = 32; cfg->NV = 2; }
    else if (nvec <= 128) { cfg->G = 32; cfg->NV = 4; }
    else if (nvec <= 256) { cfg->G = 64; cfg->NV = 4; }
    else if (nvec <= 512) { cfg->G = 128; cfg->NV = 4; }
    else if (nvec <= 1024) { cfg->G = 256; cfg->NV = 4; }
    else if (nvec <= 2048) { cfg->G = 256; cfg->NV = 8; }
    else return false;
    int per_sm = 2048 / cfg->G;
    if (per_sm > ctas_per_sm_cap) per_sm = ctas_per_sm_cap;
    long long ctas = static_cast<long long>(sm_count()) * per_sm;
    if (ctas > rows) ctas = rows;
    if (ctas < 1) ctas = 1;
    cfg->grid = static_cast<unsigned>(ctas);
    return true;
}

}  // namespace ln
}  // namespace spq

using namespace spq;
using namespace spq::ln;

extern "C" int spq_layernorm_fwd(const float* x, int64_t rows, int64_t cols, const float* weight, const float* bias, float eps,
                                 float* y, float* mean, float* rstd, spq_stream_t stream) {
    SPQ_REQUIRE(x && weight && bias && y && rows > 0 && cols > 0, "spq_layernorm_fwd: bad arguments");
    SPQ_REQUIRE((cols % 4) == 0 && aligned16(x) && aligned16(y) && aligned16(weight) && aligned16(bias),
                "spq_layernorm_fwd: normalized dim must be a multiple of 4 and pointers 16-byte aligned");
    Cfg c;
    if (!pick(rows, cols, &c, 16)) {
        set_error("spq_layernorm_fwd: normalized dim %lld > 8192 unsupported", (long long)cols);
        return SPQ_ERR_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    switch (c.NV) {
        case 1: layernorm_fwd_kernel<1><<<c.grid, c.G, 0, st>>>(x, rows, cols, weight, bias, eps, y, mean, rstd); break;
        case 2: layernorm_fwd_kernel<2><<<c.grid, c.G, 0, st>>>(x, rows, cols, weight, bias, eps, y, mean, rstd); break;
        case 4: layernorm_fwd_kernel<4><<<c.grid, c.G, 0, st>>>(x, rows, cols, weight, bias, eps, y, mean, rstd); break;
        default: layernorm_fwd_kernel<8><<<c.grid, c.G, 0, st>>>(x, rows, cols, weight, bias, eps, y, mean, rstd); break;
    }
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" size_t spq_layernorm_bwd_workspace_bytes(int64_t rows, int64_t cols) {
    Cfg c;
    if (rows <= 0 || cols <= 0 || !pick(rows, cols, &c, 8)) return 0;
    return static_cast<size_t>(c.grid) * static_cast<size_t>(cols) * 2 * sizeof(float);
}

extern "C" int spq_layernorm_bwd(const float* dy, const float* x, const float* weight, const float* mean, const float* rstd,
                                 int64_t rows, int64_t cols, float* dx, float* dweight, float* dbias, int accumulate_params,
                                 void* workspace, size_t workspace_bytes, spq_stream_t stream) {
    SPQ_REQUIRE(dy && x && weight && mean && rstd && dx && workspace && rows > 0 && cols > 0, "spq_layernorm_bwd: bad arguments");
    SPQ_REQUIRE((cols % 4) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(weight) && aligned16(workspace),
                "spq_layernorm_bwd: alignment");
    Cfg c;
    if (!pick(rows, cols, &c, 8)) {          // 8 CTAs of G threads per SM: ~64 KB of loads in flight per SM
        set_error("spq_layernorm_bwd: normalized dim %lld > 8192 unsupported", (long long)cols);
        return SPQ_ERR_UNSUPPORTED;
    }
    SPQ_REQUIRE(workspace_bytes >= spq_layernorm_bwd_workspace_bytes(rows, cols), "spq_layernorm_bwd: workspace too small");
    float* pdw = reinterpret_cast<float*>(workspace);
    float* pdb = pdw + static_cast<size_t>(c.grid) * cols;
    cudaStream_t st = as_stream(stream);
    switch (c.NV) {
        case 1: layernorm_bwd_kernel<1><<<c.grid, c.G, 0, st>>>(dy, x, weight, mean, rstd, rows, cols, dx, pdw, pdb); break;
        case 2: layernorm_bwd_kernel<2><<<c.grid, c.G, 0, st>>>(dy, x, weight, mean, rstd, rows, cols, dx, pdw, pdb); break;
        case 4: layernorm_bwd_kernel<4><<<c.grid, c.G, 0, st>>>(dy, x, weight, mean, rstd, rows, cols, dx, pdw, pdb); break;
        default: layernorm_bwd_kernel<8><<<c.grid, c.G, 0, st>>>(dy, x, weight, mean, rstd, rows, cols, dx, pdw, pdb); break;
    }
    SPQ_LAUNCH_OK();
    if (dweight || dbias) {
        colsum_finalize_kernel<<<static_cast<unsigned>((cols + 7) / 8), 256, 0, st>>>(pdw, pdb, static_cast<int>(c.grid), cols,
                                                                                            dweight, dbias, accumulate_params ? 1 : 0);
        SPQ_LAUNCH_OK();
    }
    return SPQ_OK;
}

// The second half of spq_layernorm_bwd on its own: fold the per-CTA column sums a call with dweight = dbias = null left
// in `workspace` into dweight / dbias.  Lets a training driver run the fold (a latency-bound launch nothing downstream
// waits for) on a side stream; the workspace must then be private to that backward call.
extern "C" int spq_layernorm_bwd_finalize(const void* workspace, size_t workspace_bytes, int64_t rows, int64_t cols, float* dweight,
                                          float* dbias, int accumulate_params, spq_stream_t stream) {
    SPQ_REQUIRE(workspace && rows > 0 && cols > 0 && (dweight || dbias), "spq_layernorm_bwd_finalize: bad arguments");
    Cfg c;
    if (!pick(rows, cols, &c, 8)) {
        set_error("spq_layernorm_bwd_finalize: normalized dim %lld > 8192 unsupported", (long long)cols);
        return SPQ_ERR_UNSUPPORTED;
    }
    SPQ_REQUIRE(workspace_bytes >= spq_layernorm_bwd_workspace_bytes(rows, cols), "spq_layernorm_bwd_finalize: workspace too small");
    const float* pdw = reinterpret_cast<const float*>(workspace);
    const float* pdb = pdw + static_cast<size_t>(c.grid) * cols;
    colsum_finalize_kernel<<<static_cast<unsigned>((cols + 7) / 8), 256, 0, as_stream(stream)>>>(pdw, pdb, static_cast<int>(c.grid), cols, dweight,
                                                                                               dbias, accumulate_params ? 1 : 0);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

