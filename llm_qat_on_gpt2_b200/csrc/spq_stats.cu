// Calibration statistics: per-channel / per-tensor min-max reduction (HBM-bound, one read of x).
//
// Replaces LearnableFakeQuantize._collect_statistics_batch, _reduce_min_max and
// finish_calibration of the reference (p1/quantization.py:104-139, 152-162, 174-209), which
// issue one min and one max pass per reduced dimension plus abs/gt/any/clamp/log2 temporaries.
// Here: one streaming pass (float4, L1::no_allocate) produces partial min/max (of x, or of |x|
// in log mode) and the any(|x| > eps) flag; a second tiny kernel folds the partials, applies
// log2 (after the reduction -- log2 is monotone, so min(log2|x|) == log2(min|x|) bit for bit)
// and merges into the running statistics.  No host synchronisation.
#include "spq_common.cuh"

namespace spq {
namespace stats {

constexpr int TX = 32, TY = 8;            // block = 32 x 8 threads
constexpr int COLS_PER_BLOCK_V4 = TX * 4; // 128 columns per block (float4 per thread)

struct Ws {
    int32_t* flags;   // [0] any(|x|>eps)
    float* pmin;      // [chunks, C]
    float* pmax;
};

static inline int col_chunks(int64_t rows, int64_t cols) {
    const int64_t col_tiles = (cols + COLS_PER_BLOCK_V4 - 1) / COLS_PER_BLOCK_V4;
    int64_t want = (static_cast<int64_t>(148) * 8 + col_tiles - 1) / col_tiles;   // ~8 CTAs per SM
    const int64_t max_chunks = (rows + 4 * TY - 1) / (4 * TY);                     // >= 32 rows per chunk
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

// Partial reduction over rows for a tile of columns.  VEC = 4: float4 loads (cols % 4 == 0 and x
// 16-byte aligned); VEC = 1: scalar fallback.
template <int VEC, bool LOG, typename XT>
__global__ void __launch_bounds__(TX* TY)
colstats_partial_kernel(const XT* __restrict__ x, long long rows, long long cols, float eps, long long rows_per_chunk,
                        float* __restrict__ pmin, float* __restrict__ pmax, int32_t* __restrict__ flags) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long c0 = (static_cast<long long>(blockIdx.x) * TX + tx) * VEC;
    const long long r_begin = static_cast<long long>(blockIdx.y) * rows_per_chunk;
    long long r_end = r_begin + rows_per_chunk;
    if (r_end > rows) r_end = rows;

    float mn[VEC], mx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    unsigned nan_cols = 0;   // bit j: column c0 + j saw a NaN (fminf/fmaxf drop NaN; torch keeps it)
    bool any = false;

    if (c0 < cols) {
        const XT* p = x + c0;
        long long r = r_begin + ty;
        if constexpr (VEC == 4) {
            // 4 independent 16-byte loads in flight per thread
            for (; r + 3 * TY < r_end; r += 4 * TY) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ld_stream_x4<XT>(p + (r + u * TY) * cols);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a = LOG ? fabsf(e[j]) : e[j];
                        nan_cols |= (a != a) ? (1u << j) : 0u;
                        if (LOG) any |= (a > eps);
                        mn[j] = fminf(mn[j], a);
                        mx[j] = fmaxf(mx[j], a);
                    }
                }
            }
        }
        for (; r < r_end; r += TY) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (c0 + j < cols) {
                    float a = ld_x1(p + r * cols + j);
                    if (LOG) a = fabsf(a);
                    nan_cols |= (a != a) ? (1u << j) : 0u;
                    if (LOG) any |= (a > eps);
                    mn[j] = fminf(mn[j], a);
                    mx[j] = fmaxf(mx[j], a);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        if (nan_cols & (1u << j)) { mn[j] = NAN; mx[j] = NAN; }
    __shared__ float s_mn[TY][TX * VEC + 1];
    __shared__ float s_mx[TY][TX * VEC + 1];
    __shared__ int s_any;
    if (tx == 0 && ty == 0) s_any = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) { s_mn[ty][tx * VEC + j] = mn[j]; s_mx[ty][tx * VEC + j] = mx[j]; }
    __syncthreads();
    if (LOG && any) s_any = 1;
    __syncthreads();
    const int lin = ty * TX + tx;
    for (int c = lin; c < TX * VEC; c += TX * TY) {
        float a = s_mn[0][c], b = s_mx[0][c];
#pragma unroll
        for (int y = 1; y < TY; ++y) { a = nan_min(a, s_mn[y][c]); b = nan_max(b, s_mx[y][c]); }
        const long long col = static_cast<long long>(blockIdx.x) * TX * VEC + c;
        if (col < cols) {
            pmin[static_cast<long long>(blockIdx.y) * cols + col] = a;
            pmax[static_cast<long long>(blockIdx.y) * cols + col] = b;
        }
    }
    if (LOG && lin == 0 && s_any) atomicOr(flags, 1);
}

// One warp per row (weights: channel_dim = 0).
template <bool LOG>
__global__ void __launch_bounds__(256)
rowstats_kernel(const float* __restrict__ x, long long rows, long long cols, float eps, float* __restrict__ pmin,
                float* __restrict__ pmax, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* p = x + row * cols;
    float mn = INFINITY, mx = -INFINITY;
    bool seen_nan = false, any = false;
    auto upd = [&](float a) {
        if (LOG) a = fabsf(a);
        seen_nan |= (a != a);
        if (LOG) any |= (a > eps);
        mn = fminf(mn, a);
        mx = fmaxf(mx, a);
    };
    if ((cols & 3) == 0 && aligned16_dev(p)) {
        for (long long c = lane * 4; c < cols; c += 128) {
            const float4 v = ld_stream_f4(p + c);
            upd(v.x); upd(v.y); upd(v.z); upd(v.w);
        }
    } else {
        for (long long c = lane; c < cols; c += 32) upd(__ldg(p + c));
    }
    if (seen_nan) { mn = NAN; mx = NAN; }
    mn = warp_min(mn);
    mx = warp_max(mx);
    const unsigned anyw = __ballot_sync(0xffffffffu, any);
    if (lane == 0) {
        pmin[row] = mn;
        pmax[row] = mx;
        if (LOG && anyw) atomicOr(flags, 1);
    }
}

// Fold `chunks` partials per channel, apply the log transform, merge into the running statistics.
// collapse != 0: per-tensor -- all channels fold into stat[0] (single block).
template <bool LOG>
__global__ void __launch_bounds__(1024)
stats_finalize_kernel(const float* __restrict__ pmin, const float* __restrict__ pmax, long long C, int chunks, int collapse,
                      float eps, int accumulate, const int32_t* __restrict__ flags, float* __restrict__ stat_min,
                      float* __restrict__ stat_max, int32_t* __restrict__ state) {
    const bool any = LOG ? (flags[0] != 0) : true;
    const float log_eps = LOG ? log2_cr(eps) : 0.f;
    auto emit = [&](long long idx, float a, float b) {
        if (LOG) {
            if (any) {
                a = log2_cr(a < eps ? eps : a);      // torch.clamp(min=eps) keeps NaN
                b = log2_cr(b < eps ? eps : b);
            } else {
                if (accumulate) return;              // nothing above eps: statistics unchanged
                a = log_eps; b = log_eps;            // first batch: filled with log2(eps)
            }
        }
        if (accumulate) {
            a = nan_min(stat_min[idx], a);
            b = nan_max(stat_max[idx], b);
        }
        stat_min[idx] = a;
        stat_max[idx] = b;
    };
    if (!collapse) {
        // 32 channels per block, blockDim/32 lanes share the fold over the `chunks` partials of a
        // channel (a single thread walking ~200 dependent-latency loads made this kernel as slow as
        // the streaming pass itself)
        __shared__ float fa[32][33], fb[32][33];
        const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5, nl = blockDim.x >> 5;
        const long long c = static_cast<long long>(blockIdx.x) * 32 + cl;
        float a = INFINITY, b = -INFINITY;
        if (c < C) {
#pragma unroll 4
            for (int k = sl; k < chunks; k += nl) {
                a = nan_min(a, pmin[static_cast<long long>(k) * C + c]);
                b = nan_max(b, pmax[static_cast<long long>(k) * C + c]);
            }
        }
        fa[sl][cl] = a; fb[sl][cl] = b;
        __syncthreads();
        if (sl == 0 && c < C) {
            for (int w = 1; w < nl; ++w) { a = nan_min(a, fa[w][cl]); b = nan_max(b, fb[w][cl]); }
            emit(c, a, b);
        }
    } else {
        float a = INFINITY, b = -INFINITY;
        const long long total = C * chunks;
        for (long long i = threadIdx.x; i < total; i += blockDim.x) { a = nan_min(a, pmin[i]); b = nan_max(b, pmax[i]); }
        a = warp_min(a); b = warp_max(b);
        __shared__ float sa[32], sb[32];
        if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (blockDim.x >> 5); ++w) { a = nan_min(a, sa[w]); b = nan_max(b, sb[w]); }
            emit(0, a, b);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && state && any) atomicOr(state, 1);
}

// scale / zero-point from calibrated statistics (p1/quantization.py:104-139)
__device__ __forceinline__ void finish_one(float lo, float hi, int qtype, int symmetric, float levels, float eps,
                                           float& scale, float& zp) {
    if (qtype == SPQ_LOG) {                       // :110-116
        zp = lo;
        scale = __fsub_rn(hi, lo);
    } else if (symmetric) {                       // :118-122
        float a = nan_max(fabsf(lo), fabsf(hi));
        a = (a < eps) ? eps : a;
        scale = __fdiv_rn(a, levels);
        zp = 0.f;
    } else {                                      // :123-127
        float r = __fsub_rn(hi, lo);
        r = (r < eps) ? eps : r;
        const float s = __fdiv_rn(r, levels);
        scale = s;
        zp = rintf(__fdiv_rn(-lo, s));
    }
}

__global__ void finish_calibration_kernel(const float* __restrict__ rmin, const float* __restrict__ rmax, long long n, int qtype,
                                          int symmetric, float levels, float eps, float* __restrict__ scale,
                                          float* __restrict__ zp) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, z;
    finish_one(rmin[i], rmax[i], qtype, symmetric, levels, eps, s, z);
    scale[i] = s;
    zp[i] = z;
}

// Single-batch calibration of MANY small tensors in one launch (the LoRA A/B quantisers of all 48 linears are
// recalibrated on their own weights every training step, p1/train_sp.py:125-163: 96 x (memset + 3 kernels) made
// that phase host-bound).  blockIdx.y = job; per-column jobs: blockIdx.x = block of 32 columns, 8 row lanes;
// per-row jobs: blockIdx.x = block of 8 rows, one warp each; per-tensor jobs: block 0 only.  Statistics,
// log transform and scale / zero-point are those of spq_minmax_stats + spq_finish_calibration bit for bit;
// flags[job] = any(|x| > eps) (log mode; 1 otherwise) -- a log job without data is redone by the caller.
__global__ void __launch_bounds__(256)
calibrate_many_kernel(const SpqCalibJob* __restrict__ jobs, int32_t* __restrict__ flags) {
    const SpqCalibJob jb = jobs[blockIdx.y];
    const bool LOG = jb.qtype == SPQ_LOG;
    const float eps = jb.eps;
    const float levels = jb.levels;
    __shared__ float s_a[8][33], s_b[8][33];
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    bool any = false;
    auto tf = [&](float a) { return LOG ? log2_cr(a < eps ? eps : a) : a; };      // clamp keeps NaN
    if (jb.bcast == SPQ_PER_COL) {
        const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
        const long long c = static_cast<long long>(blockIdx.x) * 32 + cl;
        if (static_cast<long long>(blockIdx.x) * 32 >= jb.cols) return;
        float a = INFINITY, b = -INFINITY;
        bool isnan_ = false;
        if (c < jb.cols) {
            for (long long r = rl; r < jb.rows; r += 8) {
                float v = __ldg(jb.x + r * jb.cols + c);
                if (LOG) v = fabsf(v);
                isnan_ |= (v != v);
                any |= (v > eps);
                a = fminf(a, v); b = fmaxf(b, v);
            }
        }
        if (isnan_) { a = NAN; b = NAN; }
        s_a[rl][cl] = a; s_b[rl][cl] = b;
        if (LOG && any) s_any = 1;
        __syncthreads();
        if (rl == 0 && c < jb.cols) {
#pragma unroll
            for (int w = 1; w < 8; ++w) { a = nan_min(a, s_a[w][cl]); b = nan_max(b, s_b[w][cl]); }
            a = tf(a); b = tf(b);
            float s, z;
            finish_one(a, b, jb.qtype, jb.symmetric, levels, eps, s, z);
            jb.rmin[c] = a; jb.rmax[c] = b; jb.scale[c] = s; jb.zp[c] = z;
        }
    } else if (jb.bcast == SPQ_PER_ROW) {
        const int lane = threadIdx.x & 31;
        const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
        if (static_cast<long long>(blockIdx.x) * 8 >= jb.rows) return;
        float a = INFINITY, b = -INFINITY;
        bool isnan_ = false;
        if (r < jb.rows) {
            for (long long c = lane; c < jb.cols; c += 32) {
                float v = __ldg(jb.x + r * jb.cols + c);
                if (LOG) v = fabsf(v);
                isnan_ |= (v != v);
                any |= (v > eps);
                a = fminf(a, v); b = fmaxf(b, v);
            }
        }
        if (__any_sync(0xffffffffu, isnan_)) { a = NAN; b = NAN; }
        a = warp_min(a); b = warp_max(b);
        if (LOG && any) s_any = 1;
        __syncthreads();
        if (lane == 0 && r < jb.rows) {
            a = tf(a); b = tf(b);
            float s, z;
            finish_one(a, b, jb.qtype, jb.symmetric, levels, eps, s, z);
            jb.rmin[r] = a; jb.rmax[r] = b; jb.scale[r] = s; jb.zp[r] = z;
        }
    } else {
        if (blockIdx.x != 0) return;
        float a = INFINITY, b = -INFINITY;
        bool isnan_ = false;
        const long long total = jb.rows * jb.cols;
        for (long long i = threadIdx.x; i < total; i += blockDim.x) {
            float v = __ldg(jb.x + i);
            if (LOG) v = fabsf(v);
            isnan_ |= (v != v);
            any |= (v > eps);
            a = fminf(a, v); b = fmaxf(b, v);
        }
        if (isnan_) { a = NAN; b = NAN; }
        a = warp_min(a); b = warp_max(b);
        if ((threadIdx.x & 31) == 0) { s_a[0][threadIdx.x >> 5] = a; s_b[0][threadIdx.x >> 5] = b; }
        if (LOG && any) s_any = 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) { a = nan_min(a, s_a[0][w]); b = nan_max(b, s_b[0][w]); }
            a = tf(a); b = tf(b);
            float s, z;
            finish_one(a, b, jb.qtype, jb.symmetric, levels, eps, s, z);
            jb.rmin[0] = a; jb.rmax[0] = b; jb.scale[0] = s; jb.zp[0] = z;
        }
    }
    if (threadIdx.x == 0 && (!LOG || s_any)) atomicOr(flags + blockIdx.y, 1);
}

}  // namespace stats
}  // namespace spq

using namespace spq;
using namespace spq::stats;

extern "C" size_t spq_stats_workspace_bytes(int64_t rows, int64_t cols, int bcast) {
    if (rows <= 0 || cols <= 0) return 256;
    size_t n;
    if (bcast == SPQ_PER_ROW) n = static_cast<size_t>(rows);
    else n = static_cast<size_t>(col_chunks(rows, cols)) * static_cast<size_t>(cols);
    return 256 + 2 * n * sizeof(float);
}

// Fold [chunks, C] partials produced by another kernel of the library (the LayerNorm-fused statistics pass of
// spq_quantize.cu) into the running statistics: the second half of spq_minmax_stats, per-column layout.
int spq::stats::finalize_partials(const float* pmin, const float* pmax, long long C, int chunks, int log_mode, float eps,
                                  int accumulate, const int32_t* flags, float* stat_min, float* stat_max, int32_t* state,
                                  cudaStream_t st) {
    const unsigned fgrid = static_cast<unsigned>((C + 31) / 32);
    const unsigned fthreads = chunks > 64 ? 1024u : 256u;
    if (log_mode)
        stats_finalize_kernel<true><<<fgrid, fthreads, 0, st>>>(pmin, pmax, C, chunks, 0, eps, accumulate, flags, stat_min, stat_max, state);
    else
        stats_finalize_kernel<false><<<fgrid, fthreads, 0, st>>>(pmin, pmax, C, chunks, 0, eps, accumulate, flags, stat_min, stat_max, state);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

template <int VEC, typename XT>
static void launch_colstats(dim3 grid, dim3 block, cudaStream_t st, int log_mode, const void* x, long long rows, long long cols,
                            float eps, long long rpc, float* pmin, float* pmax, int32_t* flags) {
    const XT* xt = static_cast<const XT*>(x);
    if (log_mode) colstats_partial_kernel<VEC, true, XT><<<grid, block, 0, st>>>(xt, rows, cols, eps, rpc, pmin, pmax, flags);
    else colstats_partial_kernel<VEC, false, XT><<<grid, block, 0, st>>>(xt, rows, cols, eps, rpc, pmin, pmax, flags);
}

extern "C" int spq_minmax_stats(const void* x, int x_is_half, int64_t rows, int64_t cols, int bcast, int log_mode, float eps,
                                float* stat_min, float* stat_max, int accumulate, int32_t* state, void* workspace,
                                size_t workspace_bytes, spq_stream_t stream) {
    SPQ_REQUIRE(x && stat_min && stat_max && workspace, "spq_minmax_stats: null pointer");
    SPQ_REQUIRE(rows > 0 && cols > 0, "spq_minmax_stats: empty tensor [%lld, %lld]", (long long)rows, (long long)cols);
    SPQ_REQUIRE(bcast == SPQ_PER_TENSOR || bcast == SPQ_PER_ROW || bcast == SPQ_PER_COL, "spq_minmax_stats: bad bcast %d", bcast);
    SPQ_REQUIRE(workspace_bytes >= spq_stats_workspace_bytes(rows, cols, bcast), "spq_minmax_stats: workspace too small");
    cudaStream_t st = as_stream(stream);
    Ws ws;
    ws.flags = reinterpret_cast<int32_t*>(workspace);
    ws.pmin = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
    SPQ_CUDA_OK(cudaMemsetAsync(ws.flags, 0, 16, st));
    long long C;
    int chunks;
    if (bcast == SPQ_PER_ROW) {
        C = rows; chunks = 1;
        ws.pmax = ws.pmin + C;
        const int warps = 8;
        const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
        SPQ_REQUIRE(!x_is_half, "spq_minmax_stats: per-row (weight) statistics take float32 input");
        const float* xf = static_cast<const float*>(x);
        if (log_mode) rowstats_kernel<true><<<grid, warps * 32, 0, st>>>(xf, rows, cols, eps, ws.pmin, ws.pmax, ws.flags);
        else rowstats_kernel<false><<<grid, warps * 32, 0, st>>>(xf, rows, cols, eps, ws.pmin, ws.pmax, ws.flags);
        SPQ_LAUNCH_OK();
    } else {
        C = cols; chunks = col_chunks(rows, cols);
        ws.pmax = ws.pmin + static_cast<size_t>(chunks) * C;
        const long long rpc = (rows + chunks - 1) / chunks;
        const bool vec = ((cols & 3) == 0) && (x_is_half ? (reinterpret_cast<uintptr_t>(x) & 7u) == 0 : aligned16(x));
        dim3 block(TX, TY);
        if (vec) {
            dim3 grid(static_cast<unsigned>((cols + COLS_PER_BLOCK_V4 - 1) / COLS_PER_BLOCK_V4), chunks);
            if (x_is_half) launch_colstats<4, __half>(grid, block, st, log_mode, x, rows, cols, eps, rpc, ws.pmin, ws.pmax, ws.flags);
            else launch_colstats<4, float>(grid, block, st, log_mode, x, rows, cols, eps, rpc, ws.pmin, ws.pmax, ws.flags);
        } else {
            dim3 grid(static_cast<unsigned>((cols + TX - 1) / TX), chunks);
            if (x_is_half) launch_colstats<1, __half>(grid, block, st, log_mode, x, rows, cols, eps, rpc, ws.pmin, ws.pmax, ws.flags);
            else launch_colstats<1, float>(grid, block, st, log_mode, x, rows, cols, eps, rpc, ws.pmin, ws.pmax, ws.flags);
        }
        SPQ_LAUNCH_OK();
    }
    const int collapse = (bcast == SPQ_PER_TENSOR) ? 1 : 0;
    const unsigned fgrid = collapse ? 1u : static_cast<unsigned>((C + 31) / 32);
    const unsigned fthreads = (!collapse && chunks > 64) ? 1024u : 256u;
    if (log_mode)
        stats_finalize_kernel<true><<<fgrid, fthreads, 0, st>>>(ws.pmin, ws.pmax, C, chunks, collapse, eps, accumulate, ws.flags, stat_min, stat_max, state);
    else
        stats_finalize_kernel<false><<<fgrid, fthreads, 0, st>>>(ws.pmin, ws.pmax, C, chunks, collapse, eps, accumulate, ws.flags, stat_min, stat_max, state);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_finish_calibration(const float* running_min, const float* running_max, int64_t n, int qtype, int symmetric,
                                      int bits, float eps, float* scale, float* zero_point, spq_stream_t stream) {
    SPQ_REQUIRE(running_min && running_max && scale && zero_point && n > 0, "spq_finish_calibration: bad arguments");
    SPQ_REQUIRE(bits >= 1 && bits <= 32, "spq_finish_calibration: bits %d", bits);
    const double levels = symmetric ? (static_cast<double>(1ull << (bits - 1)) - 1.0) : (static_cast<double>(1ull << bits) - 1.0);
    finish_calibration_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        running_min, running_max, n, qtype, symmetric, static_cast<float>(levels), eps, scale, zero_point);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

extern "C" int spq_calibrate_many(const SpqCalibJob* jobs_dev, int32_t n_jobs, int32_t max_blocks, int32_t* flags_dev,
                                  spq_stream_t stream) {
    SPQ_REQUIRE(jobs_dev && flags_dev && n_jobs > 0 && max_blocks > 0 && n_jobs <= 65535, "spq_calibrate_many: bad arguments");
    cudaStream_t st = as_stream(stream);
    SPQ_CUDA_OK(cudaMemsetAsync(flags_dev, 0, sizeof(int32_t) * static_cast<size_t>(n_jobs), st));
    calibrate_many_kernel<<<dim3(static_cast<unsigned>(max_blocks), static_cast<unsigned>(n_jobs)), 256, 0, st>>>(jobs_dev, flags_dev);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}
