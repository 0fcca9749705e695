// Shared host/device helpers for libspq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/spq_b200.h"

namespace spq {

// ---- host-side error / bookkeeping ----------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int sm_count();

#define SPQ_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            spq::set_error(__VA_ARGS__);            \
            return SPQ_ERR_INVALID;                 \
        }                                           \
    } while (0)

#define SPQ_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            spq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                           __FILE__, __LINE__);                                        \
            return SPQ_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define SPQ_LAUNCH_OK()                                                                \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            spq::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                        \
            return SPQ_ERR_CUDA;                                                       \
        }                                                                              \
        spq::count_launch();                                                           \
    } while (0)

namespace stats {
// spq_stats.cu: fold per-CTA column partials [chunks, C] into the running statistics (log2 applied after the fold)
int finalize_partials(const float* pmin, const float* pmax, long long C, int chunks, int log_mode, float eps, int accumulate,
                      const int32_t* flags, float* stat_min, float* stat_max, int32_t* state, cudaStream_t st);
}

inline cudaStream_t as_stream(spq_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__
// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    // streaming read: do not pollute L1 (the data is touched once)
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// typed streaming loads: the activation-side kernels accept float32 or float16 inputs (fp16 values are
// widened exactly, so results equal those on x.float())
template <typename XT> __device__ __forceinline__ float4 ld_stream_x4(const XT* p);
template <> __device__ __forceinline__ float4 ld_stream_x4<float>(const float* p) { return ld_stream_f4(p); }
template <> __device__ __forceinline__ float4 ld_stream_x4<__half>(const __half* p) {
    unsigned lo, hi;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float ld_x1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_x1(const __half* p) { return __half2float(__ldg(p)); }

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// torch.min/max semantics: NaN wins.
__device__ __forceinline__ float nan_min(float a, float b) { return (b < a || b != b) ? b : a; }
__device__ __forceinline__ float nan_max(float a, float b) { return (b > a || b != b) ? b : a; }

// Correctly rounded float32 log2 (up to the ~2^-29 double-rounding cases): the definition the
// oracle uses (oracle/quant_oracle.py:log2_cr) and torch-CPU realises on 99.987 % of inputs.
__device__ __forceinline__ float log2_cr(float a) { return __double2float_rn(log2(static_cast<double>(a))); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nan_min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_fmax(float v) {   // NaN-ignoring
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// saturating float -> fp16 bits
__device__ __forceinline__ unsigned short f2h_sat(float v) {
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}
// two floats -> packed fp16x2 (lo in bits 0..15), saturating: ONE F2FP instead of two conversions and a byte merge
__device__ __forceinline__ unsigned int pack_h2(float lo, float hi) {
    unsigned int r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
#endif  // __CUDACC__

}  // namespace spq
