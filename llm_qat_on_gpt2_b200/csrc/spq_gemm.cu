// Fused quant-GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma (kind::f16,
// fp32 accumulation in TMEM) -> tcgen05.ld epilogue (scale / clamp / bias / residual).
//
// Replaces F.linear(q(x), q(W), b) + LoRALayer.forward of the reference (p1/lora.py:45-54,
// 144-150): the operands are the fp16 code / dequant tensors produced by spq_quantize.cu, the
// LoRA up-projection is folded in as one more K segment (A2, B2) of the same accumulator.
//
// Layout of one CTA (320 threads, persistent over output tiles):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9  epilogue: TMEM lane quadrant (warp_idx % 4), column half ((warp_idx - 2) / 4):
//               tcgen05.ld -> scale/clamp/bias in registers -> swizzled smem tile -> TMA bulk store
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), two TMEM accumulators with
// tmem_full/tmem_empty mbarriers (MMA <-> epilogue) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>

#include "spq_common.cuh"

namespace spq {
namespace gemm {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 64;         // 64 fp16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
// Epilogue warps: EPI_WARPS / 4 per TMEM lane quadrant, each owning BN / NPART columns of its 32 rows.  8 is the
// default.  -DSPQ_EPI_WARPS=16 (four instruction streams per SM sub-partition, one staging tile per warp) was measured
// on the same box and is no faster (c_fc +GELU 179.8 vs 177.9 us, fp32 store 160 vs 152 us): the kernel is bound by the
// bytes that cross the L2 <-> SM fabric (operand tiles in, output tiles out), not by epilogue latency.
#ifndef SPQ_EPI_WARPS
#define SPQ_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = SPQ_EPI_WARPS;
static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps: 2 or 4 per TMEM lane quadrant");
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int NPART = EPI_WARPS / 4;               // column parts of a tile (one per warp of a lane quadrant)
constexpr int NUM_THREADS = 64 + EPI_THREADS;      // 2 control warps + the epilogue warps
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr uint32_t SPIN_LIMIT = 1u << 24;
constexpr int STG_TILE_BYTES = 32 * 32 * 4;   // per-epilogue-warp 32 x 32 fp32 staging tile, 128B-XOR-swizzled
#ifndef SPQ_STG_BUFS
#define SPQ_STG_BUFS (SPQ_EPI_WARPS >= 16 ? 1 : 2)
#endif
// 8 warps: two tiles per warp (chunk i+1 is computed while the TMA engine still reads chunk i); 16 warps: one tile
// per warp (the same 64 KB in total) -- the other three warps of the sub-partition cover the wait
constexpr int STG_BUFS = SPQ_STG_BUFS;

__device__ int g_abort = 0;    // watchdog: set when a pipeline wait timed out

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a broken pipeline must never hang the GPU, and must never pass silently.  On timeout (2^24
// failed try_waits, seconds) the kernel TRAPS: the launch fails, the CUDA context reports the error at the next
// synchronising call of the process (torch raises, bench.py exits non-zero without a JSON line), and no later
// kernel can run on garbage.  g_abort is kept for the mis-aligned-smem guard (spq_debug_status()).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t it = 0;; ++it) {
        if (mbar_try_wait(bar, parity)) return;
        if ((it & 1023u) == 1023u && it >= SPIN_LIMIT) {
            atomicExch(&g_abort, 1);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster (same TPC) run one 256-row MMA; each loads its own
// 128 rows of A and HALF of the B tile, so every operand byte is fetched from L2 once per pair, not once per SM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {   // same offset, in CTA `cta_rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the completion bytes are signalled on `bar_cluster_addr`, a shared::cluster address (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// Multicast variant (clusters of two CTA pairs): the box lands at the same CTA-relative offset in every CTA of
// `cta_mask`, and each destination signals the barrier at the CTA-relative offset of `bar_local` in the LEADER of its own
// pair (cta_group::2 honours the peer bit of the barrier address: bit 24 of the shared-window address, cleared here).
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_local, int c0, int c1,
                                                    uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_local & 0xFEFFFFFFu), "h"(cta_mask),
          "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// e4m3 x e4m3 -> fp32 (kind::f8f6f4, K = 32 per instruction): the operands of a <= 4-bit integer-code GEMM
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), 128B swizzle.
//   K-major : rows of 128 B (64 fp16 along K), 8-row groups 1024 B apart (SBO); LBO unused (=1).
//   MN-major: rows of 128 B (64 fp16 along M/N) indexed by k, 8-k groups 1024 B apart (SBO),
//             consecutive 64-element M/N blocks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: fp16 x fp16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_f16(int umma_m, int umma_n, int a_mn_major, int b_mn_major) {
    return (1u << 4)                                   // c_format = F32
           | (0u << 7) | (0u << 10)                    // a_format = b_format = F16
           | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16)
           | (static_cast<uint32_t>(umma_n >> 3) << 17) | (static_cast<uint32_t>(umma_m >> 4) << 24);
}

// Exact-form (erf) GELU of nn.GELU().  erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. <= 1e-7 |v|
// on the output), arranged as 12 FMUL/FFMA + MUFU.RCP + MUFU.EX2: erff() costs ~30 instructions and the IEEE
// reciprocal / denormal-guarded exp another ~10, which shows in an epilogue that has ~18 issue slots per element.
//   gelu(v) = v/2 + |v|/2 * erf(|v|/sqrt 2),   erf(x) = 1 - (a1 t + ... + a5 t^5) exp(-x^2),  t = 1/(1 + p x)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_erf(float v) {
    const float x = fabsf(v) * 0.70710678118654752f;
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((v * v) * -0.72134752044448170f));   // exp(-v^2 / 2)
    const float e = fmaf(-(p * t), ex, 1.0f);                // erf(|v| / sqrt 2)
    const float hv = 0.5f * v;
    return fmaf(fabsf(hv), e, hv);
}

struct EpiParams {
    const float* row_scale;   // [M] or null
    const float* col_scale;   // [N] or null
    const float* bias;        // [N] or null
    const float* C;           // [M, ldc] or null (added after the clamp)
    const float* alpha_dev;   // device scalar multiplied into alpha, or null
    void* D;
    long long ldc, ldd;
    long long d_stride_n;     // 1 for row-major D; TN kernel may store transposed
    float alpha;
    float clamp_abs;          // <= 0: off
    int debug;                // profiling experiments only: 1 skip epilogue after tcgen05.ld, 2 skip MMA issue,
                              // 4 disable the TMA-store path, 8 skip the TMA store instruction, 16 skip the staging writes
    int tma_store;            // 1: D is written with TMA bulk tensor stores (16 B aligned rows; residual prefetched per lane)
    int act;                  // 0: none, 1: exact (erf) GELU applied after the bias
    int store_hint;           // 1: TMA stores carry an L2 evict-first policy (outputs much larger than L2)
    float2* lse_part;         // LSE kernels: [M, lse_ld] (max, sum of exp(v - max)) of each row over each column half-tile
    long long lse_ld;
};

template <int BN, bool CTA2 = false>
struct SmemLayout {
    static constexpr int B_TILE_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;      // CTA pair: each CTA stages half of the B tile
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int EPI_BYTES = 2 * BN * 4;      // col_scale + bias of the tile
    static constexpr int STG_BYTES = EPI_WARPS * STG_BUFS * STG_TILE_BYTES;
    static constexpr int RING_BYTES = 232448 - STG_BYTES - EPI_BYTES - 256;       // what is left for the operand ring
    static constexpr int STAGES = (RING_BYTES / STAGE_BYTES) > 8 ? 8 : (RING_BYTES / STAGE_BYTES);
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    // order: [stages][staging tiles][col_scale|bias][barriers]; the dynamic smem base must be 1024 B aligned
    // (checked at run time) -- there is no room for alignment slack at BN = 256
    static constexpr int TOTAL = STAGES * STAGE_BYTES + STG_BYTES + EPI_BYTES + BAR_BYTES;
    static_assert(TOTAL <= 232448, "shared memory budget");
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// D[m,n] = epi( sum_k A[m,k] B[n,k] + sum_j A2[m,j] B2[n,j] ), all operands K-major fp16.
// Tile order of the persistent CTAs.  Few column tiles (the layer GEMMs: 3-12): row-major, the ~148 tiles in
// flight share one or two A row-blocks and all of B sits in L2.  Many column tiles (LM head: 197): groups of
// GROUP_M row-blocks are swept column by column, so the tiles in flight touch ~16 A blocks x ~9 B blocks; with
// row-major order every row-block re-read the whole 77 MB B while the 6.6 GB output stream kept evicting it
// (ncu: 7-8 GB of DRAM reads for 0.13 GB of operands).
constexpr int GROUP_M = 16;
__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
    if (n_tiles < 2 * GROUP_M) {
        m_blk = t / n_tiles;
        n_blk = t - m_blk * n_tiles;
        return;
    }
    const int per_group = GROUP_M * n_tiles;
    const int g = t / per_group;
    const int rem = t - g * per_group;
    const int gm = min(GROUP_M, m_tiles - g * GROUP_M);     // the last group may be short
    n_blk = rem / gm;
    m_blk = g * GROUP_M + (rem - n_blk * gm);
}

// PRE_C: the TMA-store epilogue adds a float32 residual (separate instantiation: its 32 prefetch registers and
// extra staging traffic stay out of the plain kernel)
// LSE: the epilogue also keeps, per output row, the running (max, sum exp) of the values it stores and writes one
// pair per column half-tile -- the log-sum-exp of the LM head's logits without a second pass over them
// CTA2: launched as clusters of two CTAs; the pair computes 256 x BN tiles with tcgen05.mma.cta_group::2 (issued
// by the even CTA), each CTA owning 128 rows of A, half of B in shared memory and its 128 rows of the accumulator.
// F8: the first K segment (A, B) holds e4m3 bytes -- 128 of them per 128-byte swizzle row, tcgen05.mma.kind::f8f6f4
// (instruction-descriptor format code 0 = E4M3, the same bits as F16 under kind::f16); the second segment (the fp16
// LoRA operands) keeps kind::f16 and lands in the same fp32 accumulator.
template <int BN, bool OUT_HALF, bool PRE_C, bool LSE = false, bool CTA2 = false, bool F8 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                const __grid_constant__ CUtensorMap tmD, int M, int N, int kb1, int kb2, EpiParams ep) {
    using L = SmemLayout<BN, CTA2>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stg_all = smem + L::STAGES * L::STAGE_BYTES;                       // EPI_WARPS x 4 KB, 1024 B aligned
    float* epi_cs = reinterpret_cast<float*>(stg_all + L::STG_BYTES);
    float* epi_bias = epi_cs + BN;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_bias + BN);
    uint64_t* empty_bar = full_bar + L::STAGES;
    uint64_t* tmem_full = empty_bar + L::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int TILE_M = CTA2 ? 2 * BM : BM;                     // rows per tile (per CTA pair in CTA2 mode)
    // CTA2 kernels run as clusters of ONE pair (2 CTAs) or TWO pairs (4 CTAs, "quad"): in a quad the pairs work on
    // vertically adjacent 256-row tiles of the same column block, so they need the same B tile -- every CTA fetches a
    // QUARTER of it and multicasts the box into the CTA at its position in the other pair: the B bytes cross the
    // L2 -> SM fabric once per quad instead of once per pair (the kernel is bound by that traffic)
    const uint32_t cluster_rank = CTA2 ? cluster_ctarank() : 0u;
    const bool quad_cluster = CTA2 && cluster_nctarank() == 4u;
    const uint32_t cta_rank = cluster_rank & 1u;                   // rank inside the CTA pair
    const uint32_t pair_id = cluster_rank >> 1;                    // which pair of the cluster (0 unless quad)
    const uint32_t lead_rank = pair_id << 1;                       // cluster rank of this pair's leader (it issues the MMAs)
    const int cpc = CTA2 ? (quad_cluster ? 4 : 2) : 1;             // CTAs per cluster
    const int tile_first = static_cast<int>(blockIdx.x) / cpc;
    const int tile_step = static_cast<int>(gridDim.x) / cpc;
    const int tile_rows = quad_cluster ? 2 * TILE_M : TILE_M;      // rows one cluster covers per scheduled tile
    const int row_off = static_cast<int>(pair_id) * TILE_M + static_cast<int>(cta_rank) * BM;   // this CTA's rows inside it
    const int n_tiles = (N + BN - 1) / BN;
    const int m_tiles = (M + tile_rows - 1) / tile_rows;
    const int num_tiles = n_tiles * m_tiles;
    const int kb_total = kb1 + kb2;

    // the 128B swizzle of TMA / UMMA / the staging tiles assumes a 1024 B aligned base
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) atomicExch(&g_abort, 2);
        __trap();
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (kb2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
        if (ep.tma_store) tma_prefetch_desc(&tmD);
        for (int s = 0; s < L::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], quad_cluster ? 2 : 1);     // quad: both pairs' MMAs must have retired before a slot is rewritten
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], EPI_WARPS * (CTA2 ? 2 : 1));     // one arrive per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "n"(L::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "n"(L::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();           // the peer's barriers exist before anything is signalled on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = tile_first; t < num_tiles; t += tile_step) {
                int m_blk, n_blk;
                tile_coords(t, m_tiles, n_tiles, m_blk, n_blk);
                const int m0 = m_blk * tile_rows + row_off;
                const int n0 = n_blk * BN + (CTA2 ? static_cast<int>(cta_rank) * (BN / 2) : 0);   // this CTA's half of B
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + A_TILE_BYTES;
                    if constexpr (CTA2) {
                        // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
                        const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), lead_rank);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        const CUtensorMap* ta = kb < kb1 ? &tmA : &tmA2;
                        const CUtensorMap* tb = kb < kb1 ? &tmB : &tmB2;
                        const int kc = kb < kb1 ? kb * (F8 ? 2 * BK : BK) : (kb - kb1) * BK;
                        tma_load_2d_pair(sa, ta, lead_bar, kc, m0);
                        if (!quad_cluster) {
                            tma_load_2d_pair(sb, tb, lead_bar, kc, n0);
                        } else {
                            // this CTA's quarter of the B tile (BN / 4 rows), into its own half-tile AND into the half-tile
                            // of the CTA at the same position in the other pair; the other quarter arrives from there
                            tma_load_2d_pair_mc(sb + pair_id * (L::B_TILE_BYTES / 2), tb, smem_u32(&full_bar[stage]), kc,
                                                n0 + static_cast<int>(pair_id) * (BN / 4),
                                                static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2))));
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        if (kb < kb1) {
                            tma_load_2d(sa, &tmA, &full_bar[stage], kb * (F8 ? 2 * BK : BK), m0);
                            tma_load_2d(sb, &tmB, &full_bar[stage], kb * (F8 ? 2 * BK : BK), n0);
                        } else {
                            tma_load_2d(sa, &tmA2, &full_bar[stage], (kb - kb1) * BK, m0);
                            tma_load_2d(sb, &tmB2, &full_bar[stage], (kb - kb1) * BK, n0);
                        }
                    }
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0 && cta_rank == 0) {              // CTA pair: the even CTA issues for both
            constexpr uint32_t idesc = make_idesc_f16(TILE_M, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = tile_first; t < num_tiles; t += tile_step) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint64_t adesc = make_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_desc_sw128(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (ep.debug & 2) break;
                        // +32 B per UMMA_K inside the 128 B swizzle span: +2 in the (addr >> 4) field
                        const uint64_t ad = adesc + static_cast<uint64_t>(2 * k), bd = bdesc + static_cast<uint64_t>(2 * k);
                        const uint32_t acc_on = (kb | k) != 0 ? 1u : 0u;
                        if (F8 && kb < kb1) {
                            if constexpr (CTA2) umma_f8_pair(tmem_d, ad, bd, idesc, acc_on);
                            else umma_f8(tmem_d, ad, bd, idesc, acc_on);
                        } else {
                            if constexpr (CTA2) umma_f16_pair(tmem_d, ad, bd, idesc, acc_on);
                            else umma_f16(tmem_d, ad, bd, idesc, acc_on);
                        }
                    }
                    if constexpr (CTA2) {
                        // frees the slot in every CTA that writes into it (quad: both pairs); the accumulator is the pair's
                        umma_commit_pair(&empty_bar[stage], static_cast<uint16_t>(quad_cluster ? 0xF : 0x3));
                        if (kb == kb_total - 1) umma_commit_pair(&tmem_full[acc], static_cast<uint16_t>(0x3u << lead_rank));
                    } else {
                        umma_commit(&empty_bar[stage]);          // frees the smem slot when the MMAs retire
                        if (kb == kb_total - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue (warps 2..9)
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        const int part = (warp - 2) >> 2;             // which part of the tile's columns
        const int epi_tid = threadIdx.x - 64;         // 0..EPI_THREADS-1
        constexpr int COLS_PER_PART = BN / NPART >= 32 ? BN / NPART : 32;
        constexpr int CHUNKS = COLS_PER_PART / 32;    // 32-column chunks per warp
        float* const stg_base = reinterpret_cast<float*>(stg_all + (warp - 2) * STG_BUFS * STG_TILE_BYTES);
        float* stg = stg_base;
        uint32_t stg_u32 = smem_u32(stg);
        int stg_buf = 0;
        int stores_in_flight = 0;                     // TMA store groups this warp has committed and not yet waited for
        int acc = 0;
        uint32_t acc_phase = 0;
        bool store_pending = false;
        const float alpha = ep.alpha * (ep.alpha_dev ? __ldg(ep.alpha_dev) : 1.0f);
        const bool active = part * COLS_PER_PART < BN;   // narrow tiles (BN = 64 with four parts) leave the upper parts idle
        uint64_t store_policy = 0;
        if (ep.store_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(store_policy));
        // residual C on the TMA-store path: the warp reads the NEXT chunk's 32 x 32 block of C coalesced (8 lanes
        // per 128 B row segment) into registers as soon as the current block has been parked in the staging tile;
        // the block goes through the same swizzled tile, so every lane then finds its own row next to its
        // accumulators (chunk 0 of the next tile is requested while that tile's MMA still runs)
        float4 cn[PRE_C ? 8 : 1];
        auto c_load = [&](int tt, int cc) {
            int m_blk, n_blk;
            tile_coords(tt, m_tiles, n_tiles, m_blk, n_blk);
            const int rb = m_blk * tile_rows + row_off + quad * 32 + (lane >> 3);
            const int ncol = n_blk * BN + part * COLS_PER_PART + cc * 32 + (lane & 7) * 4;
            const float* pc = ep.C + static_cast<long long>(rb) * ep.ldc + ncol;
#pragma unroll
            for (int i = 0; i < (PRE_C ? 8 : 1); ++i)
                cn[i] = (rb + i * 4 < M && ncol < N) ? *reinterpret_cast<const float4*>(pc + static_cast<long long>(i) * 4 * ep.ldc)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if constexpr (PRE_C) { if (active && tile_first < num_tiles) c_load(tile_first, 0); }
        for (int t = tile_first; t < num_tiles; t += tile_step) {
            int m_blk, n_blk;
            tile_coords(t, m_tiles, n_tiles, m_blk, n_blk);
            const int m0 = m_blk * tile_rows + row_off;
            const int n0 = n_blk * BN;
            // stage this tile's per-column parameters (previous tile's readers are done: barrier 1)
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            for (int j = epi_tid; j < BN; j += EPI_THREADS) {
                const int n = n0 + j;
                epi_cs[j] = (ep.col_scale && n < N) ? __ldg(ep.col_scale + n) : 1.0f;
                epi_bias[j] = (ep.bias && n < N) ? __ldg(ep.bias + n) : 0.0f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int rbase = m0 + quad * 32;                       // first global row of this warp's strip
            const int row = rbase + lane;
            const float rs = alpha * ((ep.row_scale && row < M) ? __ldg(ep.row_scale + row) : 1.0f);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + part * COLS_PER_PART);
            const int sw = lane & 7;                                // 128B swizzle: 16 B chunk index ^ (row % 8)
            float lse_m = -INFINITY, lse_s = 0.f;                   // LSE: this row over this warp's columns of the tile
            if (active) {
#pragma unroll 1
                for (int c = 0; c < CHUNKS; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
                    tmem_ld_wait();
                    const int cl = part * COLS_PER_PART + c * 32;   // first column of the chunk inside the tile
                    const int nc = n0 + cl;
                    auto c_next = [&]() {                           // issue the next chunk's residual loads
                        if (c + 1 < CHUNKS) c_load(t, c + 1);
                        else if (t + tile_step < num_tiles) c_load(t + tile_step, 0);
                    };
                    if (nc >= N || rbase >= M || (ep.debug & 1)) {              // warp-uniform
                        if constexpr (PRE_C) c_next();
                        continue;
                    }
                    if (ep.tma_store) {
                        // staging tile `stg_buf` was last handed to the TMA engine two chunks ago: at most ONE younger
                        // store group may still be reading (the other tile) when this one is overwritten
                        if (stores_in_flight >= STG_BUFS) {
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STG_BUFS - 1) : "memory");
                            __syncwarp();
                            stores_in_flight = STG_BUFS - 1;
                        }
                        stg = stg_base + stg_buf * (STG_TILE_BYTES / 4);
                        stg_u32 = smem_u32(stg);
                    }
                    if (ep.tma_store) {
                        if constexpr (PRE_C) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int rl = i * 4 + (lane >> 3);
                                *reinterpret_cast<float4*>(stg + rl * 32 + (((lane & 7) ^ (rl & 7)) << 2)) = cn[i];
                            }
                            c_next();
                            __syncwarp();
                        }
                        // finish the arithmetic in registers (this lane owns one row, 32 columns), write the row
                        // into the swizzled tile, then one lane hands the 32 x 32 block to the TMA engine
                        uint2 hold = make_uint2(0u, 0u);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 cs4 = *reinterpret_cast<const float4*>(epi_cs + cl + j);
                            const float4 bi4 = *reinterpret_cast<const float4*>(epi_bias + cl + j);
                            float4 o;
                            o.x = __uint_as_float(v[j]) * rs * cs4.x;     o.y = __uint_as_float(v[j + 1]) * rs * cs4.y;
                            o.z = __uint_as_float(v[j + 2]) * rs * cs4.z; o.w = __uint_as_float(v[j + 3]) * rs * cs4.w;
                            if (ep.clamp_abs > 0.f) {
                                o.x = fminf(fmaxf(o.x, -ep.clamp_abs), ep.clamp_abs); o.y = fminf(fmaxf(o.y, -ep.clamp_abs), ep.clamp_abs);
                                o.z = fminf(fmaxf(o.z, -ep.clamp_abs), ep.clamp_abs); o.w = fminf(fmaxf(o.w, -ep.clamp_abs), ep.clamp_abs);
                            }
                            o.x += bi4.x; o.y += bi4.y; o.z += bi4.z; o.w += bi4.w;
                            if constexpr (PRE_C) {
                                const float4 c4 = *reinterpret_cast<const float4*>(stg + lane * 32 + (((j >> 2) ^ sw) << 2));
                                o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w;
                            }
                            if (ep.act == 1) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
                            if constexpr (LSE) {
                                // the finished values stay in v[] for the chunk-level log-sum-exp below
                                v[j] = __float_as_uint(o.x); v[j + 1] = __float_as_uint(o.y);
                                v[j + 2] = __float_as_uint(o.z); v[j + 3] = __float_as_uint(o.w);
                            }
                            if constexpr (OUT_HALF) {
                                // 64 B rows, 64B swizzle: 16-byte chunk index ^ ((row / 2) % 4); one 16-byte store
                                // per 8 columns (8-byte stores put lanes l and l+8 on the same banks: ncu showed
                                // 2x the ideal shared-store wavefronts)
                                if ((j & 4) == 0) {
                                    hold = make_uint2(pack_h2(o.x, o.y), pack_h2(o.z, o.w));
                                } else {
                                    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(stg) + lane * 64 +
                                                              (((j >> 3) ^ ((lane >> 1) & 3)) << 4)) =
                                        make_uint4(hold.x, hold.y, pack_h2(o.x, o.y), pack_h2(o.z, o.w));
                                }
                            } else {
                                if (!(ep.debug & 16)) *reinterpret_cast<float4*>(stg + lane * 32 + (((j >> 2) ^ sw) << 2)) = o;
                                else if (o.x == 1.2345e-30f) ep.alpha_dev = nullptr;
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0 && !(ep.debug & 8)) {
                            if (ep.store_hint) {
                                // streaming output (>> L2, nobody reads it back soon): evict-first keeps the operand
                                // tiles of the tiles in flight resident while the store stream passes through L2
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                                                 reinterpret_cast<uint64_t>(&tmD)),
                                             "r"(stg_u32), "r"(nc), "r"(rbase), "l"(store_policy)
                                             : "memory");
                            } else {
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                                 reinterpret_cast<uint64_t>(&tmD)),
                                             "r"(stg_u32), "r"(nc), "r"(rbase)
                                             : "memory");
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        store_pending = true;
                        ++stores_in_flight;
                        stg_buf = (stg_buf + 1) % STG_BUFS;
                        if constexpr (LSE) {
                            // running (max, sum exp) of this row over the chunk: ONE rescale per 32 values -- max of the
                            // chunk first (FMNMX3), then exp2((v - max) log2 e) per value: 3.5 instructions per value
                            // where the per-quad online update (five guarded __expf) cost ~13.  Columns >= N (ragged last
                            // tile) do not take part.
                            constexpr float LOG2E = 1.4426950408889634f;
                            const bool ragged = nc + 32 > N;            // warp-uniform
                            float cm = -INFINITY;
                            if (!ragged) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) cm = fmaxf(cm, (nc + j < N) ? __uint_as_float(v[j]) : -INFINITY);
                            }
                            if (cm > -INFINITY) {
                                const float mn = fmaxf(lse_m, cm);
                                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                                if (!ragged) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) s4[j & 3] += ex2_ftz((__uint_as_float(v[j]) - mn) * LOG2E);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        s4[j & 3] += (nc + j < N) ? ex2_ftz((__uint_as_float(v[j]) - mn) * LOG2E) : 0.f;
                                }
                                lse_s = lse_s * ex2_ftz((lse_m - mn) * LOG2E) + ((s4[0] + s4[1]) + (s4[2] + s4[3]));
                                lse_m = mn;
                            }
                        }
                    } else {
                        // general path (fp16 output, residual input, unaligned D): row scale before the transpose,
                        // column scale / clamp / bias / residual after it, coalesced stores from the swizzled tile
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(stg + lane * 32 + (((j >> 2) ^ sw) << 2)) =
                                make_float4(__uint_as_float(v[j]) * rs, __uint_as_float(v[j + 1]) * rs,
                                            __uint_as_float(v[j + 2]) * rs, __uint_as_float(v[j + 3]) * rs);
                        __syncwarp();
                        const bool full = (nc + 32 <= N);
                        const bool d_vec = OUT_HALF ? (((ep.ldd & 7) == 0) && ((reinterpret_cast<uintptr_t>(ep.D) & 15u) == 0))
                                                    : (((ep.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.D) & 15u) == 0));
                        const bool c_vec = ep.C && ((ep.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15u) == 0);
                        if (full && d_vec && (!ep.C || c_vec)) {
                            // 8 lanes cover one 128 B row segment; one instruction writes 4 rows
                            const int rsub = lane >> 3, cq = lane & 7;
                            const float4 cs4 = *reinterpret_cast<const float4*>(epi_cs + cl + cq * 4);
                            const float4 bi4 = *reinterpret_cast<const float4*>(epi_bias + cl + cq * 4);
                            long long doff = static_cast<long long>(rbase + rsub) * ep.ldd + nc + cq * 4;
                            const long long coff = static_cast<long long>(rbase + rsub) * ep.ldc + nc + cq * 4;
                            float4 cv8[8];
                            if (ep.C) {                             // all eight residual loads in flight at once
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    cv8[i] = (rbase + i * 4 + rsub < M) ? *reinterpret_cast<const float4*>(ep.C + coff + static_cast<long long>(i) * 4 * ep.ldc)
                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int rl = i * 4 + rsub;
                                if (rbase + rl < M) {
                                    float4 o = *reinterpret_cast<const float4*>(stg + rl * 32 + ((cq ^ (rl & 7)) << 2));
                                    o.x *= cs4.x; o.y *= cs4.y; o.z *= cs4.z; o.w *= cs4.w;
                                    if (ep.clamp_abs > 0.f) {
                                        o.x = fminf(fmaxf(o.x, -ep.clamp_abs), ep.clamp_abs); o.y = fminf(fmaxf(o.y, -ep.clamp_abs), ep.clamp_abs);
                                        o.z = fminf(fmaxf(o.z, -ep.clamp_abs), ep.clamp_abs); o.w = fminf(fmaxf(o.w, -ep.clamp_abs), ep.clamp_abs);
                                    }
                                    o.x += bi4.x; o.y += bi4.y; o.z += bi4.z; o.w += bi4.w;
                                    if (ep.C) { o.x += cv8[i].x; o.y += cv8[i].y; o.z += cv8[i].z; o.w += cv8[i].w; }
                                    if (ep.act == 1) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
                                    if constexpr (OUT_HALF) {
                                        *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(ep.D) + doff) =
                                            make_uint2(pack_h2(o.x, o.y), pack_h2(o.z, o.w));
                                    } else {
                                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.D) + doff) = o;
                                    }
                                }
                                doff += 4 * ep.ldd;
                            }
                        } else {
                            // unaligned / ragged: one row per instruction, lane = column (still coalesced)
                            const bool col_ok = nc + lane < N;
                            const float cs1 = epi_cs[cl + lane], bi1 = epi_bias[cl + lane];
                            long long doff = static_cast<long long>(rbase) * ep.ldd + nc + lane;
                            long long coff = static_cast<long long>(rbase) * ep.ldc + nc + lane;
#pragma unroll 4
                            for (int rl = 0; rl < 32; ++rl) {
                                if (rbase + rl < M && col_ok) {
                                    float o = stg[rl * 32 + ((((lane >> 2) ^ (rl & 7)) << 2) | (lane & 3))] * cs1;
                                    if (ep.clamp_abs > 0.f) o = fminf(fmaxf(o, -ep.clamp_abs), ep.clamp_abs);
                                    o += bi1;
                                    if (ep.C) o += ep.C[coff];
                                    if (ep.act == 1) o = gelu_erf(o);
                                    if constexpr (OUT_HALF) reinterpret_cast<unsigned short*>(ep.D)[doff] = f2h_sat(o);
                                    else reinterpret_cast<float*>(ep.D)[doff] = o;
                                }
                                doff += ep.ldd;
                                coff += ep.ldc;
                            }
                        }
                        __syncwarp();                               // the tile is reused by the next chunk
                    }
                }
            }
            if constexpr (LSE) {
                if (row < M) ep.lse_part[static_cast<long long>(row) * ep.lse_ld + n_blk * NPART + part] = make_float2(lse_m, lse_s);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), lead_rank));   // the pair leader's barrier
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();           // neither CTA leaves while the other can still signal it / read its smem
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (CTA2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 row-major [rows, cols] with leading dimension ld (elements); box = 64 columns x box_rows.
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SPQ_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld box %d", static_cast<int>(r),
                  static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows);
        return SPQ_ERR_CUDA;
    }
    return SPQ_OK;
}

// e4m3 row-major [rows, cols] bytes with leading dimension ld (bytes); box = 128 columns (one swizzle row) x box_rows.
static int make_tmap_u8(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SPQ_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(2 * BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (e4m3) failed (%d) for [%lld x %lld] ld %lld box %d", static_cast<int>(r),
                  static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows);
        return SPQ_ERR_CUDA;
    }
    return SPQ_OK;
}

// row-major D [M, N], leading dimension ldd (elements): box = 32 columns x 32 rows; fp32 rows of the box are
// 128 B (128B swizzle), fp16 rows 64 B (64B swizzle) -- the swizzle keeps the per-row smem writes conflict-free
static int make_tmap_out(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, bool is_half) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SPQ_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * (is_half ? 2 : 4)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                     gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     is_half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (output) failed (%d)", static_cast<int>(r));
        return SPQ_ERR_CUDA;
    }
    return SPQ_OK;
}

template <int BN, bool OUT_HALF, bool PRE_C = false, bool LSE = false, bool F8 = false>
static int launch_nt(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tA2, const CUtensorMap& tB2,
                     const CUtensorMap& tD, int M, int N, int kb1, int kb2, const EpiParams& ep, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        SPQ_CUDA_OK(cudaFuncSetAttribute(qgemm_nt_kernel<BN, OUT_HALF, PRE_C, LSE, false, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    qgemm_nt_kernel<BN, OUT_HALF, PRE_C, LSE, false, F8><<<grid, NUM_THREADS, L::TOTAL, stream>>>(tA, tB, tA2, tB2, tD, M, N, kb1, kb2, ep);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}


// cluster size of a CTA-pair launch: 2 (one pair) or 4 (two pairs sharing the B tile by multicast).  Set by qgemm_impl
// before it builds the B tensor maps (a quad stages BN / 4 rows of B per CTA and k-block).
static thread_local int g_pair_cluster = 2;

// how many clusters of `csize` CTAs of this kernel can be co-resident (cached per kernel and size)
template <typename K>
static int max_active_clusters(K kern, cudaLaunchConfig_t cfg, int csize, int* cache) {
    if (cache[csize] < 0) {
        int n = 0;
        cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / csize * csize));
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        cache[csize] = n > 0 ? n : 0;
    }
    return cache[csize];
}

// CTA-pair launch: clusters of two (or four) CTAs, 256 x BN tiles per pair, persistent: as many clusters as fit
template <int BN, bool OUT_HALF, bool PRE_C = false, bool LSE = false, bool F8 = false>
static int launch_nt_pair(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tA2, const CUtensorMap& tB2,
                          const CUtensorMap& tD, int M, int N, int kb1, int kb2, const EpiParams& ep, cudaStream_t stream) {
    using L = SmemLayout<BN, true>;
    auto kern = qgemm_nt_kernel<BN, OUT_HALF, PRE_C, LSE, true, F8>;
    static bool attr_set = false;
    if (!attr_set) {
        SPQ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int csize = g_pair_cluster;
    const int rows_per_cluster = BM * csize;
    const int tiles = ((M + rows_per_cluster - 1) / rows_per_cluster) * ((N + BN - 1) / BN);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / csize * csize));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(csize); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent clusters: exactly as many as can be co-resident (a TPC with one SM fused off cannot host a pair)
    static int cache[5] = {-1, -1, -1, -1, -1};
    int max_clusters = max_active_clusters(kern, cfg, csize, cache);
    if (max_clusters < 1) max_clusters = 1;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(static_cast<unsigned>(csize * clusters));
    SPQ_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tA, tB, tA2, tB2, tD, M, N, kb1, kb2, ep));
    spq::count_launch();
    return SPQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Transposed-operand GEMM (weight-gradient shape): D[i,j] += alpha * is[i] * js[j] * sum_m P[m,i] Q[m,j].
// P [Mred, I] and Q [Mred, J] are row-major fp16, i.e. "MN-major" UMMA operands: a TMA box of
// 64 reduction rows x 64 columns lands as 64 rows of 128 B (128B swizzle), consecutive 64-column
// blocks 8 KB apart (LBO), 8-row groups 1 KB apart (SBO).  The reduction (over tokens) is split
// across CTAs; partial tiles are combined with fp32 atomics (RED) into the pre-zeroed D.
constexpr int TN_BOX_BYTES = 64 * 128;   // 64 reduction rows x 64 fp16
constexpr int TN_THREADS = 192;           // TMA warp, MMA warp, 4 epilogue warps (one per TMEM lane quadrant)

template <int BJ>
struct TnSmem {
    static constexpr int P_BYTES = 2 * TN_BOX_BYTES;            // 128 output rows = 2 boxes
    static constexpr int Q_BYTES = (BJ / 64) * TN_BOX_BYTES;
    static constexpr int STAGE_BYTES = P_BYTES + Q_BYTES;
    static constexpr int STAGES = (163840 / STAGE_BYTES) > 8 ? 8 : (163840 / STAGE_BYTES);
    static constexpr int TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 2) * 8 + 16 + 1024;
    static constexpr int TMEM_COLS = BJ < 32 ? 32 : BJ;          // 64 / 128 / 256: powers of two
};

struct TnEpi {
    float* part;                   // [splits][I * J] partial results, each in D's own layout
    long long stride_i, stride_j;  // D (and every partial) is addressed [i * stride_i + j * stride_j]
    long long plane;               // I * J
};

template <int BJ>
__global__ void __launch_bounds__(TN_THREADS, 1)
qgemm_tn_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, int I, int J, int kb_total,
                int kb_per_split, int j_tiles, TnEpi ep) {
    using L = TnSmem<BJ>;
    extern __shared__ uint8_t smem_raw[];
    // 1024 B alignment for the 128B swizzle.  Offset arithmetic on the __shared__ symbol (not an integer
    // round trip) so that the compiler keeps the shared address space and emits LDS/STS, not generic LD/ST.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::STAGES * L::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + L::STAGES;
    uint64_t* tmem_full = empty_bar + L::STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = (blockIdx.x / j_tiles) * BM;
    const int j0 = (blockIdx.x % j_tiles) * BJ;
    const int kb_begin = blockIdx.y * kb_per_split;
    int kb_end = kb_begin + kb_per_split;
    if (kb_end > kb_total) kb_end = kb_total;
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmQ);
        for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "n"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sp = smem + stage * L::STAGE_BYTES;
                uint8_t* sq = sp + L::P_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                tma_load_2d(sp, &tmP, &full_bar[stage], i0, kb * 64);
                tma_load_2d(sp + TN_BOX_BYTES, &tmP, &full_bar[stage], i0 + 64, kb * 64);
#pragma unroll
                for (int b = 0; b < BJ / 64; ++b) tma_load_2d(sq + b * TN_BOX_BYTES, &tmQ, &full_bar[stage], j0 + 64 * b, kb * 64);
                if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BJ, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sp = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint64_t pdesc = make_desc_sw128(sp, TN_BOX_BYTES, 1024);
                const uint64_t qdesc = make_desc_sw128(sp + L::P_BYTES, TN_BOX_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // 16 reduction rows = 2048 B per UMMA
                    umma_f16(tmem_base, pdesc + static_cast<uint64_t>(128 * k), qdesc + static_cast<uint64_t>(128 * k), idesc,
                             (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (kb == nkb - 1) umma_commit(&tmem_full[0]);
                if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (nkb > 0) {
        // epilogue: this CTA's partial tile goes to its own plane of the workspace (no atomics: the planes are
        // folded in a fixed order by tn_reduce_kernel, so the result is bitwise reproducible)
        const int quad = warp & 3;
        mbar_wait(&tmem_full[0], 0);
        tcgen05_fence_after();
        const int i = i0 + quad * 32 + lane;
        const bool i_ok = i < I;
        float* base = ep.part + static_cast<long long>(blockIdx.y) * ep.plane + static_cast<long long>(i) * ep.stride_i;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const bool row_vec = ep.stride_j == 1 && (ep.stride_i & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.part) & 15u) == 0 &&
                             (ep.plane & 3) == 0;
#pragma unroll 1
        for (int c = 0; c < BJ / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            const int jb = j0 + c * 32;
            if (!i_ok || jb >= J) continue;
            if (row_vec && jb + 32 <= J) {
                // D row-major: the lane owns 128 contiguous bytes of its row
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4)
                    *reinterpret_cast<float4*>(base + jb + jj) = make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                                                                             __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
            } else {
                // D stored transposed (stride_i == 1): for each j the 32 lanes write 32 consecutive floats
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    if (jb + jj < J) base[static_cast<long long>(jb + jj) * ep.stride_j] = __uint_as_float(v[jj]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::TMEM_COLS) : "memory");
    }
}

struct TnReduce {
    const float* part;
    int splits;
    long long plane, I, J, stride_i, stride_j;
    const float* i_scale;
    const float* j_scale;
    const float* alpha_dev;
    float alpha, clamp_abs;
    const float* gq_scale_i;     // optional: min-max fake quantisation of the result, scale per row i (GradientQuantizer)
    float gq_levels;             // 2^(b-1) - 1
    int accumulate;              // D += result (gradient accumulation straight into a .grad buffer)
    float* D;
};

// p2/quantization.py:14-26 GradientQuantizer.backward on the finished LoRA gradient: symmetric min-max fake quantisation
// with a per-row scale (true division, round-half-even, clamp to +-levels), then the weight quantiser's STE clamp
__device__ __forceinline__ float tn_finish(const TnReduce& a, float v, long long i) {
    if (a.gq_scale_i) {
        const float sc = __ldg(a.gq_scale_i + i);
        v = fminf(fmaxf(rintf(__fdiv_rn(v, sc)), -a.gq_levels), a.gq_levels) * sc;
    }
    if (a.clamp_abs > 0.f) v = fminf(fmaxf(v, -a.clamp_abs), a.clamp_abs);
    return v;
}

// D[e] = clamp(alpha * alpha_dev * is[i] * js[j] * sum_s part[s][e]): the splits are added in index order.
__global__ void __launch_bounds__(64) tn_reduce_kernel(TnReduce a) {
    const float al = a.alpha * (a.alpha_dev ? __ldg(a.alpha_dev) : 1.0f);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const bool vec = (a.plane & 3) == 0 && (reinterpret_cast<uintptr_t>(a.part) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.D) & 15u) == 0 &&
                     (((a.stride_j == 1) ? a.J : a.I) & 3) == 0;
    if (vec) {
        for (long long e4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e4 < (a.plane >> 2); e4 += stride) {
            const long long e = e4 << 2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int s = 0;
            for (; s + 8 <= a.splits; s += 8) {           // eight independent loads in flight, added in index order
                float4 p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) p[u] = *reinterpret_cast<const float4*>(a.part + static_cast<long long>(s + u) * a.plane + e);
#pragma unroll
                for (int u = 0; u < 8; ++u) { acc.x += p[u].x; acc.y += p[u].y; acc.z += p[u].z; acc.w += p[u].w; }
            }
            for (; s < a.splits; ++s) {
                const float4 p = *reinterpret_cast<const float4*>(a.part + static_cast<long long>(s) * a.plane + e);
                acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            }
            float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                long long i, j;
                if (a.stride_j == 1) { i = (e + u) / a.J; j = (e + u) - i * a.J; }
                else { j = (e + u) / a.I; i = (e + u) - j * a.I; }
                o[u] = tn_finish(a, o[u] * al * (a.i_scale ? __ldg(a.i_scale + i) : 1.0f) * (a.j_scale ? __ldg(a.j_scale + j) : 1.0f), i);
            }
            if (a.accumulate) {
                const float4 d = *reinterpret_cast<const float4*>(a.D + e);
                o[0] += d.x; o[1] += d.y; o[2] += d.z; o[3] += d.w;
            }
            *reinterpret_cast<float4*>(a.D + e) = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < a.plane; e += stride) {
            float acc = a.part[e];
            for (int s = 1; s < a.splits; ++s) acc += a.part[static_cast<long long>(s) * a.plane + e];
            long long i, j;
            if (a.stride_j == 1) { i = e / a.J; j = e - i * a.J; }
            else { j = e / a.I; i = e - j * a.I; }
            const float v = tn_finish(a, acc * al * (a.i_scale ? __ldg(a.i_scale + i) : 1.0f) * (a.j_scale ? __ldg(a.j_scale + j) : 1.0f), i);
            a.D[e] = a.accumulate ? a.D[e] + v : v;
        }
    }
}

// reduction splits: enough CTAs for two waves, never more than the k-blocks there are
static void tn_splits(int64_t Mred, int64_t I, int64_t J, int bj, int& splits, int& per, int& j_tiles, int& tiles) {
    const int i_tiles = static_cast<int>((I + BM - 1) / BM);
    j_tiles = static_cast<int>((J + bj - 1) / bj);
    tiles = i_tiles * j_tiles;
    const int kb_total = static_cast<int>((Mred + 63) / 64);
    const int sms = sm_count() > 0 ? sm_count() : 148;
    splits = (sms + tiles - 1) / tiles;               // one wave of CTAs: every extra split is another plane to fold
    if (splits > kb_total) splits = kb_total;
    if (splits < 1) splits = 1;
    per = (kb_total + splits - 1) / splits;
    splits = (kb_total + per - 1) / per;
}
static int tn_bj(int64_t J) { return J <= 64 ? 64 : J <= 128 ? 128 : 256; }

template <int BJ>
static int launch_tn(const CUtensorMap& tP, const CUtensorMap& tQ, int I, int J, int kb_total, int splits, int per, int j_tiles,
                     int tiles, const TnEpi& ep, cudaStream_t stream) {
    using L = TnSmem<BJ>;
    static bool attr_set = false;
    if (!attr_set) {
        SPQ_CUDA_OK(cudaFuncSetAttribute(qgemm_tn_kernel<BJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    dim3 grid(tiles, splits);
    qgemm_tn_kernel<BJ><<<grid, TN_THREADS, L::TOTAL, stream>>>(tP, tQ, I, J, kb_total, per, j_tiles, ep);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}

// fp16 row-major [rows, cols]: box = 64 columns x 64 rows (MN-major operand boxes)
static int make_tmap_mn(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld) {
    return make_tmap(tm, base, rows, cols, ld, 64);
}

}  // namespace gemm
}  // namespace spq

using namespace spq;
using namespace spq::gemm;

extern "C" int spq_debug_status(int* aborted_host) {
    int v = 0;
    SPQ_CUDA_OK(cudaDeviceSynchronize());
    SPQ_CUDA_OK(cudaMemcpyFromSymbol(&v, g_abort, sizeof(int)));
    if (aborted_host) *aborted_host = v;
    if (v) {
        int zero = 0;
        SPQ_CUDA_OK(cudaMemcpyToSymbol(g_abort, &zero, sizeof(int)));
    }
    return SPQ_OK;
}

// tile width: the widest BN that still gives every SM a tile; otherwise the one with most tiles
static int pick_bn(int64_t M, int64_t N, int sms) {
    const long long mt = (M + BM - 1) / BM;
    auto tiles = [&](int b) { return mt * ((N + b - 1) / b); };
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (tiles(256) >= sms) return 256;
    if (tiles(128) >= sms) return 128;
    return 64;
}

static int qgemm_impl(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N,
                      int64_t K, const spq_half_t* A2, int64_t lda2, const spq_half_t* B2, int64_t ldb2, int64_t K2,
                      float alpha, const float* row_scale, const float* col_scale, const float* bias, float clamp_abs,
                      const float* C, int64_t ldc, void* D, int64_t ldd, int d_is_half, int activation,
                      float* lse_part, int64_t lse_ld, spq_stream_t stream, bool f8 = false) {
    SPQ_REQUIRE(A && B && D, "spq_qgemm: null operand");
    SPQ_REQUIRE(M > 0 && N > 0 && K > 0, "spq_qgemm: empty problem %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
    SPQ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "spq_qgemm: dimension overflow");
    SPQ_REQUIRE((lda % (f8 ? 16 : 8)) == 0 && (ldb % (f8 ? 16 : 8)) == 0 && lda >= K && ldb >= K,
                "spq_qgemm: lda/ldb must be multiples of 16 bytes and >= K");
    SPQ_REQUIRE(aligned16(A) && aligned16(B), "spq_qgemm: operands must be 16-byte aligned");
    SPQ_REQUIRE(K2 == 0 || (A2 && B2 && (lda2 % 8) == 0 && (ldb2 % 8) == 0 && aligned16(A2) && aligned16(B2)),
                "spq_qgemm: bad second K segment");
    SPQ_REQUIRE(ldd >= N && (!C || ldc >= N), "spq_qgemm: ldd/ldc < N");

    const int sms = sm_count();
    if (sms <= 0) {
        set_error("spq_qgemm: no CUDA device");
        return SPQ_ERR_CUDA;
    }
    const int bn = pick_bn(M, N, sms);
    // CTA-pair mode (cta_group::2, 256 x 256 tiles): wide outputs with enough 256-row tiles for every SM pair
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("SPQ_GEMM_PAIR"); pair_env = e ? atoi(e) : 1; }      // SPQ_GEMM_PAIR=0: A/B switch
    const bool pair = pair_env != 0 && bn == 256 &&
                      ((M + 2 * BM - 1) / (2 * BM)) * ((N + bn - 1) / bn) >= sms / 2;
    // Two pairs per cluster with the B tile multicast (SPQ_GEMM_CLUSTER4=1).  OFF by default: measured on the same box it
    // is ~10 % slower (c_attn 114.0 vs 106.7 us, c_fc + GELU 197 vs 178 us, main loop alone 123 vs 110 us).  Multicast
    // halves the B bytes REQUESTED from L2 but every SM still RECEIVES its half tile, and quads fit 128 of the 148 SMs:
    // the kernel is bound by the bytes that enter and leave each SM (~43 B/clk/SM, the rate cuBLAS also stops at), not
    // by L2 slice bandwidth.  Kept because it is bit-identical (tests/test_gpu_gemm_cluster.py) and documents the bound.
    static int quad_env = -2;
    if (quad_env == -2) { const char* e = getenv("SPQ_GEMM_CLUSTER4"); quad_env = e ? atoi(e) : 0; }
    const bool quad = pair && quad_env > 0;
    g_pair_cluster = quad ? 4 : 2;
    const int b_rows = pair ? (quad ? bn / 4 : bn / 2) : bn;      // rows of B each CTA fetches per k-block
    CUtensorMap tA, tB, tA2, tB2;
    int rc;
    if (f8) {
        if ((rc = make_tmap_u8(&tA, A, M, K, lda, BM)) != SPQ_OK) return rc;
        if ((rc = make_tmap_u8(&tB, B, N, K, ldb, b_rows)) != SPQ_OK) return rc;
    } else {
        if ((rc = make_tmap(&tA, A, M, K, lda, BM)) != SPQ_OK) return rc;
        if ((rc = make_tmap(&tB, B, N, K, ldb, b_rows)) != SPQ_OK) return rc;
    }
    if (K2 > 0) {
        if ((rc = make_tmap(&tA2, A2, M, K2, lda2, BM)) != SPQ_OK) return rc;
        if ((rc = make_tmap(&tB2, B2, N, K2, ldb2, b_rows)) != SPQ_OK) return rc;
    } else {
        tA2 = tA;
        tB2 = tB;
    }
    EpiParams ep;
    ep.row_scale = row_scale; ep.col_scale = col_scale; ep.bias = bias; ep.C = C; ep.alpha_dev = nullptr;
    ep.D = D; ep.ldc = ldc; ep.ldd = ldd; ep.d_stride_n = 1; ep.alpha = alpha; ep.clamp_abs = clamp_abs;
    ep.act = activation;
    {
        // SPQ_GEMM_STORE_HINT: 0 never, 1 always, unset: outputs of at least 1 GB (the LM head's logits)
        static int hint_env = -2;
        if (hint_env == -2) { const char* e = getenv("SPQ_GEMM_STORE_HINT"); hint_env = e ? atoi(e) : -1; }
        const double out_bytes = static_cast<double>(M) * static_cast<double>(N) * (d_is_half ? 2.0 : 4.0);
        ep.store_hint = hint_env >= 0 ? (hint_env != 0) : (out_bytes >= 1073741824.0);
    }
    ep.lse_part = reinterpret_cast<float2*>(lse_part); ep.lse_ld = lse_ld;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("SPQ_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
        ep.debug = dbg;
    }
    const int kb1 = static_cast<int>(f8 ? (K + 2 * BK - 1) / (2 * BK) : (K + BK - 1) / BK);
    const int kb2 = static_cast<int>((K2 + BK - 1) / BK);
    cudaStream_t st = as_stream(stream);
    const int m = static_cast<int>(M), n = static_cast<int>(N);
    // TMA store path: output rows 16-byte aligned; a residual C must be float4-addressable
    CUtensorMap tD = tA;
    // (TMA clips the inner dimension at 16-byte granularity: with N % 4 != 0 (fp16: N % 8) the 1-3 floats (1-7
    // halves) of row padding after column N are overwritten too, so the row must have that padding)
    const int64_t gran = d_is_half ? 8 : 4;
    const bool c_ok = !C || (!d_is_half && (ldc % 4) == 0 && (N % 4) == 0 && aligned16(C));   // fp16 D + residual: general path
    ep.tma_store = (c_ok && (ldd % gran) == 0 && ldd >= (N + gran - 1) / gran * gran && aligned16(D) && !(ep.debug & 4)) ? 1 : 0;
    if (ep.tma_store && (rc = make_tmap_out(&tD, D, M, N, ldd, d_is_half != 0)) != SPQ_OK) return rc;
    if (lse_part) {
        SPQ_REQUIRE(ep.tma_store && !d_is_half && !C, "spq_qgemm_lse: needs float32 output with 16-byte aligned, padded rows and no residual");
        SPQ_REQUIRE(lse_ld >= NPART * ((N + bn - 1) / bn) && (reinterpret_cast<uintptr_t>(lse_part) & 7u) == 0, "spq_qgemm_lse: partials buffer too narrow");
        if (pair) return launch_nt_pair<256, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (bn == 256) return launch_nt<256, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (bn == 128) return launch_nt<128, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        return launch_nt<64, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    }
    if (f8) {
        // e4m3 code GEMM (eval configurations: <= 4-bit, per-tensor input scale): float32 or fp16 output, no residual
        SPQ_REQUIRE(!C && !lse_part, "spq_qgemm_f8: no residual / log-sum-exp epilogue");
        if (pair) {
            if (d_is_half) return launch_nt_pair<256, true, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
            return launch_nt_pair<256, false, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        }
        if (d_is_half) {
            if (bn == 256) return launch_nt<256, true, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
            if (bn == 128) return launch_nt<128, true, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
            return launch_nt<64, true, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        }
        if (bn == 256) return launch_nt<256, false, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (bn == 128) return launch_nt<128, false, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        return launch_nt<64, false, false, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    }
    if (pair) {
        if (d_is_half) return launch_nt_pair<256, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (ep.tma_store && C) return launch_nt_pair<256, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        return launch_nt_pair<256, false>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    }
    if (d_is_half) {
        if (bn == 256) return launch_nt<256, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (bn == 128) return launch_nt<128, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        return launch_nt<64, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    }
    if (ep.tma_store && C) {
        if (bn == 256) return launch_nt<256, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        if (bn == 128) return launch_nt<128, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
        return launch_nt<64, false, true>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    }
    if (bn == 256) return launch_nt<256, false>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    if (bn == 128) return launch_nt<128, false>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
    return launch_nt<64, false>(tA, tB, tA2, tB2, tD, m, n, kb1, kb2, ep, st);
}

extern "C" int spq_qgemm(const spq_half_t* A, int64_t lda, const spq_half_t* B, int64_t ldb, int64_t M, int64_t N,
                         int64_t K, const spq_half_t* A2, int64_t lda2, const spq_half_t* B2, int64_t ldb2, int64_t K2,
                         float alpha, const float* row_scale, const float* col_scale, const float* bias, float clamp_abs,
                         const float* C, int64_t ldc, void* D, int64_t ldd, int d_is_half, int activation,
                         spq_stream_t stream) {
    return qgemm_impl(A, lda, B, ldb, M, N, K, A2, lda2, B2, ldb2, K2, alpha, row_scale, col_scale, bias, clamp_abs, C, ldc,
                      D, ldd, d_is_half, activation, nullptr, 0, stream);
}

extern "C" int spq_qgemm_f8(const uint8_t* A, int64_t lda, const uint8_t* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                            const spq_half_t* A2, int64_t lda2, const spq_half_t* B2, int64_t ldb2, int64_t K2, float alpha,
                            const float* row_scale, const float* col_scale, const float* bias, void* D, int64_t ldd,
                            int d_is_half, int activation, spq_stream_t stream) {
    return qgemm_impl(A, lda, B, ldb, M, N, K, A2, lda2, B2, ldb2, K2, alpha, row_scale, col_scale, bias, 0.f, nullptr, 0, D, ldd,
                      d_is_half, activation, nullptr, 0, stream, true);
}

extern "C" int64_t spq_qgemm_lse_parts(int64_t M, int64_t N) {
    const int sms = sm_count();
    if (M <= 0 || N <= 0 || sms <= 0) return 0;
    const int bn = pick_bn(M, N, sms);
    return NPART * ((N + bn - 1) / bn);
}

extern "C" int spq_qgemm_lse(const spq_half_t* A, int64_t lda, const spq_half_t* B, int64_t ldb, int64_t M, int64_t N,
                             int64_t K, float alpha, const float* row_scale, const float* col_scale, const float* bias,
                             float* D, int64_t ldd, float* lse_part, int64_t lse_ld, spq_stream_t stream) {
    SPQ_REQUIRE(lse_part, "spq_qgemm_lse: null partials buffer");
    return qgemm_impl(A, lda, B, ldb, M, N, K, nullptr, 0, nullptr, 0, 0, alpha, row_scale, col_scale, bias, 0.f, nullptr, 0,
                      D, ldd, 0, 0, lse_part, lse_ld, stream);
}

extern "C" size_t spq_gemm_tn_workspace_bytes(int64_t Mred, int64_t I, int64_t J) {
    if (Mred <= 0 || I <= 0 || J <= 0) return 0;
    int splits, per, j_tiles, tiles;
    tn_splits(Mred, I, J, tn_bj(J), splits, per, j_tiles, tiles);
    return static_cast<size_t>(splits) * static_cast<size_t>(I) * static_cast<size_t>(J) * sizeof(float);
}

extern "C" int spq_gemm_tn(const spq_half_t* P, int64_t ldp, const spq_half_t* Q, int64_t ldq, int64_t Mred, int64_t I, int64_t J,
                           float alpha, const float* alpha_dev, const float* i_scale, const float* j_scale, float clamp_abs,
                           const float* gq_scale_i, int gq_bits, int accumulate,
                           float* D, int64_t d_stride_i, int64_t d_stride_j, void* workspace, size_t workspace_bytes,
                           spq_stream_t stream) {
    SPQ_REQUIRE(!gq_scale_i || (gq_bits >= 2 && gq_bits <= 24), "spq_gemm_tn: gradient quantiser bits %d", gq_bits);
    SPQ_REQUIRE(P && Q && D && workspace, "spq_gemm_tn: null operand");
    SPQ_REQUIRE(Mred > 0 && I > 0 && J > 0 && Mred < (1ll << 31) && I < (1ll << 31) && J < (1ll << 31), "spq_gemm_tn: bad shape");
    SPQ_REQUIRE((ldp % 8) == 0 && (ldq % 8) == 0 && ldp >= I && ldq >= J && aligned16(P) && aligned16(Q),
                "spq_gemm_tn: leading dimensions must be multiples of 8 and operands 16-byte aligned");
    SPQ_REQUIRE((d_stride_i == 1 && d_stride_j == I) || (d_stride_j == 1 && d_stride_i == J),
                "spq_gemm_tn: D must be dense, [I, J] row-major or its transpose");
    SPQ_REQUIRE(workspace_bytes >= spq_gemm_tn_workspace_bytes(Mred, I, J) && aligned16(workspace), "spq_gemm_tn: workspace too small");
    if (sm_count() <= 0) {
        set_error("spq_gemm_tn: no CUDA device");
        return SPQ_ERR_CUDA;
    }
    cudaStream_t st = as_stream(stream);
    CUtensorMap tP, tQ;
    int rc;
    if ((rc = make_tmap_mn(&tP, P, Mred, I, ldp)) != SPQ_OK) return rc;
    if ((rc = make_tmap_mn(&tQ, Q, Mred, J, ldq)) != SPQ_OK) return rc;
    const int bj = tn_bj(J);
    int splits, per, j_tiles, tiles;
    tn_splits(Mred, I, J, bj, splits, per, j_tiles, tiles);
    TnEpi ep;
    ep.part = static_cast<float*>(workspace); ep.stride_i = d_stride_i; ep.stride_j = d_stride_j; ep.plane = I * J;
    const int kb_total = static_cast<int>((Mred + 63) / 64);
    const int i = static_cast<int>(I), j = static_cast<int>(J);
    if (bj == 64) rc = launch_tn<64>(tP, tQ, i, j, kb_total, splits, per, j_tiles, tiles, ep, st);
    else if (bj == 128) rc = launch_tn<128>(tP, tQ, i, j, kb_total, splits, per, j_tiles, tiles, ep, st);
    else rc = launch_tn<256>(tP, tQ, i, j, kb_total, splits, per, j_tiles, tiles, ep, st);
    if (rc != SPQ_OK) return rc;
    TnReduce ra;
    ra.part = ep.part; ra.splits = splits; ra.plane = ep.plane; ra.I = I; ra.J = J; ra.stride_i = d_stride_i; ra.stride_j = d_stride_j;
    ra.i_scale = i_scale; ra.j_scale = j_scale; ra.alpha_dev = alpha_dev; ra.alpha = alpha; ra.clamp_abs = clamp_abs; ra.D = D;
    ra.gq_scale_i = gq_scale_i; ra.gq_levels = gq_scale_i ? static_cast<float>((1 << (gq_bits - 1)) - 1) : 0.f;
    ra.accumulate = accumulate ? 1 : 0;
    long long blocks = (ep.plane / 4 + 63) / 64;
    const long long cap = static_cast<long long>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    tn_reduce_kernel<<<static_cast<unsigned>(blocks), 64, 0, st>>>(ra);
    SPQ_LAUNCH_OK();
    return SPQ_OK;
}
