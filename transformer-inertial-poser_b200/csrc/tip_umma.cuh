// tcgen05 GEMM engine of the TIP hot path (sm_100a): C[M,N] = A[M,K] * W[N,K]^T with fp32-parity
// accuracy from three FP16 tensor-core products per tile (error-compensated split,
// a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with hi = fp16(s*x), lo = fp16(s*x - hi), s a power of
// two keeping the operands in fp16's normal range; 22 significant bits like a 3xTF32 split, but at
// the f16 MMA rate and half the operand bytes.  A single TF32/BF16 pass misses the 1e-4 parity bar
// by 100x).  Accumulation is fp32 in TMEM; the epilogue un-scales exactly.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D tiles (128B-swizzled, K-major) of the four
//                                operand planes into a multi-stage shared-memory ring (mbarrier
//                                complete_tx)
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16),
//                                3 per k-step, accumulating in TMEM; tcgen05.commit frees the smem
//                                stage / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld the fp32 accumulator (double-buffered in TMEM so the next
//                                tile's MMAs overlap), fused bias / ReLU / dropout, or residual +
//                                LayerNorm over the full 256-wide row, optional FP16 hi/lo split of
//                                the output for the next GEMM, coalesced-by-row global stores.
#pragma once
#include <cuda.h>

#include "tip_common.cuh"
#include "tip_simt.cuh"

namespace tip {

enum UmmaGemmId { UG_IN = 0, UG_QKV, UG_OUT, UG_FF1, UG_FF2, UG_IH, UG_HEAD_R, UG_HEAD_E, UG_COUNT };
constexpr bool UMMA_AVAILABLE = true;

constexpr int UM_BM = 128;          // UMMA M (cta_group::1)
constexpr int UM_BK = 64;           // fp16 elements per k-block = one 128-byte swizzle row

// CG = CTAs per tile: 1 (M = 128, tcgen05 cta_group::1) or 2 (a CTA pair computes a 256 x BN tile with
// cta_group::2 MMAs: each CTA stages its own 128 rows of A and HALF of the B tile, so the L2->SM operand
// traffic per flop is half that of two independent CTAs).
// BK = fp16 elements per k-block: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B: half-size stages, so
// a 128 x 256 tile gets a FOUR-deep ring in the same shared memory its 64-wide k-blocks only allow two stages of).
template <int BN, int CG = 1, int BK = UM_BK> struct UmmaCfg {
    static constexpr int B_ROWS = BN / CG;                      // rows of the B tile staged by one CTA
    static constexpr int STAGES = (BK == 32) ? 4 : (B_ROWS == 256) ? 2 : (B_ROWS == 64) ? 4 : 3;
    static constexpr int A_BYTES = UM_BM * BK * 2;              // one plane of A per stage
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;   // double-buffered accumulator (power of two)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 8 * 4096 /*epilogue staging*/ + 1024 /*align slack*/ +
                                      320 /*barriers, seed slot, tile-scheduler ring*/ + 1024 /*LN row stats*/;
};

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug traps (surfaces as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are counted on an mbarrier that may live in the PEER CTA of the pair
// (`bar_cluster` is a shared::cluster address, e.g. from mapa)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier at the same shared offset in BOTH CTAs of the pair once the MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t gets row (lane base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// same load without the wait: the caller issues several and then one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
          "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
          "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
}  // namespace ptx

// K-major, 128B-swizzled shared-memory operand descriptor (tile rows x 128 bytes, 8-row groups
// 1024 bytes apart).  Bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// the same for 64-byte rows (BK = 32): SWIZZLE_64B (layout type 4), 8-row groups 512 bytes apart
template <int BK> __device__ __forceinline__ uint64_t umma_smem_desc_bk(uint32_t saddr) {
    if constexpr (BK == 64) return umma_smem_desc(saddr);
    else return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
                ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A/B FP16 (format 0 at bits 7-9 / 10-12),
// both K-major, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- epilogue helpers ----------------------------------------------------------------------------
// Each epilogue warp owns a 32-row x 32-column fp32 staging tile in shared memory (4 KB, 16-byte
// chunks XOR-swizzled by row so both access patterns below are bank-conflict free):
//   "row" pattern   : thread t <-> row t, all 8 chunks (how tcgen05.ld delivers the accumulator)
//   "coal" pattern  : lane l <-> chunk l%8 of rows i*4 + l/8, i = 0..7 (4 full 128-byte lines per
//                     warp instruction in global memory)
__device__ __forceinline__ float4* stg_ptr(float* stg, int row, int chunk) {
    return reinterpret_cast<float4*>(stg) + row * 8 + (chunk ^ (row & 7));
}
__device__ __forceinline__ void stg_write_row(float* stg, int lane, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) *stg_ptr(stg, lane, j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stg_read_row(const float* stg, int lane, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = *stg_ptr(const_cast<float*>(stg), lane, j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
// Lean FP16 hi/lo split of four (already scaled) values: Veltkamp splitting in fp32 (hi = x rounded to
// 11 significant bits, exact in fp16's normal range; lo = x - hi, exact) + two packed conversions
// per plane, instead of three scalar F2F conversions per value.  Saturates at +-65504.
__device__ __forceinline__ void veltkamp11(float x, float& hi, float& lo) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    const float t = __fmul_rn(x, 8193.f);
    hi = __fsub_rn(t, __fsub_rn(t, x));
    lo = __fsub_rn(x, hi);
}
__device__ __forceinline__ void half_split_store4_fast(__half* hi_p, __half* lo_p, float4 x) {
    float h0, h1, h2, h3, l0, l1, l2, l3;
    veltkamp11(x.x, h0, l0); veltkamp11(x.y, h1, l1); veltkamp11(x.z, h2, l2); veltkamp11(x.w, h3, l3);
    const __half2 ha = __floats2half2_rn(h0, h1), hb = __floats2half2_rn(h2, h3);
    const __half2 la = __floats2half2_rn(l0, l1), lb = __floats2half2_rn(l2, l3);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&ha); uh.y = *reinterpret_cast<const uint32_t*>(&hb);
    ul.x = *reinterpret_cast<const uint32_t*>(&la); ul.y = *reinterpret_cast<const uint32_t*>(&lb);
    *reinterpret_cast<uint2*>(hi_p) = uh;
    *reinterpret_cast<uint2*>(lo_p) = ul;
}

constexpr int UM_EPI_WARPS = 8;                 // two warps per TMEM lane quarter (column halves)
constexpr int UM_THREADS = 64 + 32 * UM_EPI_WARPS;

// DROP (LayerNorm variant only): dropout1 / dropout2 compiled in.  The LayerNorm epilogue is issue-bound, so the
// deterministic kernel carries none of the mask arithmetic (the non-LN kernels take a warp-uniform branch instead).
template <int BN, bool LN, bool OUT_HALF, int CG = 1, bool DROP = false, int BK = UM_BK>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                 const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                 const __grid_constant__ CUtensorMap mapC0, const __grid_constant__ CUtensorMap mapC1,
                 const __grid_constant__ CUtensorMap mapR_hi, const __grid_constant__ CUtensorMap mapR_lo,
                 int M, int N, int K, int m_tile0, int m_tile_cnt, Epi ep) {
    const bool pdl_early = ep.pdl_early != 0;
    using Cfg = UmmaCfg<BN, CG, BK>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(BK == 64 || (BK == 32 && !LN && CG == 1), "64-byte k-blocks: plain single-CTA tiles only");
    static_assert(CG == 1 || (CG == 2 && (BN == 256 || (BN == 128 && !LN))), "pair tiles: 256 x 256 (plain or LayerNorm) or 256 x 128 (plain)");
    // 128B-swizzled operand tiles need 1024-byte alignment.  The kernel has no static shared memory,
    // so the dynamic window starts at shared offset 0; keeping the pointer un-cast preserves the
    // shared address space (LDS/STS instead of generic LD/ST for the epilogue staging).
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);          // 8 x 4 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + UM_EPI_WARPS * 4096);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint64_t* rfull_bar = bars + 2 * STAGES + 5;  //           LN: residual tile landed in the ring (TMA -> epilogue)
    // dynamic tile scheduler (ep.sched != null; plain single-CTA tiles): the producer thread draws tile indices from a global
    // counter and hands them to the MMA thread and the epilogue warps through a 4-deep ring.  With several execution lanes
    // a kernel often starts on the SMs another lane's narrow kernel leaves free and gets the rest later: CTAs that are
    // resident early then simply take more tiles (a static round-robin makes the kernel wait for its last CTA's share).
    uint64_t* sfull_bar = bars + 2 * STAGES + 8;      // [4] producer -> consumers
    uint64_t* sempty_bar = bars + 2 * STAGES + 12;    // [4] consumers (MMA thread + 8 epilogue warps) -> producer
    volatile int* stile = reinterpret_cast<volatile int*>(bars + 2 * STAGES + 16);   // [4]
    float* row_stat = reinterpret_cast<float*>(bars + 2 * STAGES + 20);   // (16-byte aligned: float4 reads)   // LN: [2 halves][128 rows] partials, then mean/rstd; plain: [2][256] bias slices
    // LN fast path (no dropout): after a tile's last k-block the operand ring is idle, so the producer parks the
    // residual tile there (128 KB, the same 64-column swizzled boxes the GEMMs read xa / xb with) and the
    // epilogue stages its output boxes in the ring's last 64 KB; both stages go back to the producer when the
    // tile's epilogue is done.
    const bool ln_ring = LN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (N + BN - 1) / BN;
    const int total_tiles = n_tiles * (m_tile_cnt / CG);       // m-tiles [m_tile0, m_tile0 + m_tile_cnt), CG per tile
    const int num_kb = K / BK;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;   // rank 0 = leader: issues the pair's MMAs
    const int grp = blockIdx.x / CG, n_grp = gridDim.x / CG;   // persistent loop over tiles, one CTA group per tile
    const bool dyn = (CG == 1 && !LN) && ep.sched != nullptr;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mapA_hi); ptx::prefetch_tmap(&mapA_lo);
        ptx::prefetch_tmap(&mapB_hi); ptx::prefetch_tmap(&mapB_lo);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        // the accumulator of a pair is released by the epilogue warps of BOTH CTAs (on the leader's barrier)
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], UM_EPI_WARPS * CG); }
        ptx::mbar_init(rfull_bar, 1);
        for (int s = 0; s < 4; ++s) { ptx::mbar_init(&sfull_bar[s], 1); ptx::mbar_init(&sempty_bar[s], 1 + UM_EPI_WARPS); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { if constexpr (CG == 2) ptx::tmem_alloc2(tmem_slot, Cfg::TMEM_COLS); else ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS); }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) { cluster_arrive(); cluster_wait(); }   // the peer's barriers exist before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();          // PDL: everything above overlapped the previous kernel; its results are needed from here on
    if (pdl_early) griddep_launch();
#define TIP_TS(i) do { if (ep.tbuf && blockIdx.x == 0 && lane == 0) ep.tbuf[i] = ptx::globaltimer_ns(); } while (0)
    if (warp == 2) TIP_TS(0);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t uses[STAGES];                 // fills of each stage so far (k-blocks and residual tiles)
#pragma unroll
            for (int s = 0; s < STAGES; ++s) uses[s] = 0;
            if constexpr (LN) { ptx::prefetch_tmap(&mapR_hi); ptx::prefetch_tmap(&mapR_lo); }
            int tile = grp;
            if (dyn) {                                  // first tile: drawn and published (ring slot 0 is free)
                tile = atomicAdd(ep.sched, 1);
                stile[0] = tile;
                ptx::mbar_arrive(&sfull_bar[0]);
            }
            for (int pit = 0; tile < total_tiles; ++pit) {
                int tile_next = tile + n_grp;
                if (dyn) tile_next = atomicAdd(ep.sched, 1);       // in flight while this tile's loads are issued
                const int m0 = (m_tile0 + (tile / n_tiles) * CG + (int)cta_rank) * UM_BM;
                const int n0 = (tile % n_tiles) * BN + (int)cta_rank * Cfg::B_ROWS;      // this CTA's slice of the B tile
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], (uses[stage] & 1u) ^ 1u);
                    uses[stage]++;
                    uint8_t* s = smem + stage * Cfg::STAGE_BYTES;
                    if constexpr (CG == 2) {
                        // both CTAs' bytes are counted on the LEADER's full barrier (it expects 2 stages' worth)
                        const uint32_t fb = map_to_cta(ptx::smem_u32(&full_bar[stage]), 0u);
                        if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        ptx::tma_load_2d_pair(s, &mapA_hi, fb, kb * BK, m0);
                        ptx::tma_load_2d_pair(s + Cfg::A_BYTES, &mapA_lo, fb, kb * BK, m0);
                        ptx::tma_load_2d_pair(s + 2 * Cfg::A_BYTES, &mapB_hi, fb, kb * BK, n0);
                        ptx::tma_load_2d_pair(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapB_lo, fb, kb * BK, n0);
                    } else {
                        ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        ptx::tma_load_2d(s, &mapA_hi, &full_bar[stage], kb * BK, m0);
                        ptx::tma_load_2d(s + Cfg::A_BYTES, &mapA_lo, &full_bar[stage], kb * BK, m0);
                        ptx::tma_load_2d(s + 2 * Cfg::A_BYTES, &mapB_hi, &full_bar[stage], kb * BK, n0);
                        ptx::tma_load_2d(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapB_lo, &full_bar[stage], kb * BK, n0);
                    }
                    if (++stage == STAGES) stage = 0;
                }
                if constexpr (LN) {
                    if (ln_ring) {
                        // residual tile [128 rows x 256 cols] hi + lo -> ring bytes [0, 128 KB): box (plane p, column
                        // block cb) at (p * 4 + cb) * 16 KB.  Needs the whole ring: every stage must have been read.
#pragma unroll
                        for (int s2 = 0; s2 < STAGES; ++s2) { ptx::mbar_wait(&empty_bar[s2], (uses[s2] & 1u) ^ 1u); uses[s2]++; }
                        ptx::mbar_expect_tx(rfull_bar, 8 * Cfg::A_BYTES);
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) {
                            ptx::tma_load_2d(smem + cb * Cfg::A_BYTES, &mapR_hi, rfull_bar, cb * UM_BK, m0);
                            ptx::tma_load_2d(smem + (4 + cb) * Cfg::A_BYTES, &mapR_lo, rfull_bar, cb * UM_BK, m0);
                        }
                    }
                }
                if (dyn) {                              // publish the next tile (or the end marker) to the consumers
                    const int q = (pit + 1) & 3;
                    ptx::mbar_wait(&sempty_bar[q], (((pit + 1) >> 2) & 1) ^ 1);
                    stile[q] = tile_next;
                    ptx::mbar_arrive(&sfull_bar[q]);
                }
                tile = tile_next;
            }
            // the last CTA to run dry re-arms the counters for the next launch that uses this slot
            if (dyn && atomicAdd(ep.sched + 1, 1) == (int)gridDim.x - 1) { ep.sched[0] = 0; ep.sched[1] = 0; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(UM_BM * CG, BN);
            int stage = 0; uint32_t phase = 0;
            for (int it = 0;; ++it) {
                int tile = grp + it * n_grp;
                if (dyn) {
                    ptx::mbar_wait(&sfull_bar[it & 3], (it >> 2) & 1);
                    tile = stile[it & 3];
                    ptx::mbar_arrive(&sempty_bar[it & 3]);
                }
                if (tile >= total_tiles) break;
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    if (kb == 0 && it == 0) TIP_TS(1);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = umma_smem_desc_bk<BK>(sa), a_lo = umma_smem_desc_bk<BK>(sa + Cfg::A_BYTES);
                    const uint64_t b_hi = umma_smem_desc_bk<BK>(sa + 2 * Cfg::A_BYTES);
                    const uint64_t b_lo = umma_smem_desc_bk<BK>(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 16 fp16 = 32 bytes along K
                        if constexpr (CG == 2) {
                            ptx::umma_f16_pair(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            ptx::umma_f16_pair(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            ptx::umma_f16_pair(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                        } else {
                            ptx::umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            ptx::umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            ptx::umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                        }
                    }
                    // smem stage free (in both CTAs of a pair) once these MMAs retire
                    if constexpr (CG == 2) ptx::umma_commit_pair(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (each CTA's epilogue reads its own 128 rows from its own TMEM)
                if constexpr (CG == 2) ptx::umma_commit_pair(&tfull_bar[as]); else ptx::umma_commit(&tfull_bar[as]);
                if (it == 0) TIP_TS(2);
            }
        }
    } else {
        // ================= epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp-2)/4 ====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CH = BN / 64;                               // 32-column chunks per warp
        float* stg = staging + (warp - 2) * 1024;
        const int crow = lane >> 3, cchunk = lane & 7;            // "coal" pattern coordinates
        const bool vec_ok = (ep.ldc & 3) == 0;
        const float asc = ep.acc_scale ? __ldg(ep.acc_scale) : 1.f;    // 1 / (s_a * s_w), a power of two
        const float osc = OUT_HALF ? ACT_SCALE : 1.f;             // output planes hold ACT_SCALE * x
        const float relu_floor = ep.relu ? 0.f : -INFINITY;
        // dropout on the GEMM output (element index = row * N + col): threshold and scale are kernel parameters
        // (constant bank, no registers), the site seed is parked in shared memory by the first epilogue thread --
        // the LayerNorm epilogue holds a 128-column row in registers and has none to spare
        volatile uint64_t* seed_slot = reinterpret_cast<volatile uint64_t*>(bars + 2 * STAGES + 6);
        for (int it = 0;; ++it) {
            int tile = grp + it * n_grp;
            if (dyn) {
                ptx::mbar_wait(&sfull_bar[it & 3], (it >> 2) & 1);
                tile = stile[it & 3];
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&sempty_bar[it & 3]);
            }
            if (tile >= total_tiles) break;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (m_tile0 + (tile / n_tiles) * CG + (int)cta_rank) * UM_BM, n0 = (tile % n_tiles) * BN;
            const int rbase = m0 + quarter * 32;                  // first row of this warp
            // lean path: full tile, vector stores, no dropout (warp-uniform); everything else -> slow path
            const bool fast = (m0 + UM_BM <= M) && (n0 + BN <= N) && vec_ok && !(ep.drop_p > 0.f) && !(ep.dbg & 7);
            if constexpr (!LN) {
                // this tile's bias slice (pre-multiplied by the output scale) -> shared memory while the MMAs still run:
                // the kernel leaves ~3 KB of L1, so a __ldg in the chunk loop is an exposed L2 round trip
                const int et = (int)threadIdx.x - 64;
                // (with dropout on this GEMM's output the kept elements' 1/(1-p) rides on the bias and the accumulator scale)
                if (et < BN) row_stat[(it & 1) * 256 + et] = (n0 + et < N) ? __ldg(ep.bias + n0 + et) * osc * (ep.drop_thr ? ep.drop_inv : 1.f) : 0.f;
                if (it == 0 && et == 0) *seed_slot = ep.drop_thr ? site_seed(ep.seed_ptr, ep.seed) : 0ull;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            ptx::mbar_wait(&tfull_bar[as], aphase);
            if (warp == 2 && it == 0) TIP_TS(3);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
            float v[32];
            if constexpr (!LN) {
                if (ep.tma_out && (n0 + BN <= N) && !(ep.dbg & 7)) {
                    // thread = accumulator row: bias/ReLU/split in registers, 32x32 output box through a
                    // swizzled 4 KB shared tile, written to global by TMA (no transposition, no LSU stores)
                    const float sc = asc * osc * (ep.drop_thr ? ep.drop_inv : 1.f);
                    uint8_t* sbuf = reinterpret_cast<uint8_t*>(stg);
                    const float* bs = row_stat + (it & 1) * 256 + half * (BN / 2);
                    // the next chunk's accumulator columns are requested from tensor memory before this chunk is processed
                    // (tcgen05.ld latency under the bias / split / store work instead of in front of it)
                    ptx::tmem_ld32(t_acc, v);
#pragma unroll 1
                    for (int c = 0; c < CH; ++c) {
                        const int colb = n0 + half * (BN / 2) + c * 32;
                        float vn[32];
                        if (c + 1 < CH) ptx::tmem_ld32_nowait(t_acc + (c + 1) * 32, vn);
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 b = *reinterpret_cast<const float4*>(bs + c * 32 + 4 * j4);       // broadcast
                            v[4 * j4 + 0] = fmaxf(fmaf(v[4 * j4 + 0], sc, b.x), relu_floor);
                            v[4 * j4 + 1] = fmaxf(fmaf(v[4 * j4 + 1], sc, b.y), relu_floor);
                            v[4 * j4 + 2] = fmaxf(fmaf(v[4 * j4 + 2], sc, b.z), relu_floor);
                            v[4 * j4 + 3] = fmaxf(fmaf(v[4 * j4 + 3], sc, b.w), relu_floor);
                        }
                        if (ep.drop_thr) {                            // reference: dropout(relu(linear1(x))) -- warp-uniform branch
                            const uint64_t g0 = ((uint64_t)(rbase + lane) * N + colb) >> 2;
                            const uint64_t dseed = *seed_slot;
                            const uint32_t thr_hi = ep.drop_thr << 16;
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {               // one hash per four columns; a dropped element is one compare + select
                                const uint64_t h = hash_u64(dseed, g0 + j4);
                                const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
                                if ((hl << 16) < thr_hi) v[4 * j4 + 0] = 0.f;
                                if (hl < thr_hi) v[4 * j4 + 1] = 0.f;
                                if ((hh << 16) < thr_hi) v[4 * j4 + 2] = 0.f;
                                if (hh < thr_hi) v[4 * j4 + 3] = 0.f;
                            }
                        }
                        if constexpr (OUT_HALF) {
                            // two [32 rows][64 B] tiles (hi, lo), 64B-swizzled: chunk ^= (row >> 1) & 3.  The two planes are
                            // separate bulk groups, so while one plane's box is still being read out of shared memory
                            // the other plane's tile can already be rewritten (wait_group.read 1, not 0).
                            const int sw = (lane >> 1) & 3;
                            uint4 uh[4], ul[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float h0, h1, l0, l1;
                                __half2 t;
                                veltkamp11(v[8 * j + 0], h0, l0); veltkamp11(v[8 * j + 1], h1, l1);
                                t = __floats2half2_rn(h0, h1); uh[j].x = *reinterpret_cast<uint32_t*>(&t);
                                t = __floats2half2_rn(l0, l1); ul[j].x = *reinterpret_cast<uint32_t*>(&t);
                                veltkamp11(v[8 * j + 2], h0, l0); veltkamp11(v[8 * j + 3], h1, l1);
                                t = __floats2half2_rn(h0, h1); uh[j].y = *reinterpret_cast<uint32_t*>(&t);
                                t = __floats2half2_rn(l0, l1); ul[j].y = *reinterpret_cast<uint32_t*>(&t);
                                veltkamp11(v[8 * j + 4], h0, l0); veltkamp11(v[8 * j + 5], h1, l1);
                                t = __floats2half2_rn(h0, h1); uh[j].z = *reinterpret_cast<uint32_t*>(&t);
                                t = __floats2half2_rn(l0, l1); ul[j].z = *reinterpret_cast<uint32_t*>(&t);
                                veltkamp11(v[8 * j + 6], h0, l0); veltkamp11(v[8 * j + 7], h1, l1);
                                t = __floats2half2_rn(h0, h1); uh[j].w = *reinterpret_cast<uint32_t*>(&t);
                                t = __floats2half2_rn(l0, l1); ul[j].w = *reinterpret_cast<uint32_t*>(&t);
                            }
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // hi tile free
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((j ^ sw) << 4)) = uh[j];
                            ptx::fence_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                ptx::tma_store_2d(&mapC0, sbuf, colb, rbase);
                                ptx::bulk_commit();
                                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");              // lo tile free
                            }
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sbuf + 2048 + lane * 64 + ((j ^ sw) << 4)) = ul[j];
                            ptx::fence_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                ptx::tma_store_2d(&mapC1, sbuf + 2048, colb, rbase);
                                ptx::bulk_commit();
                            }
                        } else {
                            if (lane == 0) ptx::bulk_wait_read0();        // previous box has left the shared tile
                            __syncwarp();
                            // one [32 rows][128 B] fp32 tile, 128B-swizzled: chunk ^= row & 7
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            ptx::fence_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                ptx::tma_store_2d(&mapC0, sbuf, colb, rbase);
                                ptx::bulk_commit();
                            }
                        }
                        if (c + 1 < CH) {
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = vn[j];
                        }
                    }
                } else if (fast) {
#pragma unroll 1
                    for (int c = 0; c < CH; ++c) {
                        const int col = n0 + half * (BN / 2) + c * 32 + cchunk * 4;
                        float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
                        b4.x *= osc; b4.y *= osc; b4.z *= osc; b4.w *= osc;
                        const float sc = asc * osc;
                        ptx::tmem_ld32(t_acc + c * 32, v);
                        stg_write_row(stg, lane, v);
                        __syncwarp();
                        const size_t obase = (size_t)(rbase + crow) * ep.ldc + col;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 x = *stg_ptr(stg, i * 4 + crow, cchunk);
                            x.x = fmaxf(fmaf(x.x, sc, b4.x), relu_floor); x.y = fmaxf(fmaf(x.y, sc, b4.y), relu_floor);
                            x.z = fmaxf(fmaf(x.z, sc, b4.z), relu_floor); x.w = fmaxf(fmaf(x.w, sc, b4.w), relu_floor);
                            const size_t off = obase + (size_t)(i * 4) * ep.ldc;
                            if (ep.dbg & 8) { if (x.x == 123.456f) ep.out[off] = x.y; continue; }     // experiment: no stores
                            if (ep.dbg & 16) {                                                         // experiment: no split math
                                if constexpr (OUT_HALF) { *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out) + off) = make_uint2(__float_as_uint(x.x), __float_as_uint(x.y));
                                    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out_lo) + off) = make_uint2(__float_as_uint(x.z), __float_as_uint(x.w)); continue; }
                            }
                            if constexpr (OUT_HALF) half_split_store4_fast(reinterpret_cast<__half*>(ep.out) + off,
                                                                           reinterpret_cast<__half*>(ep.out_lo) + off, x);
                            else *reinterpret_cast<float4*>(ep.out + off) = x;
                        }
                        __syncwarp();
                    }
                } else if (!OUT_HALF && !(ep.drop_p > 0.f) && !(ep.dbg & 7)) {
                    // fp32 output with arbitrary N / row pitch (the head: ldc = size_s = 131): lane <-> column,
                    // one fully coalesced 128-byte row segment per store instruction
#pragma unroll 1
                    for (int c = 0; c < CH; ++c) {
                        const int colb = n0 + half * (BN / 2) + c * 32;
                        if (colb >= N) break;                                 // warp-uniform
                        ptx::tmem_ld32(t_acc + c * 32, v);
                        stg_write_row(stg, lane, v);
                        __syncwarp();
                        const int col = colb + lane;
                        const float bv = col < N ? __ldg(ep.bias + col) : 0.f;
                        const int nrow = min(32, M - rbase);
#pragma unroll 4
                        for (int r = 0; r < nrow; ++r) {
                            const float x = reinterpret_cast<const float*>(stg_ptr(stg, r, lane >> 2))[lane & 3];
                            if (col < N) ep.out[(size_t)(rbase + r) * ep.ldc + col] = fmaxf(fmaf(x, asc, bv), relu_floor);
                        }
                        __syncwarp();
                    }
                } else {
                    const float inv_keep = ep.drop_inv;
                    const uint64_t dseed = *seed_slot;
#pragma unroll 1
                    for (int c = 0; c < CH; ++c) {
                        const int colb = n0 + half * (BN / 2) + c * 32;     // first column of the chunk
                        if (colb >= N) break;                                 // warp-uniform
                        ptx::tmem_ld32(t_acc + c * 32, v);
                        stg_write_row(stg, lane, v);
                        __syncwarp();
                        const int col = colb + cchunk * 4;
#pragma unroll 1
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + crow, row = rbase + r;
                            const float4 x4 = *stg_ptr(stg, r, cchunk);
                            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (row < M && col + q < N && !(ep.dbg & 1)) {
                                    float x = fmaf(xs[q], asc, (ep.dbg & 2) ? 0.f : __ldg(ep.bias + col + q));
                                    x = fmaxf(x, relu_floor);
                                    if (ep.drop_p > 0.f) x *= dropout_factor(ep.drop_p, inv_keep, dseed, (uint64_t)row * N + col + q);
                                    const size_t off = (size_t)row * ep.ldc + col + q;
                                    if constexpr (OUT_HALF) {
                                        __half hi, lo;
                                        half_split(x * ACT_SCALE, hi, lo);
                                        reinterpret_cast<__half*>(ep.out)[off] = hi;
                                        reinterpret_cast<__half*>(ep.out_lo)[off] = lo;
                                    } else {
                                        ep.out[off] = x;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            } else {
                // LayerNorm epilogue: out = LN(acc*asc + bias [dropout] + residual) * gamma + beta, BN == 256
                // == the whole row.  Residual and output are FP16 hi/lo planes of ACT_SCALE * x.
                if (ln_ring) {
                    // ---- fast path: thread = accumulator row, its 128 columns stay in registers after ONE TMEM read ----
                    // residual: from the swizzled boxes the producer parked in the ring; bias / gamma / beta: from
                    // shared memory (the kernel runs with ~3 KB of L1, so every __ldg would be an L2 round trip);
                    // output: 32x32 boxes staged in the ring's tail (two 4 KB buffers per warp) and TMA-stored.
                    float* cvec = staging;                            // [0,256) bias, [256,512) gamma*16, [512,768) beta*16
                    if (it == 0) {
                        const int t = (int)threadIdx.x - 64;          // 0..255 among the epilogue threads
                        cvec[t] = __ldg(ep.bias + t) * (DROP ? ep.drop_inv : 1.f);       // kept elements carry 1/(1-p)
                        cvec[256 + t] = __ldg(ep.gamma + t) * ACT_SCALE;
                        cvec[512 + t] = __ldg(ep.beta + t) * ACT_SCALE;
                        if (t == 0) *seed_slot = ep.drop_thr ? site_seed(ep.seed_ptr, ep.seed) : 0ull;
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    const int trow = quarter * 32 + lane;             // row within the tile
                    const int col0 = half * (BN / 2);
                    float x[BN / 2];
#pragma unroll
                    for (int c = 0; c < CH; ++c) ptx::tmem_ld32_nowait(t_acc + c * 32, *reinterpret_cast<float(*)[32]>(&x[c * 32]));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // the accumulator is in registers: hand the TMEM buffer back to the MMA warp now
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) ptx::mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tempty_bar[as]), 0u));
                        else ptx::mbar_arrive(&tempty_bar[as]);
                    }
                    ptx::mbar_wait(rfull_bar, (uint32_t)(it & 1));    // residual tile landed (issued right after the last k-block)
                    float rsum = 0.f;
                    uint64_t sd_row = 0ull;                           // DROP: site seed + this row's first hash group (one shared-memory read per tile)
                    if constexpr (DROP) sd_row = *seed_slot;
                    const uint8_t* rrow = smem + trow * 128;          // this row inside every 16 KB box
                    const int rsw = trow & 7;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int cb = half * 2 + (c >> 1);           // 64-column box of this chunk
                        const uint8_t* bh = rrow + cb * Cfg::A_BYTES;
                        const uint8_t* bl = rrow + (4 + cb) * Cfg::A_BYTES;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {                 // 16-byte chunk = 8 columns
                            const int pos = (((c & 1) * 4 + i) ^ rsw) << 4;
                            const uint4 h4 = *reinterpret_cast<const uint4*>(bh + pos);
                            const uint4 l4 = *reinterpret_cast<const uint4*>(bl + pos);
                            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w};
                            const uint32_t lw[4] = {l4.x, l4.y, l4.z, l4.w};
                            const float4 b0 = *reinterpret_cast<const float4*>(cvec + col0 + c * 32 + i * 8);       // broadcast
                            const float4 b1 = *reinterpret_cast<const float4*>(cvec + col0 + c * 32 + i * 8 + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            // dropout1 / dropout2 of the encoder layer (DROP): on the sub-layer output, before the residual
                            // add.  One hash per four columns; element e of a group is dropped when its 16-bit lane of the
                            // hash is below drop_thr; the kept ones carry 1/(1-p) through the pre-scaled asc / bias.
                            uint32_t hw2[4] = {0u, 0u, 0u, 0u};
                            if constexpr (DROP) {
                                const uint64_t g0 = ((uint64_t)(m0 + trow) * N + col0 + c * 32 + i * 8) >> 2;
                                const uint64_t sd = sd_row;
                                const uint64_t ha = hash_u64(sd, g0), hb2 = hash_u64(sd, g0 + 1);
                                hw2[0] = (uint32_t)ha; hw2[1] = (uint32_t)(ha >> 32); hw2[2] = (uint32_t)hb2; hw2[3] = (uint32_t)(hb2 >> 32);
                            }
                            const float ascd = DROP ? asc * ep.drop_inv : asc;
                            const uint32_t thr_hi = ep.drop_thr << 16;
#pragma unroll
                            for (int q2 = 0; q2 < 4; ++q2) {
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[q2]));
                                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[q2]));
                                const int j = c * 32 + i * 8 + q2 * 2;
                                float s0 = fmaf(x[j], ascd, bb[q2 * 2]), s1 = fmaf(x[j + 1], ascd, bb[q2 * 2 + 1]);
                                if constexpr (DROP) {
                                    if ((hw2[q2] << 16) < thr_hi) s0 = 0.f;          // lane 0: low 16 bits
                                    if (hw2[q2] < thr_hi) s1 = 0.f;                  // lane 1: high 16 bits
                                }
                                const float v0 = fmaf(hf.x + lf.x, 1.f / ACT_SCALE, s0);
                                const float v1 = fmaf(hf.y + lf.y, 1.f / ACT_SCALE, s1);
                                x[j] = v0; x[j + 1] = v1;
                                rsum += v0 + v1;
                            }
                        }
                    }
                    if (warp == 2 && it == 0) TIP_TS(4);
                    row_stat[half * 128 + trow] = rsum;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                    const float mean = (row_stat[trow] + row_stat[128 + trow]) * (1.f / BN);
                    float q2s = 0.f;
#pragma unroll
                    for (int j = 0; j < BN / 2; ++j) { const float d = x[j] - mean; q2s = fmaf(d, d, q2s); }
                    row_stat[256 + half * 128 + trow] = q2s;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                    const float var = (row_stat[256 + trow] + row_stat[256 + 128 + trow]) * (1.f / BN);
                    const float ca = rsqrtf(var + 1e-5f), cb2 = -mean * ca;
                    if (warp == 2 && it == 0) TIP_TS(5);
                    uint8_t* obuf = smem + 8 * Cfg::A_BYTES + (warp - 2) * 8192;      // two 4 KB buffers (hi 2 KB | lo 2 KB)
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int colb = col0 + c * 32;
                        uint8_t* sbuf = obuf + (c & 1) * 4096;
                        if (c >= 2) {                                 // the box stored two chunks ago has left this buffer
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 g0 = *reinterpret_cast<const float4*>(cvec + 256 + colb + 8 * j);
                            const float4 g1 = *reinterpret_cast<const float4*>(cvec + 256 + colb + 8 * j + 4);
                            const float4 e0 = *reinterpret_cast<const float4*>(cvec + 512 + colb + 8 * j);
                            const float4 e1 = *reinterpret_cast<const float4*>(cvec + 512 + colb + 8 * j + 4);
                            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                            const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                            uint32_t uh[4], ul[4];
#pragma unroll
                            for (int p2 = 0; p2 < 4; ++p2) {
                                const int jj = c * 32 + 8 * j + 2 * p2;
                                const float y0 = fmaf(fmaf(x[jj], ca, cb2), gg[2 * p2], ee[2 * p2]);
                                const float y1 = fmaf(fmaf(x[jj + 1], ca, cb2), gg[2 * p2 + 1], ee[2 * p2 + 1]);
                                float h0, h1, l0, l1;
                                veltkamp11(y0, h0, l0); veltkamp11(y1, h1, l1);
                                __half2 t2 = __floats2half2_rn(h0, h1); uh[p2] = *reinterpret_cast<uint32_t*>(&t2);
                                t2 = __floats2half2_rn(l0, l1); ul[p2] = *reinterpret_cast<uint32_t*>(&t2);
                            }
                            const int off = lane * 64 + ((j ^ sw) << 4);
                            *reinterpret_cast<uint4*>(sbuf + off) = make_uint4(uh[0], uh[1], uh[2], uh[3]);
                            *reinterpret_cast<uint4*>(sbuf + 2048 + off) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
                        }
                        ptx::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapC0, sbuf, colb, rbase);
                            ptx::tma_store_2d(&mapC1, sbuf + 2048, colb, rbase);
                            ptx::bulk_commit();
                        }
                    }
                    if (warp == 2 && it == 0) TIP_TS(6);
                    // give the ring back to the producer once every warp's boxes have been read out of it
                    if (lane == 0) ptx::bulk_wait_read0();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 64) {
#pragma unroll
                        for (int s2 = 0; s2 < STAGES; ++s2) ptx::mbar_arrive(&empty_bar[s2]);
                    }
                    continue;                                         // tempty already arrived
                }
                // (the LayerNorm epilogue always takes the ring path above)
            }
            if (warp == 2 && it == 0) TIP_TS(6);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                      // 8 (x2 CTAs) arrivals free the accumulator
                if constexpr (CG == 2) ptx::mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tempty_bar[as]), 0u));
                else ptx::mbar_arrive(&tempty_bar[as]);
            }
        }
    }
    if (!pdl_early) griddep_launch();                              // late trigger: the next kernel's launch + prologue overlap only this tail
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();                 // outstanding TMA stores of this warp
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) { cluster_arrive(); cluster_wait(); }   // neither CTA leaves while the pair's MMAs / arrivals can touch it
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc2(tmem_base, Cfg::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
    if (warp == 2) TIP_TS(7);
#undef TIP_TS
}

// ------------------------------------------------------------------------------------------------
// Host side: TMA descriptors of every operand plane and the launch table.
struct UmmaOperand { CUtensorMap hi, lo; };
struct UmmaOutput { CUtensorMap c0, c1; bool valid = false; };   // 32x32 store boxes (hi/lo fp16 planes, or one fp32 plane)
struct UmmaMaps {
    UmmaOperand a_xin, a_xa, a_xb, a_att, a_hid, a_hs;                    // activations (box 32 x 128)
    UmmaOperand w_in, w_qkv[MAX_LAYERS], w_o[MAX_LAYERS], w_1[MAX_LAYERS], w_2[MAX_LAYERS], w_ih, w_l;
    UmmaOperand w_qkv256[MAX_LAYERS], w_1256[MAX_LAYERS], w_ih256;        // 256-row boxes of the same planes (wide tiles)
    UmmaOperand a_att32, a_hid32, w_o32[MAX_LAYERS];                       // 32-column k-blocks of the LayerNorm GEMMs' operands (two-row-tile kernel, tip_umma_ln2.cuh; W2: w_2k32)
    UmmaOperand w_2h[MAX_LAYERS];                                          // W2, 128-row boxes: each CTA of a LayerNorm pair tile stages half of B
    UmmaOperand w_o64[MAX_LAYERS], w_264[MAX_LAYERS];                     // 64-row boxes (skinny-M LayerNorm GEMMs, 4 CTAs per row tile)
    UmmaOperand w_qkv64[MAX_LAYERS], w_164[MAX_LAYERS], w_ih64;           // 64-row boxes: each CTA of a 256 x 128 pair tile stages half of B
    UmmaOperand a_xa32, a_xb32, w_qkv256k32[MAX_LAYERS], w_1256k32[MAX_LAYERS];   // 32-column (64-byte, SWIZZLE_64B) k-blocks: 128 x 256 tiles with a 4-stage ring
    UmmaOperand w_1k32[MAX_LAYERS], w_2k32[MAX_LAYERS];                            // W1 (128-row boxes) / W2 (256-row boxes), 32-column k-blocks: fused FFN kernel
    UmmaOperand w_qkvr[MAX_LAYERS];                                                // re-ordered in_proj rows, 192-row x 32-column boxes (fused QKV + attention)
    UmmaOperand a_xin32, a_hs32, w_in256k32, w_l192k32;                            // ... in_linear as ONE 128 x 256 tile per row tile, the head as ONE 128 x 192 tile
    UmmaOutput o_pre;                                                      // fp32 [rows][256] scratch of the un-fused LayerNorm path
    UmmaOutput o_xa, o_xb, o_hid, o_qkv, o_gi;                             // TMA-store targets
    UmmaOperand w_hh;                                                      // resident A operand of the recurrence (box 64 x 64)
    int num_sms = 148;
    int ln_grid = 0;                                                       // CTAs per fused-LayerNorm launch (0 = one per row tile)
    int ln_pair_min_k = 0;                                                 // LayerNorm GEMMs with K >= this run on CTA pairs (0 = never)
    bool attrs_set = false;
};

typedef CUresult (*tip_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline tip_encode_tiled_fn umma_encode_fn() {
    static tip_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tip_encode_tiled_fn>(p);
    }
    return fn;
}

// 2-D fp16 row-major [rows][cols] plane, box = 64 columns (128 bytes, swizzled) x box_rows rows
inline bool umma_make_map(CUtensorMap* map, const __half* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int bk = UM_BK) {
    tip_encode_tiled_fn enc = umma_encode_fn();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)bk, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// output box: 32 rows x 32 columns; fp16 planes (64-byte rows, SWIZZLE_64B) or fp32 (128-byte rows, SWIZZLE_128B)
inline bool umma_make_store_map(CUtensorMap* map, void* base, bool is_half, uint64_t rows, uint64_t cols) {
    tip_encode_tiled_fn enc = umma_encode_fn();
    if (!enc) return false;
    const size_t es = is_half ? sizeof(__half) : sizeof(float);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * es};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride,
               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, is_half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int umma_build_maps(UmmaMaps& mp, const float* blob, const PackOff& o, const Dims& d, float* xin,
                           size_t plane_xin, float* xa, float* xb, float* att, size_t plane_e, float* hid,
                           size_t plane_f, float* hs, size_t plane_r, float* qkv, float* gi, float* pre, int cap_rows,
                           std::string& err) {
    bool ok = true;
    auto outp = [&](UmmaOutput& op, float* p, size_t plane, int cols, bool is_half) {
        if (is_half) {
            __half* h = reinterpret_cast<__half*>(p);
            op.valid = umma_make_store_map(&op.c0, h, true, cap_rows, cols) && umma_make_store_map(&op.c1, h + plane, true, cap_rows, cols);
        } else {
            op.valid = umma_make_store_map(&op.c0, p, false, cap_rows, cols);
            op.c1 = op.c0;
        }
        ok = ok && op.valid;
    };
    outp(mp.o_xa, xa, plane_e, E, true);
    outp(mp.o_xb, xb, plane_e, E, true);
    outp(mp.o_hid, hid, plane_f, F, true);
    outp(mp.o_qkv, qkv, (size_t)cap_rows * 3 * E, 3 * E, true);      // FP16 hi/lo planes of 16*q|k|v for the mma attention
    outp(mp.o_gi, gi, 0, R, false);
    outp(mp.o_pre, pre, 0, E, false);        // fp32 [rows][256] scratch of the un-fused LayerNorm path
    // activation planes: hi at the start of the buffer, lo `plane` halves later (same bytes as one fp32 plane)
    auto act = [&](UmmaOperand& op, const float* p, size_t plane, int cols, int bk = UM_BK) {
        const __half* h = reinterpret_cast<const __half*>(p);
        ok = ok && umma_make_map(&op.hi, h, cap_rows, cols, UM_BM, bk) && umma_make_map(&op.lo, h + plane, cap_rows, cols, UM_BM, bk);
    };
    auto wgt = [&](UmmaOperand& op, size_t hi, size_t lo, int rows, int cols, int bn, int bk = UM_BK) {
        ok = ok && umma_make_map(&op.hi, reinterpret_cast<const __half*>(blob + hi), rows, cols, bn, bk) &&
             umma_make_map(&op.lo, reinterpret_cast<const __half*>(blob + lo), rows, cols, bn, bk);
    };
    act(mp.a_xin, xin, plane_xin, d.kin_pad);
    act(mp.a_xa, xa, plane_e, E);
    act(mp.a_xb, xb, plane_e, E);
    act(mp.a_xa32, xa, plane_e, E, 32);
    act(mp.a_xin32, xin, plane_xin, d.kin_pad, 32);
    act(mp.a_hs32, hs, plane_r, R, 32);
    wgt(mp.w_in256k32, o.win_hi, o.win_lo, E, d.kin_pad, 256, 32);
    wgt(mp.w_l192k32, o.wl_hi, o.wl_lo, HEAD_NPAD, d.khead, 192, 32);      // rows 144..191 of the box: out of bounds -> zero fill
    act(mp.a_xb32, xb, plane_e, E, 32);
    act(mp.a_att, att, plane_e, E);
    act(mp.a_att32, att, plane_e, E, 32);
    act(mp.a_hid32, hid, plane_f, F, 32);
    act(mp.a_hid, hid, plane_f, F);
    act(mp.a_hs, hs, plane_r, R);
    wgt(mp.w_in, o.win_hi, o.win_lo, E, d.kin_pad, 128);
    for (int l = 0; l < d.layers; ++l) {
        const LayerOff& L = o.layer[l];
        wgt(mp.w_qkv[l], L.wqkv_hi, L.wqkv_lo, 3 * E, E, 128);
        wgt(mp.w_o[l], L.wo_hi, L.wo_lo, E, E, 256);
        wgt(mp.w_1[l], L.w1_hi, L.w1_lo, F, E, 128);
        wgt(mp.w_qkv256[l], L.wqkv_hi, L.wqkv_lo, 3 * E, E, 256);
        wgt(mp.w_o64[l], L.wo_hi, L.wo_lo, E, E, 64);
        wgt(mp.w_264[l], L.w2_hi, L.w2_lo, E, F, 64);
        wgt(mp.w_1256[l], L.w1_hi, L.w1_lo, F, E, 256);
        wgt(mp.w_qkv64[l], L.wqkv_hi, L.wqkv_lo, 3 * E, E, 64);
        wgt(mp.w_qkv256k32[l], L.wqkv_hi, L.wqkv_lo, 3 * E, E, 256, 32);
        wgt(mp.w_qkvr[l], L.wqkvr_hi, L.wqkvr_lo, 3 * E, E, 192, 32);
        wgt(mp.w_1k32[l], L.w1_hi, L.w1_lo, F, E, 128, 32);
        wgt(mp.w_2k32[l], L.w2_hi, L.w2_lo, E, F, 256, 32);
        wgt(mp.w_1256k32[l], L.w1_hi, L.w1_lo, F, E, 256, 32);
        wgt(mp.w_164[l], L.w1_hi, L.w1_lo, F, E, 64);
        wgt(mp.w_2[l], L.w2_hi, L.w2_lo, E, F, 256);
        wgt(mp.w_2h[l], L.w2_hi, L.w2_lo, E, F, 128);
        wgt(mp.w_o32[l], L.wo_hi, L.wo_lo, E, E, 256, 32);
    }
    if (d.with_rnn) wgt(mp.w_ih, o.wih_hi, o.wih_lo, R, E, 128);
    if (d.with_rnn) wgt(mp.w_ih256, o.wih_hi, o.wih_lo, R, E, 256);
    if (d.with_rnn) wgt(mp.w_ih64, o.wih_hi, o.wih_lo, R, E, 64);
    if (d.with_rnn) wgt(mp.w_hh, o.whh_hi, o.whh_lo, R, R, 16);     // 16-row boxes: hi/lo rows interleave per TMEM quarter
    wgt(mp.w_l, o.wl_hi, o.wl_lo, HEAD_NPAD, d.khead, 128);
    if (!ok) { err = "cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor-map arguments)"; return TIP_ERR_CUDA; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&mp.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (!mp.attrs_set) {
        cudaFuncSetAttribute(umma_gemm_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, true, true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, false, true, 1, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 1, 32>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<192, false, false, 1, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<192, 1, 32>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<128, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<128, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_gemm_kernel<256, true, true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 2>::SMEM_BYTES);
        mp.attrs_set = true;
    }
    return TIP_OK;
}

inline int wide_kind() {       // TIP_BN256: 1 = 128 x 256 single-CTA tiles, 64-wide k-blocks (2-stage ring: exposes TMA latency, measured no gain);
                               //            2 = the same tiles with 32-wide k-blocks (4-stage ring), qkv and ff1 only
    static const int v = getenv("TIP_BN256") ? atoi(getenv("TIP_BN256")) : 0;
    return v;
}
inline bool wide_mode() { return wide_kind() == 1; }
inline bool one_tile_mode() {  // TIP_ONE_TILE=1: in_linear (N = 256) and the head (N <= 144) as ONE tile per 128-row tile (A read once; 32-wide k-blocks)
    static const int v = getenv("TIP_ONE_TILE") ? atoi(getenv("TIP_ONE_TILE")) : 0;
    return v != 0;
}
inline int pair_kind() {       // TIP_PAIR: 1 = 256 x 256 pair tiles, 2 = 256 x 128 pair tiles (each CTA: its 128 rows of A, 64 rows of B)
    static const int v = getenv("TIP_PAIR") ? atoi(getenv("TIP_PAIR")) : 0;
    return v;
}
inline bool pair_mode_unused() {      // experiment (TIP_PAIR=1): CTA-pair tiles for the wide non-LN GEMMs.  Measured: the mainloop becomes MMA-bound (4.2 us per 256x256 tile) but the 128x256 epilogue per CTA (5 us) is then the critical path -> no gain over 128x128 single-CTA tiles
    static const int v = getenv("TIP_PAIR") ? atoi(getenv("TIP_PAIR")) : 0;
    return v != 0;
}

// skinny: (LayerNorm GEMMs at small M) run the plain GEMM with 64-column tiles -- 4 CTAs per row tile instead of 1 --
// into the fp32 scratch `o_pre`; the caller follows with resid_ln_kernel.
inline void umma_gemm(UmmaMaps& mp, int which, int layer, int M, int N, int K, const Epi& ep_in, bool ln,
                      cudaStream_t st, int m_tile0 = 0, int m_tile_cnt = -1, bool skinny = false) {
    pdl_kind() = 1;
    const UmmaOperand *A = nullptr, *B = nullptr, *B256 = nullptr, *B64 = nullptr;
    const UmmaOutput* C = nullptr;
    switch (which) {
        case UG_IN:     A = &mp.a_xin; B = &mp.w_in; C = &mp.o_xa; break;
        case UG_QKV:    A = &mp.a_xa;  B = &mp.w_qkv[layer]; B256 = &mp.w_qkv256[layer]; B64 = &mp.w_qkv64[layer]; C = &mp.o_qkv; break;
        case UG_OUT:    A = &mp.a_att; B = &mp.w_o[layer]; C = &mp.o_xb; break;
        case UG_FF1:    A = &mp.a_xb;  B = &mp.w_1[layer]; B256 = &mp.w_1256[layer]; B64 = &mp.w_164[layer]; C = &mp.o_hid; break;
        case UG_FF2:    A = &mp.a_hid; B = &mp.w_2[layer]; C = &mp.o_xa; break;
        case UG_IH:     A = &mp.a_xa;  B = &mp.w_ih; B256 = &mp.w_ih256; B64 = &mp.w_ih64; C = &mp.o_gi; break;
        case UG_HEAD_R: A = &mp.a_hs;  B = &mp.w_l; break;      // y (ldc = size_s, unaligned): plain stores
        default:        A = &mp.a_xa;  B = &mp.w_l; break;      // UG_HEAD_E
    }
    Epi ep = ep_in;
    ep.tma_out = (C && C->valid && !getenv("TIP_NO_TMA_STORE")) ? 1 : 0;
    const CUtensorMap& c0 = C ? C->c0 : A->hi;
    const CUtensorMap& c1 = C ? C->c1 : A->lo;
    // residual of the LayerNorm GEMMs, read through the same 64-column operand boxes the next GEMM uses
    const UmmaOperand* Rm = (which == UG_OUT) ? &mp.a_xa : (which == UG_FF2) ? &mp.a_xb : A;
    const int m_tiles = m_tile_cnt >= 0 ? m_tile_cnt : (M + UM_BM - 1) / UM_BM;
    if (skinny) {
        const UmmaOperand* B64 = (which == UG_OUT) ? &mp.w_o64[layer] : &mp.w_264[layer];
        ep.tma_out = (mp.o_pre.valid && !getenv("TIP_NO_TMA_STORE")) ? 1 : 0;
        const int tiles = m_tiles * (N / 64);
        launch_k(umma_gemm_kernel<64, false, false>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<64>::SMEM_BYTES, st, A->hi, A->lo, B64->hi, B64->lo, mp.o_pre.c0, mp.o_pre.c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        return;
    }
    if (ln && which == UG_FF2 && mp.ln_pair_min_k > 0 && K >= mp.ln_pair_min_k && (m_tiles % 2) == 0 && m_tiles >= 32) {
        // LayerNorm GEMM on CTA pairs (cta_group::2, 256 x 256 tiles): each CTA stages its own 128 rows of A and HALF of the
        // weight k-block, so the bytes a CTA pulls through its L2 port per row tile drop from A + W to A + W / 2 (ff2:
        // 1.77 -> 1.28 MB; the SM's port moves ~78 GB/s, which is what paces this kernel, not the MMAs)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * std::min(m_tiles / 2, mp.num_sms / 2));
        cfg.blockDim = dim3(UM_THREADS);
        cfg.dynamicSmemBytes = UmmaCfg<256, 2>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (ep.drop_thr)
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<256, true, true, 2, true>, A->hi, A->lo, mp.w_2h[layer].hi, mp.w_2h[layer].lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<256, true, true, 2>, A->hi, A->lo, mp.w_2h[layer].hi, mp.w_2h[layer].lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if (ln) {
        const int tiles = mp.ln_grid > 0 ? std::min(m_tiles, mp.ln_grid) : m_tiles;   // BN = 256 = the whole row; CTAs loop over the row tiles
        if (ep.drop_thr)
            launch_k(umma_gemm_kernel<256, true, true, 1, true>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256>::SMEM_BYTES, st, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            launch_k(umma_gemm_kernel<256, true, true>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256>::SMEM_BYTES, st, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if (B64 && (N % 128) == 0 && (m_tiles % 2) == 0 && (m_tiles / 2) * (N / 128) >= 32 && pair_kind() == 2) {
        // CTA pairs, 256 x 128 tiles: the pair's two CTAs stage their own 128 rows of A and 64 rows of B each -- 48 KB of
        // operands per CTA per k-block instead of 64 KB -- and keep the 128 x 128 epilogue of the single-CTA tiles
        const int units = (m_tiles / 2) * (N / 128);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * std::min(units, mp.num_sms / 2));
        cfg.blockDim = dim3(UM_THREADS);
        cfg.dynamicSmemBytes = UmmaCfg<128, 2>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (ep.out_lo)
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<128, false, true, 2>, A->hi, A->lo, B64->hi, B64->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<128, false, false, 2>, A->hi, A->lo, B64->hi, B64->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if (B256 && (N % 256) == 0 && (m_tiles % 2) == 0 && (m_tiles / 2) * (N / 256) >= 32 && pair_kind() == 1) {
        // CTA pairs (cta_group::2): 256 x 256 tiles, each CTA stages its 128 rows of A and half of B
        const int units = (m_tiles / 2) * (N / 256);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * std::min(units, mp.num_sms / 2));
        cfg.blockDim = dim3(UM_THREADS);
        cfg.dynamicSmemBytes = UmmaCfg<256, 2>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (ep.out_lo)
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<256, false, true, 2>, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            cudaLaunchKernelEx(&cfg, umma_gemm_kernel<256, false, false, 2>, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if ((which == UG_HEAD_R || which == UG_HEAD_E) && one_tile_mode() && m_tiles >= 64) {
        const UmmaOperand* A32 = (which == UG_HEAD_R) ? &mp.a_hs32 : &mp.a_xa32;
        launch_k(umma_gemm_kernel<192, false, false, 1, false, 32>, dim3(std::min(m_tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<192, 1, 32>::SMEM_BYTES, st, A32->hi, A32->lo, mp.w_l192k32.hi, mp.w_l192k32.lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if (which == UG_IN && ep.out_lo && one_tile_mode() && m_tiles >= 64) {
        launch_k(umma_gemm_kernel<256, false, true, 1, false, 32>, dim3(std::min(m_tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256, 1, 32>::SMEM_BYTES, st, mp.a_xin32.hi, mp.a_xin32.lo, mp.w_in256k32.hi, mp.w_in256k32.lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if ((which == UG_QKV || which == UG_FF1) && ep.out_lo && m_tiles * (N / 256) >= mp.num_sms && wide_kind() == 2) {
        const UmmaOperand* A32 = (which == UG_QKV) ? &mp.a_xa32 : &mp.a_xb32;
        const UmmaOperand* B32 = (which == UG_QKV) ? &mp.w_qkv256k32[layer] : &mp.w_1256k32[layer];
        const int tiles = m_tiles * (N / 256);
        launch_k(umma_gemm_kernel<256, false, true, 1, false, 32>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256, 1, 32>::SMEM_BYTES, st, A32->hi, A32->lo, B32->hi, B32->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else if (B256 && (N % 256) == 0 && m_tiles * (N / 256) >= mp.num_sms && wide_mode()) {
        // wide tiles: A is re-used over 256 columns (25 % less L2->SM operand traffic per flop)
        const int tiles = m_tiles * (N / 256);
        if (ep.out_lo)
            launch_k(umma_gemm_kernel<256, false, true>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256>::SMEM_BYTES, st, A->hi, A->lo, B256->hi, B256->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            launch_k(umma_gemm_kernel<256, false, false>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<256>::SMEM_BYTES, st, A->hi, A->lo, B256->hi, B256->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    } else {
        const int tiles = m_tiles * ((N + 127) / 128);
        if (ep.out_lo)
            launch_k(umma_gemm_kernel<128, false, true>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<128>::SMEM_BYTES, st, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
        else
            launch_k(umma_gemm_kernel<128, false, false>, dim3(std::min(tiles, mp.num_sms)), dim3(UM_THREADS), UmmaCfg<128>::SMEM_BYTES, st, A->hi, A->lo, B->hi, B->lo, c0, c1, Rm->hi, Rm->lo, M, N, K, m_tile0, m_tiles, ep);
    }
}

}  // namespace tip
