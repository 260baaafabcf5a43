// GEMM variant with the A operand resident in TENSOR MEMORY: C[M, N] = A[M, 256] * W[N, 256]^T, N a multiple of 128.
//
// Why: with the 3-product FP16 split every k-step of a 128 x 128 tile reads 3 x (A 4 KB + B 4 KB) of operands from shared
// memory -- 24 KB per 192 tensor-pipe clocks = the SM's whole 128 B/clk -- while TMA writes the next stages into the
// same memory, and every 128 x 128 tile pulls 256 KB of operands plus 64 KB of output through the SM's port to L2, which
// moves ~78 GB/s (loads and stores together; measured): that port is what paces the wide GEMMs of the path
// (in_linear, qkv, ff1, rnn_ih: K = 256, N = 256 .. 1024).  Here
//   * the A tile of a row tile (128 x 256, FP16 hi + lo = 128 KB) is loaded ONCE -- TMA into the operand ring, then
//     `tcgen05.cp` (shared -> tensor memory, 128 lanes x 256 bit per instruction, the same swizzled K-major descriptors an
//     SS MMA would read) into 256 columns of tensor memory -- and feeds `tcgen05.mma` with A in TMEM (TS form) for every
//     n-tile the CTA computes in that row tile;
//   * only W is streamed afterwards (TMA, 128B-swizzled 64-wide k-blocks, 5-stage ring): half the shared-memory operand
//     reads per MMA and up to 40 % fewer bytes through L2.  Same products in the same order as the plain kernel: the
//     outputs are bit-identical.
// Work split: the (row tile, n-tile) pairs of the GEMM in row-major order are cut into gridDim.x contiguous runs of
// (nearly) equal length, one per CTA; a run that crosses into the next row tile reloads A there (1.8 us of latency, which
// is why the kernel only ties the plain one on 148 CTAs).  Its use is the NARROW launch -- one CTA per two row tiles, each
// A load amortised over all n-tiles -- of handles that run as execution lanes.
// Accumulators: two 128-column TMEM buffers (the epilogue of n-tile i overlaps the MMAs of n-tile i + 1; an accumulator is
// released as soon as it sits in registers).  Epilogue = bias (slice fetched one tile ahead) / ReLU / dropout / FP16 hi-lo
// split or fp32, 32 x 32 boxes through four rotating 2 KB shared tiles per warp, written by TMA stores.
#pragma once
#include "tip_umma.cuh"

namespace tip {

constexpr int AT_BN = 128;
constexpr int AT_PLANE_BYTES = 128 * 128;                  // one plane of a k-block: 128 rows x 128 B (A or W)
constexpr int AT_MAX_N_PER_UNIT = 1024;
// STAGES = depth of the operand ring.  CG = CTAs per work unit: 1, or 2 = a CTA PAIR (tcgen05 cta_group::2, M = 256): each CTA
// keeps its own 128-row A tile in its own tensor memory and stages only HALF of every W k-block (64 of the 128 rows), so
// the bytes a CTA pulls through its L2 port per n-tile drop from W + out = 192 KB to W / 2 + out = 128 KB -- the port
// (~78 GB/s per SM, shared by TMA loads and stores) is what paces this kernel.  Stage = one k-block of both planes
// (CG = 1: 32 KB) or 16 KB (CG = 2: one plane of an A k-block, or both planes of a W half k-block); per epilogue warp
// 8 KB of output staging (chunk outputs double-buffered), 4 KB when the ring takes 6 x 32 KB.
template <int STAGES, int CG = 1> struct AtmCfg {
    static constexpr int STAGE_BYTES = (CG == 2) ? AT_PLANE_BYTES : 2 * AT_PLANE_BYTES;
    static constexpr int STG_WARP_BYTES = (STAGES * STAGE_BYTES > 5 * 32768) ? 4096 : 8192;
    static constexpr int NBUF = STG_WARP_BYTES / 4096;     // 4 KB = one chunk's output (fp16 hi + lo tiles, or one fp32 tile)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + UM_EPI_WARPS * STG_WARP_BYTES + 2 * AT_BN * 4 /*bias slices*/ + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(CG == 1 || STAGES == 8, "pair kernel: 8 x 16 KB (the A tile's 8 plane k-blocks fill the ring once)");
};
constexpr int AT_TMEM_COLS = 512;                          // [0,128) / [128,256) accumulators, [256,384) A hi, [384,512) A lo
constexpr int AT_A_COL0 = 256;

namespace ptx {
// shared -> tensor memory: 128 lanes x 256 bit (lane = row of a K-major tile, 16 fp16 = 8 columns)
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b_pair(uint32_t taddr, uint64_t sdesc) {     // both CTAs: own shared memory -> own tensor memory
    asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
}  // namespace ptx

template <bool OUT_HALF, int AT_STAGES, int CG = 1>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_atm_gemm_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                     const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                     const __grid_constant__ CUtensorMap mapC0, const __grid_constant__ CUtensorMap mapC1,
                     int M, int N, int m_tile0, int m_tiles, Epi ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    using Cfg = AtmCfg<AT_STAGES, CG>;
    constexpr int AT_STAGE_BYTES = Cfg::STAGE_BYTES;
    uint8_t* staging = smem + AT_STAGES * AT_STAGE_BYTES;                                  // 8 x (8 | 4) KB
    float* sbias = reinterpret_cast<float*>(staging + UM_EPI_WARPS * Cfg::STG_WARP_BYTES); // [2][128]: this and the next n-tile's slice
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 2 * AT_BN);
    uint64_t* full_bar = bars;                       // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + AT_STAGES;          // [STAGES] MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * AT_STAGES;      // [2] MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * AT_STAGES + 2; // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * AT_STAGES + 4);
    volatile uint64_t* seed_slot = reinterpret_cast<volatile uint64_t*>(bars + 2 * AT_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int num_kb = E / UM_BK;                // 4 k-blocks of 64
    // this CTA's (pair's) run [u0, u1) of (row tile [pair], n-tile) pairs, u = mt * n_tiles + nt
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;       // rank 0 = leader: issues the pair's copies and MMAs
    const int grp = blockIdx.x / CG, n_grp = gridDim.x / CG;
    const int n_tiles = N / AT_BN;
    const int total = (m_tiles / CG) * n_tiles;
    const int u0 = (int)(((long long)grp * total) / n_grp);
    const int u1 = (int)(((long long)(grp + 1) * total) / n_grp);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mapA_hi); ptx::prefetch_tmap(&mapA_lo);
        ptx::prefetch_tmap(&mapB_hi); ptx::prefetch_tmap(&mapB_lo);
        for (int s = 0; s < AT_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        // (pair: the accumulators are released by the epilogue warps of BOTH CTAs, on the leader's barrier)
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], UM_EPI_WARPS * CG); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { if constexpr (CG == 2) ptx::tmem_alloc2(tmem_slot, AT_TMEM_COLS); else ptx::tmem_alloc(tmem_slot, AT_TMEM_COLS); }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) { cluster_arrive(); cluster_wait(); }   // the peer's barriers exist before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();
    if (ep.pdl_early) griddep_launch();
#define AT_TS(i) do { if (ep.tbuf && blockIdx.x == 0 && lane == 0) ep.tbuf[i] = ptx::globaltimer_ns(); } while (0)
    if (warp == 2) AT_TS(0);

    if (warp == 0) {
        // ================= TMA producer: the A tile's four k-blocks, then W for every n-tile of the unit =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = u0; u < u1; ++u) {
                const int mt = u / n_tiles, nt = u - mt * n_tiles;
                if (u == u0 || nt == 0) {                       // a new row tile: its A k-blocks go through the ring first
                    const int m0 = (m_tile0 + mt * CG + (int)cta_rank) * UM_BM;
                    if constexpr (CG == 2) {
                        // one plane of one k-block (16 KB) per stage; both CTAs' bytes are counted on the LEADER's barrier
                        for (int kp = 0; kp < 2 * num_kb; ++kp) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                            const uint32_t fb = map_to_cta(ptx::smem_u32(&full_bar[stage]), 0u);
                            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * AT_STAGE_BYTES);
                            ptx::tma_load_2d_pair(smem + stage * AT_STAGE_BYTES, (kp & 1) ? &mapA_lo : &mapA_hi, fb, (kp >> 1) * UM_BK, m0);
                            if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    } else {
                        for (int kb = 0; kb < num_kb; ++kb) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* s = smem + stage * AT_STAGE_BYTES;
                            ptx::mbar_expect_tx(&full_bar[stage], AT_STAGE_BYTES);
                            ptx::tma_load_2d(s, &mapA_hi, &full_bar[stage], kb * UM_BK, m0);
                            ptx::tma_load_2d(s + AT_PLANE_BYTES, &mapA_lo, &full_bar[stage], kb * UM_BK, m0);
                            if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* s = smem + stage * AT_STAGE_BYTES;
                    if constexpr (CG == 2) {
                        // this CTA's 64 rows of the W k-block, hi | lo (8 KB each)
                        const uint32_t fb = map_to_cta(ptx::smem_u32(&full_bar[stage]), 0u);
                        const int n0 = nt * AT_BN + (int)cta_rank * (AT_BN / 2);
                        if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * AT_STAGE_BYTES);
                        ptx::tma_load_2d_pair(s, &mapB_hi, fb, kb * UM_BK, n0);
                        ptx::tma_load_2d_pair(s + AT_STAGE_BYTES / 2, &mapB_lo, fb, kb * UM_BK, n0);
                    } else {
                        ptx::mbar_expect_tx(&full_bar[stage], AT_STAGE_BYTES);
                        ptx::tma_load_2d(s, &mapB_hi, &full_bar[stage], kb * UM_BK, nt * AT_BN);
                        ptx::tma_load_2d(s + AT_PLANE_BYTES, &mapB_lo, &full_bar[stage], kb * UM_BK, nt * AT_BN);
                    }
                    if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: A shared -> tensor memory, then A from tensor memory x W from the ring =================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(UM_BM * CG, AT_BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int u = u0; u < u1; ++u, ++it) {
                const int nt = u % n_tiles;
                if (u == u0 || nt == 0) {
                    // A tile: shared -> tensor memory (tcgen05.cp runs in issue order behind the MMAs that still read the
                    // previous row tile's A); pair: one instruction copies in BOTH CTAs, each from its own shared memory
                    if constexpr (CG == 2) {
                        for (int kp = 0; kp < 2 * num_kb; ++kp) {
                            ptx::mbar_wait(&full_bar[stage], phase);
                            if (u == u0 && kp == 0) AT_TS(1);
                            ptx::tc_fence_after();
                            const uint64_t ad = umma_smem_desc(ptx::smem_u32(smem + stage * AT_STAGE_BYTES));
#pragma unroll
                            for (int k = 0; k < UM_BK / 16; ++k)
                                ptx::tmem_cp_128x256b_pair(tmem_base + (uint32_t)(AT_A_COL0 + (kp & 1) * 128 + (kp >> 1) * 32 + k * 8),
                                                           ad + (uint64_t)((k * 32) >> 4));
                            ptx::umma_commit_pair(&empty_bar[stage]);
                            if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    } else {
                        for (int kb = 0; kb < num_kb; ++kb) {
                            ptx::mbar_wait(&full_bar[stage], phase);
                            if (u == u0 && kb == 0) AT_TS(1);
                            ptx::tc_fence_after();
                            const uint32_t sa = ptx::smem_u32(smem + stage * AT_STAGE_BYTES);
                            const uint64_t a_hi = umma_smem_desc(sa), a_lo = umma_smem_desc(sa + AT_PLANE_BYTES);
#pragma unroll
                            for (int k = 0; k < UM_BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                const uint32_t col = (uint32_t)(AT_A_COL0 + kb * 32 + k * 8);           // 16 k = 8 columns
                                ptx::tmem_cp_128x256b(tmem_base + col, a_hi + adv);
                                ptx::tmem_cp_128x256b(tmem_base + col + 128u, a_lo + adv);
                            }
                            ptx::umma_commit(&empty_bar[stage]);               // slot free once the copies have read it
                            if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                    if (u == u0) AT_TS(2);
                }
                const int as = it & 1;
                ptx::mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                if (it < 8) AT_TS(40 + it);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * AT_BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    if (it == 0 && kb == 0) AT_TS(3);
                    ptx::tc_fence_after();
                    const uint32_t sb = ptx::smem_u32(smem + stage * AT_STAGE_BYTES);
                    // pair: each CTA holds 64 of the tile's 128 W rows (hi at +0, lo at +8 KB) at this offset
                    const uint64_t b_hi = umma_smem_desc(sb), b_lo = umma_smem_desc(sb + (CG == 2 ? AT_STAGE_BYTES / 2 : AT_PLANE_BYTES));
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        const uint32_t ah = tmem_base + (uint32_t)(AT_A_COL0 + kb * 32 + k * 8);
                        const uint32_t al = ah + 128u;
                        if constexpr (CG == 2) {
                            ptx::umma_f16_ts_pair(d_tmem, al, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            ptx::umma_f16_ts_pair(d_tmem, ah, b_lo + adv, idesc, 1u);
                            ptx::umma_f16_ts_pair(d_tmem, ah, b_hi + adv, idesc, 1u);
                        } else {
                            ptx::umma_f16_ts(d_tmem, al, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            ptx::umma_f16_ts(d_tmem, ah, b_lo + adv, idesc, 1u);
                            ptx::umma_f16_ts(d_tmem, ah, b_hi + adv, idesc, 1u);
                        }
                    }
                    if constexpr (CG == 2) ptx::umma_commit_pair(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                }
                if constexpr (CG == 2) ptx::umma_commit_pair(&tfull_bar[as]); else ptx::umma_commit(&tfull_bar[as]);
                if (it < 8) AT_TS(4 + it);
            }
        }
    } else {
        // ================= warps 2..9: epilogue; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const float asc = ep.acc_scale ? __ldg(ep.acc_scale) : 1.f;
        const float osc = OUT_HALF ? ACT_SCALE : 1.f;
        const float dinv = ep.drop_thr ? ep.drop_inv : 1.f;
        const float relu_floor = ep.relu ? 0.f : -INFINITY;
        const float sc = asc * osc * dinv;
        uint8_t* sbuf = staging + (warp - 2) * Cfg::STG_WARP_BYTES;
        // bias slice of an n-tile (pre-multiplied by the output scale [and 1/(1-p)]): fetched one tile ahead into a register,
        // parked in the [2][128] shared slices at the top of its tile
        const int et = (int)threadIdx.x - 64;
        float bnext = (et < AT_BN && u0 < u1) ? __ldg(ep.bias + (u0 % n_tiles) * AT_BN + et) * osc * dinv : 0.f;
        if (et == 0) *seed_slot = ep.drop_thr ? site_seed(ep.seed_ptr, ep.seed) : 0ull;
        int it = 0;
        for (int u = u0; u < u1; ++u, ++it) {
            const int as = it & 1;
            const int mt = u / n_tiles, nt = u - mt * n_tiles;
            const int n0 = nt * AT_BN;
            const int rbase = (m_tile0 + mt * CG + (int)cta_rank) * UM_BM + quarter * 32;
            if (et < AT_BN) {
                sbias[as * AT_BN + et] = bnext;
                if (u + 1 < u1) bnext = __ldg(ep.bias + ((u + 1) % n_tiles) * AT_BN + et) * osc * dinv;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // (slice `as` was last read two tiles ago: every warp is past that)
            ptx::mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            if (warp == 2 && it < 8) AT_TS(12 + it);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * AT_BN + half * 64);
            float v[2][32];
            ptx::tmem_ld32_nowait(t_acc, v[0]);
            ptx::tmem_ld32_nowait(t_acc + 32, v[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                           // the accumulator sits in registers: hand it back
                if constexpr (CG == 2) ptx::mbar_arrive_cluster(map_to_cta(ptx::smem_u32(&tempty_bar[as]), 0u));
                else ptx::mbar_arrive(&tempty_bar[as]);
            }
            if (warp == 2 && it < 8) AT_TS(20 + it);
            const float* bs = sbias + as * AT_BN + half * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int colb = n0 + half * 64 + c * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b = *reinterpret_cast<const float4*>(bs + c * 32 + 4 * j4);       // broadcast
                    v[c][4 * j4 + 0] = fmaxf(fmaf(v[c][4 * j4 + 0], sc, b.x), relu_floor);
                    v[c][4 * j4 + 1] = fmaxf(fmaf(v[c][4 * j4 + 1], sc, b.y), relu_floor);
                    v[c][4 * j4 + 2] = fmaxf(fmaf(v[c][4 * j4 + 2], sc, b.z), relu_floor);
                    v[c][4 * j4 + 3] = fmaxf(fmaf(v[c][4 * j4 + 3], sc, b.w), relu_floor);
                }
                if (ep.drop_thr) {                                      // dropout on the output (ff1): element index = row * N + col
                    const uint64_t g0 = ((uint64_t)(rbase + lane) * N + colb) >> 2;
                    const uint64_t dseed = *seed_slot;
                    const uint32_t thr_hi = ep.drop_thr << 16;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const uint64_t h = hash_u64(dseed, g0 + j4);
                        const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
                        if ((hl << 16) < thr_hi) v[c][4 * j4 + 0] = 0.f;
                        if (hl < thr_hi) v[c][4 * j4 + 1] = 0.f;
                        if ((hh << 16) < thr_hi) v[c][4 * j4 + 2] = 0.f;
                        if (hh < thr_hi) v[c][4 * j4 + 3] = 0.f;
                    }
                }
                if constexpr (OUT_HALF) {
                    // [32 rows][64 B] tiles (hi, lo) of chunk c, 64B-swizzled: 16-byte chunk ^= (row >> 1) & 3.  Four tiles per
                    // warp in rotation (c0 hi, c0 lo, c1 hi, c1 lo), one bulk group each: a tile is rewritten four groups later
                    const int sw = (lane >> 1) & 3;
                    uint4 uh[4], ul[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float h0, h1, l0, l1;
                        __half2 t;
                        veltkamp11(v[c][8 * j + 0], h0, l0); veltkamp11(v[c][8 * j + 1], h1, l1);
                        t = __floats2half2_rn(h0, h1); uh[j].x = *reinterpret_cast<uint32_t*>(&t);
                        t = __floats2half2_rn(l0, l1); ul[j].x = *reinterpret_cast<uint32_t*>(&t);
                        veltkamp11(v[c][8 * j + 2], h0, l0); veltkamp11(v[c][8 * j + 3], h1, l1);
                        t = __floats2half2_rn(h0, h1); uh[j].y = *reinterpret_cast<uint32_t*>(&t);
                        t = __floats2half2_rn(l0, l1); ul[j].y = *reinterpret_cast<uint32_t*>(&t);
                        veltkamp11(v[c][8 * j + 4], h0, l0); veltkamp11(v[c][8 * j + 5], h1, l1);
                        t = __floats2half2_rn(h0, h1); uh[j].z = *reinterpret_cast<uint32_t*>(&t);
                        t = __floats2half2_rn(l0, l1); ul[j].z = *reinterpret_cast<uint32_t*>(&t);
                        veltkamp11(v[c][8 * j + 6], h0, l0); veltkamp11(v[c][8 * j + 7], h1, l1);
                        t = __floats2half2_rn(h0, h1); uh[j].w = *reinterpret_cast<uint32_t*>(&t);
                        t = __floats2half2_rn(l0, l1); ul[j].w = *reinterpret_cast<uint32_t*>(&t);
                    }
                    uint8_t* th = sbuf + (Cfg::NBUF == 2 ? c * 4096 : 0);
                    uint8_t* tl = th + 2048;
                    if constexpr (Cfg::NBUF == 2) {
                        // four tiles per warp in rotation (c0 hi, c0 lo, c1 hi, c1 lo), one bulk group each
                        if (lane == 0) ptx::bulk_wait_read<2>();          // hi AND lo tile of this chunk's previous use have been read
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(th + lane * 64 + ((j ^ sw) << 4)) = uh[j];
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(tl + lane * 64 + ((j ^ sw) << 4)) = ul[j];
                        ptx::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapC0, th, colb, rbase);
                            ptx::bulk_commit();
                            ptx::tma_store_2d(&mapC1, tl, colb, rbase);
                            ptx::bulk_commit();
                        }
                    } else {
                        // two tiles (hi, lo): while one plane's box is still being read out the other can be rewritten
                        if (lane == 0) ptx::bulk_wait_read<1>();          // hi tile free
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(th + lane * 64 + ((j ^ sw) << 4)) = uh[j];
                        ptx::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapC0, th, colb, rbase);
                            ptx::bulk_commit();
                            ptx::bulk_wait_read<1>();                     // lo tile free
                        }
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(tl + lane * 64 + ((j ^ sw) << 4)) = ul[j];
                        ptx::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapC1, tl, colb, rbase);
                            ptx::bulk_commit();
                        }
                    }
                } else {
                    uint8_t* tf = sbuf + (Cfg::NBUF == 2 ? c * 4096 : 0);   // one [32 rows][128 B] fp32 tile per chunk, 128B-swizzled
                    if (lane == 0) ptx::bulk_wait_read<Cfg::NBUF - 1>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(tf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
                    ptx::fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&mapC0, tf, colb, rbase);
                        ptx::bulk_commit();
                    }
                }
            }
            if (warp == 2 && it < 8) AT_TS(28 + it);
        }
    }
    if (!ep.pdl_early) griddep_launch();
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();
    if (warp == 2) AT_TS(36);
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) { cluster_arrive(); cluster_wait(); }   // neither CTA leaves while the pair's MMAs / arrivals can touch it
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc2(tmem_base, AT_TMEM_COLS); else ptx::tmem_dealloc(tmem_base, AT_TMEM_COLS);
    }
}

// grid_cap: CTAs of the launch (<= 0: one per SM); the (row tile, n-tile) pairs are dealt to them in contiguous runs
template <int STAGES>
inline void launch_atm_gemm_s(const UmmaMaps& mp, const UmmaOperand& A, const UmmaOperand& B, const UmmaOutput& C,
                              int M, int N, int m_tile0, int m_tiles, int grid_cap, const Epi& ep, cudaStream_t st) {
    static bool attrs = false;
    if (!attrs) {
        cudaFuncSetAttribute(umma_atm_gemm_kernel<true, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtmCfg<STAGES>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_atm_gemm_kernel<false, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtmCfg<STAGES>::SMEM_BYTES);
        attrs = true;
    }
    pdl_kind() = 1;
    const dim3 grid(std::min(grid_cap > 0 ? grid_cap : mp.num_sms, m_tiles * (N / AT_BN)));
    if (ep.out_lo)
        launch_k(umma_atm_gemm_kernel<true, STAGES>, grid, dim3(UM_THREADS), AtmCfg<STAGES>::SMEM_BYTES, st, A.hi, A.lo, B.hi, B.lo, C.c0, C.c1, M, N, m_tile0, m_tiles, ep);
    else
        launch_k(umma_atm_gemm_kernel<false, STAGES>, grid, dim3(UM_THREADS), AtmCfg<STAGES>::SMEM_BYTES, st, A.hi, A.lo, B.hi, B.lo, C.c0, C.c1, M, N, m_tile0, m_tiles, ep);
}
// CTA pairs: B64 = the weight planes with 64-row boxes; m_tiles must be even; grid_cap counts CTAs (rounded down to pairs)
inline void launch_atm_pair_gemm(const UmmaMaps& mp, const UmmaOperand& A, const UmmaOperand& B64, const UmmaOutput& C,
                                 int M, int N, int m_tile0, int m_tiles, int grid_cap, const Epi& ep, cudaStream_t st) {
    static bool attrs = false;
    if (!attrs) {
        cudaFuncSetAttribute(umma_atm_gemm_kernel<true, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtmCfg<8, 2>::SMEM_BYTES);
        cudaFuncSetAttribute(umma_atm_gemm_kernel<false, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtmCfg<8, 2>::SMEM_BYTES);
        attrs = true;
    }
    pdl_kind() = 1;
    const int units = (m_tiles / 2) * (N / AT_BN);
    const int pairs = std::max(1, std::min((grid_cap > 0 ? grid_cap : mp.num_sms) / 2, units));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(UM_THREADS);
    cfg.dynamicSmemBytes = AtmCfg<8, 2>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (ep.out_lo)
        cudaLaunchKernelEx(&cfg, umma_atm_gemm_kernel<true, 8, 2>, A.hi, A.lo, B64.hi, B64.lo, C.c0, C.c1, M, N, m_tile0, m_tiles, ep);
    else
        cudaLaunchKernelEx(&cfg, umma_atm_gemm_kernel<false, 8, 2>, A.hi, A.lo, B64.hi, B64.lo, C.c0, C.c1, M, N, m_tile0, m_tiles, ep);
}
inline void launch_atm_gemm(const UmmaMaps& mp, const UmmaOperand& A, const UmmaOperand& B, const UmmaOutput& C,
                            int M, int N, int m_tile0, int m_tiles, int grid_cap, const Epi& ep, cudaStream_t st) {
    static const int stages = getenv("TIP_ATM_STAGES") ? atoi(getenv("TIP_ATM_STAGES")) : 5;
    if (stages >= 6) launch_atm_gemm_s<6>(mp, A, B, C, M, N, m_tile0, m_tiles, grid_cap, ep, st);
    else if (stages == 5) launch_atm_gemm_s<5>(mp, A, B, C, M, N, m_tile0, m_tiles, grid_cap, ep, st);
    else launch_atm_gemm_s<4>(mp, A, B, C, M, N, m_tile0, m_tiles, grid_cap, ep, st);
}

}  // namespace tip
