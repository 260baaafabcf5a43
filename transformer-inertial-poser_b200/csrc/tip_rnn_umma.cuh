// tanh RNN recurrence on the tensor cores (reference simple_transformer_with_state.py:95-99):
//   h_t = tanh(gi[:, t, :] + W_hh h_{t-1}),  h_0 = 0,  gi = x W_ih^T + b_ih + b_hh (previous GEMM).
//
// The recurrence is the path's only serial chain.  A thread-block cluster of 8 CTAs carries up to
// RU_N = 24 windows through all L steps:
//   * CTA c owns hidden units [64c, 64c+64).  Its slice of W_hh (FP16 hi/lo planes of s_w * W_hh,
//     128 KB) is TMA-loaded ONCE into 128B-swizzled shared memory and stays resident as the A
//     operand; nothing is re-streamed from L2 inside the time loop.  The hi and lo rows are STACKED
//     along M (128 rows per k-block; TMEM quarter q holds [hi rows of units 16q..16q+15 | lo rows
//     of the same units]) ...
//   * ... and the B operand stacks the FP16 hi/lo planes of 16*h_{t-1} along N (48 rows per k-block:
//     24 windows hi, 24 windows lo), so ONE tcgen05.mma.kind::f16 (M=128, N=48, K=16) per k-step
//     yields all partial products of the error-compensated split (hi*hi, hi*lo, lo*hi, lo*lo): 32
//     MMAs per time step instead of 96.  The issue rate of the single MMA thread, not the tensor
//     pipe, bounds this tiny GEMM, so fewer/larger instructions are what matters.
//   * eight epilogue warps read the accumulator (tcgen05.ld), fold the four partial products (two
//     column groups + a lane^16 shuffle), add gi, apply tanh, split to FP16 hi/lo and write this
//     CTA's k-block of the next B operand in the swizzled layout;
//   * that k-block (6 KB) is pushed to the 7 peer CTAs with cp.async.bulk.shared::cluster, the bytes
//     counted on an mbarrier in each destination, which is also what the MMA issuer of the next step
//     waits on -- no cluster-wide barrier in the loop.  B is double-buffered over the step parity.
#pragma once
#include "tip_umma.cuh"

namespace tip {

constexpr int RU_CTAS = 8;
constexpr int RU_UNITS = R / RU_CTAS;                 // 64 hidden units per CTA (128 stacked A rows)
constexpr int RU_N = 24;                              // windows per cluster pass (48 stacked B rows)
constexpr int RU_EPI_WARPS = 8;
constexpr int RU_THREADS = 64 + 32 * RU_EPI_WARPS;    // warp 0: MMA issuer / TMEM owner; warps 1..8: epilogue; warp 9: second MMA issuer
constexpr int RU_A_KB_BYTES = 2 * RU_UNITS * 128;     // 16 KB: one k-block (64 k) of the stacked W_hh slice
constexpr int RU_A_BYTES = 8 * RU_A_KB_BYTES;         // 128 KB
constexpr int RU_B_KB_BYTES = 2 * RU_N * 128;         // 6 KB: one k-block of stacked h (hi rows, then lo rows)
constexpr int RU_B_BYTES = 8 * RU_B_KB_BYTES;         // 48 KB per parity
constexpr int RU_SMEM_BYTES = RU_A_BYTES + 2 * RU_B_BYTES + 256;
constexpr uint32_t RU_PUSH_BYTES = (RU_CTAS - 1) * RU_B_KB_BYTES;         // what the 7 peers deliver per step
// ISS issuing threads (template parameter of the kernel): with ISS = 2 a second warp issues the MMAs of k-blocks 4..7
// into a second accumulator (64 TMEM columns further) while warp 0 issues k-blocks 0..3; each commits on acc_full
// (count ISS) and the epilogue adds the partial sums.  (An earlier experiment with 4 accumulators fed by ONE thread
// left the issue phase at 1.0 us -- 31 ns per M=128 x N=48 x K=16 MMA -- and so did the A-in-TMEM variant: the
// single issuing thread paces these small MMAs, not the accumulate dependency or the operand read.)
constexpr int RU_NACC = 2;                           // accumulators allocated
constexpr int RU_TMEM_COLS = 64 * RU_NACC;
// A_TMEM variant: the W_hh slice lives in TENSOR MEMORY instead of shared memory (128 lanes x 256 32-bit
// columns = 128 stacked rows x 512 fp16 k), so each MMA reads only the small h operand from shared memory;
// the shared-memory read of the 4 KB A slice per MMA is what bounds the issue rate of the SS form.
constexpr int RU_TMEM_COLS_A = 512;
constexpr int RU_A_COL0 = 64 * RU_NACC;              // first TMEM column of the A operand (the accumulators come first)

// ST_ASYNC: the 6 KB tile goes to the 7 peers as 16-byte st.async stores issued by all epilogue threads (bytes counted
// on the destination's mbarrier like the bulk copies) instead of 7 cp.async.bulk pushes issued by 7 threads.
// PIPE: one mbarrier per (parity, source CTA) instead of one per parity; every CTA pushes its tile to the peers in ring
// order (rank+1 first) from ONE thread, so the tiles reach a CTA staggered in time (rank-1's first), and the issuer
// consumes the k-blocks in that order: the MMAs of a step overlap the DSMEM all-gather instead of following it.
// NW windows per cluster (24 or 20): the all-gather moves NW * 64 units * 4 bytes per CTA pair and step and is what
// paces a step, so when the batch fits the co-resident clusters either way the smaller group is faster (the MMA keeps
// N = 48: B-tile rows [0, NW) = hi planes, [NW, 2 NW) = lo planes, the rest unused).
template <bool A_TMEM, bool ST_ASYNC = false, int ISS = 1, bool PIPE = false, int NW = RU_N>
__global__ void __cluster_dims__(RU_CTAS, 1, 1) __launch_bounds__(RU_THREADS, 1)
rnn_umma_kernel(const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                const float* __restrict__ gi, __half* __restrict__ hs_hi, __half* __restrict__ hs_lo,
                const float* __restrict__ acc_scale, int B, int L, unsigned long long* tbuf,
                const __half* __restrict__ whh_hi, const __half* __restrict__ whh_lo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* sA = smem;                                    // [8 kb][128 stacked rows x 128 B]
    uint8_t* sB = smem + RU_A_BYTES;                       // [2 parities][8 kb][48 stacked rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RU_A_BYTES + 2 * RU_B_BYTES);
    uint64_t* w_full = bars;                               // W_hh slice landed
    uint64_t* h_full = bars + 1;                           // [2] h_t complete in a parity buffer (local arrive + 7 pushes)
    uint64_t* acc_full = bars + 3;                         // MMAs of the step retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    uint64_t* hk_full = bars + 8;                          // PIPE: [2 parities][8 source CTAs]
    static_assert(NW == 24 || NW == 20, "window group: 24 or 20");
    constexpr int WH = NW / 2, WL = NW / 4;                // windows per column half (warp group) / per lane
    constexpr uint32_t PUSH_KB = 2 * NW * 128;             // bytes of a k-block tile that carry data (pushed to the peers)
    static_assert(!PIPE || (!ST_ASYNC && ISS == 1), "PIPE builds on the bulk-copy, single-issuer kernel");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_clusters = gridDim.x / RU_CTAS, cluster_id = blockIdx.x / RU_CTAS;
    const uint32_t sB_s = ptx::smem_u32(sB);

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&mapW_hi); ptx::prefetch_tmap(&mapW_lo);
        ptx::mbar_init(w_full, 1);
        ptx::mbar_init(&h_full[0], 1);
        ptx::mbar_init(&h_full[1], 1);
        ptx::mbar_init(acc_full, ISS);
        if constexpr (PIPE) for (int i = 0; i < 2 * RU_CTAS; ++i) ptx::mbar_init(&hk_full[i], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) ptx::tmem_alloc(tmem_slot, A_TMEM ? RU_TMEM_COLS_A : RU_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (A_TMEM) {
        // resident A operand in tensor memory: TMEM lane 32q + l holds stacked row (q, l) = [hi rows of units
        // 16q..16q+15 | lo rows of the same units]; 32-bit column c holds k = 2c, 2c+1 (K-major, packed pairs)
        if (warp >= 1 && warp <= RU_EPI_WARPS) {
            const int q = warp & 3, ch = (warp - 1) >> 2;
            const int unit = (int)rank * RU_UNITS + q * 16 + (lane & 15);
            const __half* src = ((lane >> 4) ? whh_lo : whh_hi) + (size_t)unit * R + ch * 256;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i * 64) + j);
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(RU_A_COL0 + ch * 128 + i * 32);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                    ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                    : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        ptx::tc_fence_before();
        __syncthreads();
        ptx::tc_fence_after();
    }
    if (!A_TMEM && threadIdx.x == 0) {
        // resident A operand.  Stacked row order of a k-block: for q = 0..3: 16 hi rows then 16 lo rows of
        // units 16q..16q+15, so both partial rows of a unit sit in the same TMEM lane quarter.
        ptx::mbar_expect_tx(w_full, RU_A_BYTES);
        for (int kb = 0; kb < 8; ++kb)
            for (int q = 0; q < 4; ++q) {
                uint8_t* dst = sA + kb * RU_A_KB_BYTES + q * 32 * 128;
                ptx::tma_load_2d(dst, &mapW_hi, w_full, kb * 64, (int)rank * RU_UNITS + q * 16);
                ptx::tma_load_2d(dst + 16 * 128, &mapW_lo, w_full, kb * 64, (int)rank * RU_UNITS + q * 16);
            }
    }
    cluster_arrive();          // every CTA's barriers are initialised before any peer pushes into it
    cluster_wait();
    griddep_wait();            // PDL: the W_hh slice (a weight) was requested above, while rnn_ih was still running
    griddep_launch();

    uint32_t hpar0 = 0, hpar1 = 0; // phase parities of h_full[0/1] (MMA issuer)
    uint32_t apar = 0;             // phase parity of acc_full (epilogue threads)

    for (int rb = cluster_id; rb * NW < B; rb += n_clusters) {
        const int b0 = rb * NW;
        if (warp == 0 || warp == 1 + RU_EPI_WARPS) {
            // ================= MMA issuer(s): warp 0 -> k-blocks [0, 8 / ISS), warp 9 -> the rest =================
            const int iss = (warp == 0) ? 0 : 1;
            if (lane == 0 && iss < ISS) {
                constexpr uint32_t idesc = umma_idesc_f16(2 * RU_UNITS, 2 * RU_N);
                if constexpr (!A_TMEM) ptx::mbar_wait(w_full, 0);   // completes once; later waits return immediately
                const uint32_t a0 = ptx::smem_u32(sA);
                const uint32_t d_acc = tmem_base + (uint32_t)(iss * 64);
                const int kb0 = iss * (8 / ISS), kb1 = kb0 + 8 / ISS;
                for (int t = 1; t < L; ++t) {              // step 0 has h = 0: no product
                    const int cur = t & 1;
                    const uint32_t hpar = cur ? hpar1 : hpar0;
                    if (cur) hpar1 ^= 1; else hpar0 ^= 1;
                    if constexpr (!PIPE) {
                        ptx::mbar_wait(&h_full[cur], hpar);
                        if (tbuf && blockIdx.x == 0 && iss == 0 && (t == 20 || t == 21)) tbuf[(t - 20) * 4 + 0] = ptx::globaltimer_ns();
                        ptx::tc_fence_after();
                    }
                    const uint32_t bb = sB_s + (uint32_t)cur * RU_B_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 8 / ISS; ++kk) {
                        int kb = kb0 + kk;
                        if constexpr (PIPE) {
                            kb = (int)((rank - (uint32_t)kk) & (RU_CTAS - 1));      // arrival order: own tile, then rank-1, rank-2, ...
                            ptx::mbar_wait(&hk_full[cur * RU_CTAS + kb], hpar);
                            if (kk == 0 && tbuf && blockIdx.x == 0 && (t == 20 || t == 21)) tbuf[(t - 20) * 4 + 0] = ptx::globaltimer_ns();
                            ptx::tc_fence_after();
                        }
                        const uint64_t ad = umma_smem_desc(a0 + kb * RU_A_KB_BYTES);
                        const uint64_t bd = umma_smem_desc(bb + kb * RU_B_KB_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            const uint32_t acc = (kk | k) ? 1u : 0u;
                            if constexpr (A_TMEM) {
                                const uint32_t a_t = tmem_base + (uint32_t)(RU_A_COL0 + (kb * 4 + k) * 8);   // 16 k = 8 columns
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\t"
                                    "setp.ne.b32 p, %4, 0;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                    ::"r"(d_acc), "r"(a_t), "l"(bd + adv), "r"(idesc), "r"(acc) : "memory");
                            } else {
                                ptx::umma_f16(d_acc, ad + adv, bd + adv, idesc, acc);
                            }
                        }
                    }
                    (void)kb1;
                    ptx::umma_commit(acc_full);
                    if (tbuf && blockIdx.x == 0 && iss == 0 && t == 20) tbuf[1] = ptx::globaltimer_ns();
                }
            }
        } else {
            // ===== epilogue: warps 1..8; TMEM quarter = warp % 4, window half ch = (warp-1)/4 =====
            const int q = warp & 3;
            const int ch = (warp - 1) >> 2;
            const int lh = lane >> 4;                            // 0: hi-row lane, 1: lo-row lane of the same unit
            const int ul = q * 16 + (lane & 15);                 // unit within the CTA
            const int unit = (int)rank * RU_UNITS + ul;
            const int nb = WH * ch + WL * lh;                    // first of the WL windows this lane finishes
            const float asc = __ldg(acc_scale);                  // 1 / (s_w * 16)
            const int et = (int)threadIdx.x - 32;                // 0..255 among the epilogue threads
            for (int t = 0; t < L; ++t) {
                const int nxt = (t + 1) & 1;
                float g[WL];                                     // gi of this lane's outputs (latency overlaps the MMA wait)
#pragma unroll
                for (int j = 0; j < WL; ++j)
                    g[j] = __ldg(gi + ((size_t)min(b0 + nb + j, B - 1) * L + t) * R + unit);
                float pre[WL];
                if (t > 0) {
                    ptx::mbar_wait(acc_full, apar);
                    apar ^= 1;
                    if (tbuf && blockIdx.x == 0 && t == 20 && et == 0) tbuf[2] = ptx::globaltimer_ns();
                    ptx::tc_fence_after();
                    // columns [WH ch, WH ch + WH) (x h_hi) and [NW + WH ch, ...) (x h_lo) of this lane's stacked row
                    float s[WH];
#pragma unroll
                    for (int j = 0; j < WH; ++j) s[j] = 0.f;
                    const uint32_t ta0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(WH * ch);
#pragma unroll
                    for (int a = 0; a < ISS; ++a) {
                        uint32_t r0[12], r1[12];
                        const uint32_t ta = ta0 + (uint32_t)(a * 64);
                        // WH = 12: x4 + x8 columns; WH = 10: x8 + x2
                        if constexpr (WH == 12) {
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]) : "r"(ta) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11])
                                         : "r"(ta + 4) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]) : "r"(ta + NW) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11])
                                         : "r"(ta + NW + 4) : "memory");
                        } else {
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7])
                                         : "r"(ta) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                                         : "=r"(r0[8]), "=r"(r0[9]) : "r"(ta + 8) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7])
                                         : "r"(ta + NW) : "memory");
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                                         : "=r"(r1[8]), "=r"(r1[9]) : "r"(ta + NW + 8) : "memory");
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < WH; ++j) s[j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    }
                    ptx::tc_fence_before();
                    // fold the hi-row and lo-row lanes of a unit; lane half lh keeps windows WL lh .. WL lh + WL - 1
#pragma unroll
                    for (int j = 0; j < WL; ++j) {
                        const float mine = lh ? s[j + WL] : s[j];
                        const float send = lh ? s[j] : s[j + WL];
                        pre[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < WL; ++j) pre[j] = 0.f;
                }
                // h_t -> this CTA's k-block of the next B operand (swizzled rows: windows hi, then windows lo)
                uint8_t* tile = sB + nxt * RU_B_BYTES + (int)rank * RU_B_KB_BYTES;
#pragma unroll
                for (int j = 0; j < WL; ++j) {
                    const int n = nb + j;
                    const float h = tanhf(fmaf(pre[j], asc, g[j]));
                    __half hi, lo;
                    half_split(h * ACT_SCALE, hi, lo);
                    const int off_h = ((((ul >> 3) ^ (n & 7))) << 4) + (ul & 7) * 2;
                    const int off_l = ((((ul >> 3) ^ ((NW + n) & 7))) << 4) + (ul & 7) * 2;
                    *reinterpret_cast<__half*>(tile + n * 128 + off_h) = hi;
                    *reinterpret_cast<__half*>(tile + (NW + n) * 128 + off_l) = lo;
                }
                ptx::fence_async_smem();                              // generic writes -> async proxy (MMA, bulk copies)
                asm volatile("bar.sync 1, 256;" ::: "memory");        // the whole k-block is written
                if (tbuf && blockIdx.x == 0 && t == 20 && et == 0) tbuf[3] = ptx::globaltimer_ns();
                if (t + 1 < L) {
                    // publish: local arrival + the 7 peers' bytes complete h_full[nxt] in every CTA
                    if constexpr (PIPE) {
                        if (et == 32) ptx::mbar_arrive(&hk_full[nxt * RU_CTAS + rank]);                 // own tile: written above
                        else if (et >= 33 && et < 33 + RU_CTAS && (uint32_t)(et - 33) != rank)
                            ptx::mbar_expect_tx(&hk_full[nxt * RU_CTAS + (et - 33)], PUSH_KB);    // the tile source (et-33) will push
                        if (et == 0) {
                            const uint32_t src = ptx::smem_u32(tile);
                            const uint32_t bar = ptx::smem_u32(&hk_full[nxt * RU_CTAS + rank]);
#pragma unroll
                            for (uint32_t c = 1; c < RU_CTAS; ++c) {
                                const uint32_t dstc = (rank + c) & (RU_CTAS - 1);
                                dsmem_bulk_push(map_to_cta(src, dstc), src, PUSH_KB, map_to_cta(bar, dstc));
                            }
                        }
                    } else {
                    if (et == 32) ptx::mbar_expect_tx(&h_full[nxt], (RU_CTAS - 1) * PUSH_KB);
                    if constexpr (ST_ASYNC) {
                        const uint32_t src = ptx::smem_u32(tile);
                        const uint32_t bar = ptx::smem_u32(&h_full[nxt]);
                        for (int i = et; i < (int)PUSH_KB / 16; i += 32 * RU_EPI_WARPS) {
                            const uint4 v = *reinterpret_cast<const uint4*>(tile + i * 16);
#pragma unroll
                            for (uint32_t c = 1; c < RU_CTAS; ++c) {
                                const uint32_t dstc = (rank + c) & (RU_CTAS - 1);
                                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                             ::"r"(map_to_cta(src + i * 16, dstc)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                                               "r"(map_to_cta(bar, dstc)) : "memory");
                            }
                        }
                    } else if (et < RU_CTAS && (uint32_t)et != rank) {
                        const uint32_t src = ptx::smem_u32(tile);
                        dsmem_bulk_push(map_to_cta(src, (uint32_t)et), src, PUSH_KB,
                                        map_to_cta(ptx::smem_u32(&h_full[nxt]), (uint32_t)et));
                    }
                    }
                }
                // hs[b, t, 64c .. 64c+63] (hi / lo planes) from the tile: 16-byte chunks, un-swizzled
                for (int i = et; i < 2 * NW * 8; i += 32 * RU_EPI_WARPS) {
                    const int row = i >> 3, pc = i & 7, lc = pc ^ (row & 7);
                    const int plane = row >= NW, n = row - plane * NW;
                    if (b0 + n < B) {
                        const uint4 v = *reinterpret_cast<const uint4*>(tile + row * 128 + pc * 16);
                        __half* dst = (plane ? hs_lo : hs_hi) + ((size_t)(b0 + n) * L + t) * R + rank * RU_UNITS + lc * 8;
                        *reinterpret_cast<uint4*>(dst) = v;
                    }
                }
                // the tile of parity nxt is rewritten two steps later; by then every peer has consumed this
                // push (it had to, to produce the h this CTA waits for) and the loop above has finished
            }
        }
        // next row block reuses both h buffers: every CTA must have finished its last MMAs / reads
        __syncthreads();
        cluster_arrive();
        cluster_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, A_TMEM ? RU_TMEM_COLS_A : RU_TMEM_COLS);
    }
}

}  // namespace tip
