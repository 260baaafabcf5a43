// LayerNorm GEMM (out_proj / ff2 + bias [+ dropout] + residual + LayerNorm, N = 256 = the whole row) for NARROW launches: a work
// unit is a PAIR of adjacent 128-row tiles that share every W k-block.
//
// Why: phase timestamps of the one-tile-per-CTA kernel (tools/umma_phases.py) show that it, too, runs at the SM's L2 port rate
// (~80-90 GB/s: ff2 streams A 512 KB + W 1 MB per tile, then parks 128 KB of residual and writes 128 KB of output), and W -- the
// same bytes for every row tile -- is two thirds of the mainloop's traffic.  In throughput mode a CTA owns two row
// tiles anyway; here it loads each W k-block ONCE and issues the MMAs of both tiles against it (two 256-column fp32
// accumulators = all 512 TMEM columns): 2.5 MB instead of 3.5 MB through the port per pair for ff2, 1.0 instead of 1.25 MB
// for out_proj.  (CTA pairs with cta_group::2 do not get this: with three MMAs per k-step re-reading B, the half of B that
// lives in the peer's shared memory crosses the SM-to-SM fabric three times.)
//   * 32-wide k-blocks (64-byte rows, SWIZZLE_64B): a stage = A0 hi/lo + A1 hi/lo (4 x 8 KB) + W hi/lo (2 x 16 KB) = 64 KB,
//     three stages = the same 192 KB ring as the one-tile kernel;
//   * after the last k-block both accumulators are complete; the epilogue of tile 0, then of tile 1, is the one-tile kernel's
//     LayerNorm epilogue unchanged (thread = accumulator row, residual tile parked in the idle ring by TMA, output boxes
//     staged in the ring's tail and TMA-stored); the producer parks tile 1's residual as soon as tile 0's epilogue hands the
//     ring back.
// Same products in the same order per accumulator element as the one-tile kernel: bit-identical outputs.
#pragma once
#include "tip_umma.cuh"

namespace tip {

constexpr int L2_BK = 32;
constexpr int L2_STAGES = 3;
constexpr int L2_A_BYTES = UM_BM * L2_BK * 2;                        // 8 KB: one plane of one row tile's k-block
constexpr int L2_B_BYTES = 256 * L2_BK * 2;                          // 16 KB: one plane of the W k-block (256 rows)
constexpr int L2_STAGE_BYTES = 4 * L2_A_BYTES + 2 * L2_B_BYTES;      // 64 KB
constexpr int L2_RBOX = UM_BM * 128;                                 // 16 KB: residual box, 128 rows x 64 columns (128-byte rows)
constexpr int L2_SMEM_BYTES = L2_STAGES * L2_STAGE_BYTES + 8 * 4096 /*bias / gamma / beta*/ + 1024 /*align slack*/ + 320 + 1024 /*row stats*/;
static_assert(L2_STAGES * L2_STAGE_BYTES >= 8 * L2_RBOX + UM_EPI_WARPS * 8192, "ring holds the residual tile and the output staging");

template <bool DROP>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_ln2_gemm_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,     // 128 x 32 boxes
                     const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,     // 256 x 32 boxes
                     const __grid_constant__ CUtensorMap mapC0, const __grid_constant__ CUtensorMap mapC1,         // 32 x 32 store boxes
                     const __grid_constant__ CUtensorMap mapR_hi, const __grid_constant__ CUtensorMap mapR_lo,     // 128 x 64 boxes
                     int M, int N, int K, int m_tile0, int m_tile_cnt, Epi ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    float* staging = reinterpret_cast<float*>(smem + L2_STAGES * L2_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L2_STAGES * L2_STAGE_BYTES + UM_EPI_WARPS * 4096);
    uint64_t* full_bar = bars;                        // [3] TMA -> MMA
    uint64_t* empty_bar = bars + L2_STAGES;           // [3] MMA (k-blocks) / epilogue (ring hand-back) -> TMA
    uint64_t* tfull_bar = bars + 2 * L2_STAGES;       //     both accumulators of the unit complete
    uint64_t* tempty_bar = bars + 2 * L2_STAGES + 1;  // [2] accumulator t is in the epilogue's registers
    uint64_t* rfull_bar = bars + 2 * L2_STAGES + 3;   //     residual tile landed in the ring
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * L2_STAGES + 4);
    volatile uint64_t* seed_slot = reinterpret_cast<volatile uint64_t*>(bars + 2 * L2_STAGES + 5);
    float* row_stat = reinterpret_cast<float*>(bars + 2 * L2_STAGES + 8);      // [2 halves][128 rows] partials, twice

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = (m_tile_cnt + 1) / 2;           // pairs of row tiles; the last one may hold a single tile
    const int num_kb = K / L2_BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mapA_hi); ptx::prefetch_tmap(&mapA_lo);
        ptx::prefetch_tmap(&mapB_hi); ptx::prefetch_tmap(&mapB_lo);
        ptx::prefetch_tmap(&mapR_hi); ptx::prefetch_tmap(&mapR_lo);
        for (int s = 0; s < L2_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(&tempty_bar[0], UM_EPI_WARPS); ptx::mbar_init(&tempty_bar[1], UM_EPI_WARPS);
        ptx::mbar_init(rfull_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();
    if (ep.pdl_early) griddep_launch();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t uses[L2_STAGES];                  // fills of each stage so far (k-blocks and residual tiles)
#pragma unroll
            for (int s = 0; s < L2_STAGES; ++s) uses[s] = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int m0a = (m_tile0 + 2 * unit) * UM_BM, m0b = m0a + UM_BM;
                const int ntile = (2 * unit + 1 < m_tile_cnt) ? 2 : 1;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], (uses[stage] & 1u) ^ 1u);
                    uses[stage]++;
                    uint8_t* s = smem + stage * L2_STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], L2_STAGE_BYTES);
                    ptx::tma_load_2d(s, &mapA_hi, &full_bar[stage], kb * L2_BK, m0a);
                    ptx::tma_load_2d(s + L2_A_BYTES, &mapA_lo, &full_bar[stage], kb * L2_BK, m0a);
                    // (a lone last tile: the partner's rows belong to nobody or to another batch part -- loaded, multiplied, never stored)
                    ptx::tma_load_2d(s + 2 * L2_A_BYTES, &mapA_hi, &full_bar[stage], kb * L2_BK, m0b);
                    ptx::tma_load_2d(s + 3 * L2_A_BYTES, &mapA_lo, &full_bar[stage], kb * L2_BK, m0b);
                    ptx::tma_load_2d(s + 4 * L2_A_BYTES, &mapB_hi, &full_bar[stage], kb * L2_BK, 0);
                    ptx::tma_load_2d(s + 4 * L2_A_BYTES + L2_B_BYTES, &mapB_lo, &full_bar[stage], kb * L2_BK, 0);
                    if (++stage == L2_STAGES) stage = 0;
                }
                for (int t = 0; t < ntile; ++t) {
                    // residual tile of row tile t -> ring bytes [0, 128 KB): box (plane p, column block cb) at (p * 4 + cb) * 16 KB.
                    // Needs the whole ring: every stage read by the MMAs (t = 0) / handed back by tile 0's epilogue (t = 1).
#pragma unroll
                    for (int s2 = 0; s2 < L2_STAGES; ++s2) { ptx::mbar_wait(&empty_bar[s2], (uses[s2] & 1u) ^ 1u); uses[s2]++; }
                    ptx::mbar_expect_tx(rfull_bar, 8 * L2_RBOX);
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        ptx::tma_load_2d(smem + cb * L2_RBOX, &mapR_hi, rfull_bar, cb * UM_BK, t ? m0b : m0a);
                        ptx::tma_load_2d(smem + (4 + cb) * L2_RBOX, &mapR_lo, rfull_bar, cb * UM_BK, t ? m0b : m0a);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: both row tiles against every W k-block =================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(UM_BM, 256);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
                ptx::mbar_wait(&tempty_bar[0], (uint32_t)((it & 1) ^ 1));
                ptx::mbar_wait(&tempty_bar[1], (uint32_t)((it & 1) ^ 1));
                ptx::tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * L2_STAGE_BYTES);
                    const uint64_t b_hi = umma_smem_desc_bk<32>(sa + 4 * L2_A_BYTES), b_lo = umma_smem_desc_bk<32>(sa + 4 * L2_A_BYTES + L2_B_BYTES);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint64_t a_hi = umma_smem_desc_bk<32>(sa + 2 * t * L2_A_BYTES), a_lo = umma_smem_desc_bk<32>(sa + (2 * t + 1) * L2_A_BYTES);
                        const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
#pragma unroll
                        for (int k = 0; k < L2_BK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 16 fp16 = 32 bytes along K
                            ptx::umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            ptx::umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            ptx::umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                        }
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == L2_STAGES) { stage = 0; phase ^= 1; }
                }
                // (the ring is then lent to the epilogue for the residual tiles / output staging: three hand-backs per tile, i.e.
                //  one full phase flip of every stage barrier per tile on the producer's side; the MMA side tracks full_bar only)
                ptx::umma_commit(tfull_bar);
            }
        }
    } else {
        // ================= epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const float asc = ep.acc_scale ? __ldg(ep.acc_scale) : 1.f;
        float* cvec = staging;                            // [0,256) bias, [256,512) gamma*16, [512,768) beta*16
        {
            const int t = (int)threadIdx.x - 64;          // 0..255 among the epilogue threads
            cvec[t] = __ldg(ep.bias + t) * (DROP ? ep.drop_inv : 1.f);       // kept elements carry 1/(1-p)
            cvec[256 + t] = __ldg(ep.gamma + t) * ACT_SCALE;
            cvec[512 + t] = __ldg(ep.beta + t) * ACT_SCALE;
            if (t == 0) *seed_slot = ep.drop_thr ? site_seed(ep.seed_ptr, ep.seed) : 0ull;
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        uint32_t rcount = 0;
        int it = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
            const int ntile = (2 * unit + 1 < m_tile_cnt) ? 2 : 1;
            ptx::mbar_wait(tfull_bar, (uint32_t)(it & 1));
            ptx::tc_fence_after();
            for (int t = 0; t < 2; ++t) {
                if (t >= ntile) {                          // nothing to store for the missing partner: just release its accumulator
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[t]);
                    continue;
                }
                const int m0 = (m_tile0 + 2 * unit + t) * UM_BM;
                const int rbase = m0 + quarter * 32;
                const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 256 + half * 128);
                {
                    const int trow = quarter * 32 + lane;             // row within the tile
                    const int col0 = half * (128);
                    float x[128];
#pragma unroll
                    for (int c = 0; c < 4; ++c) ptx::tmem_ld32_nowait(t_acc + c * 32, *reinterpret_cast<float(*)[32]>(&x[c * 32]));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // the accumulator is in registers: hand the TMEM buffer back to the MMA warp now
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[t]);
                    ptx::mbar_wait(rfull_bar, rcount & 1u); rcount++;    // residual tile landed (issued right after the last k-block)
                    float rsum = 0.f;
                    uint64_t sd_row = 0ull;                           // DROP: site seed + this row's first hash group (one shared-memory read per tile)
                    if constexpr (DROP) sd_row = *seed_slot;
                    const uint8_t* rrow = smem + trow * 128;          // this row inside every 16 KB box
                    const int rsw = trow & 7;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cb = half * 2 + (c >> 1);           // 64-column box of this chunk
                        const uint8_t* bh = rrow + cb * L2_RBOX;
                        const uint8_t* bl = rrow + (4 + cb) * L2_RBOX;
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
                    row_stat[half * 128 + trow] = rsum;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                    const float mean = (row_stat[trow] + row_stat[128 + trow]) * (1.f / 256);
                    float q2s = 0.f;
#pragma unroll
                    for (int j = 0; j < 128; ++j) { const float d = x[j] - mean; q2s = fmaf(d, d, q2s); }
                    row_stat[256 + half * 128 + trow] = q2s;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                    const float var = (row_stat[256 + trow] + row_stat[256 + 128 + trow]) * (1.f / 256);
                    const float ca = rsqrtf(var + 1e-5f), cb2 = -mean * ca;
                    uint8_t* obuf = smem + 8 * L2_RBOX + (warp - 2) * 8192;      // two 4 KB buffers (hi 2 KB | lo 2 KB)
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
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
                    // give the ring back to the producer once every warp's boxes have been read out of it
                    if (lane == 0) ptx::bulk_wait_read0();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 64) {
#pragma unroll
                        for (int s2 = 0; s2 < L2_STAGES; ++s2) ptx::mbar_arrive(&empty_bar[s2]);
                    }
                }
            }
        }
    }
    if (!ep.pdl_early) griddep_launch();
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// which: UG_OUT or UG_FF2; grid_cap: CTAs of the launch (<= 0: one per unit)
inline void launch_ln2_gemm(const UmmaMaps& mp, int which, int layer, int M, int N, int K, int m_tile0, int m_tiles, int grid_cap,
                            const Epi& ep, cudaStream_t st) {
    static bool attrs = false;
    if (!attrs) {
        cudaFuncSetAttribute(umma_ln2_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM_BYTES);
        cudaFuncSetAttribute(umma_ln2_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM_BYTES);
        attrs = true;
    }
    pdl_kind() = 1;
    const UmmaOperand& A = (which == UG_OUT) ? mp.a_att32 : mp.a_hid32;
    const UmmaOperand& B = (which == UG_OUT) ? mp.w_o32[layer] : mp.w_2k32[layer];
    const UmmaOperand& R = (which == UG_OUT) ? mp.a_xa : mp.a_xb;                  // residual through the 64-column operand boxes
    const UmmaOutput& C = (which == UG_OUT) ? mp.o_xb : mp.o_xa;
    const int units = (m_tiles + 1) / 2;
    const dim3 grid(std::min(units, grid_cap > 0 ? grid_cap : units));
    if (ep.drop_thr)
        launch_k(umma_ln2_gemm_kernel<true>, grid, dim3(UM_THREADS), L2_SMEM_BYTES, st, A.hi, A.lo, B.hi, B.lo, C.c0, C.c1, R.hi, R.lo, M, N, K, m_tile0, m_tiles, ep);
    else
        launch_k(umma_ln2_gemm_kernel<false>, grid, dim3(UM_THREADS), L2_SMEM_BYTES, st, A.hi, A.lo, B.hi, B.lo, C.c0, C.c1, R.hi, R.lo, M, N, K, m_tile0, m_tiles, ep);
}

}  // namespace tip
