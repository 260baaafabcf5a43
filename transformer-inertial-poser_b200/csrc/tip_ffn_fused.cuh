// Fused feed-forward block of an encoder layer (reference :91 -> nn.TransformerEncoderLayer):
//   x_out = LayerNorm2(x + dropout2(W2 * dropout(relu(W1 x + b1)) + b2))
// in ONE kernel per layer: the 1024-wide hidden activation never leaves the SM.
//
// Round 1 ran this as two kernels: ff1 (148 SMs, 128 x 128 tiles, `hid` written as FP16 hi/lo planes: 42 MB per layer)
// and ff2 + LayerNorm (80 SMs, whole-row tiles, `hid` read back).  Here one CTA owns a 128-row tile end to end:
//
//   for each of the 8 hidden chunks c (128 hidden units):
//     ff1(c):  hidacc[c & 1] (TMEM, 128 cols)  = x_tile (128 x 256) * W1[c]^T          8 k-blocks of 32: A and W1 streamed by TMA
//     epi(c):  relu(hidacc * s + b1) [dropout] -> FP16 hi/lo split -> `hidA` in shared memory, already in the K-major
//              SWIZZLE_64B layout a tcgen05 A operand wants (4 k-blocks of 32 hidden units)
//     ff2(c):  outacc (TMEM, 256 cols)        += hidA (128 x 128) * W2[:, c]^T          4 k-blocks of 32: W2 streamed by TMA
//   then the LayerNorm epilogue of the GEMM engine (residual parked in the idle operand ring by TMA, row in registers,
//   TMA-stored output boxes).
//
// One MMA-issuing thread runs the static schedule ff1(0), ff1(1), ff2(0), ff1(2), ff2(1), ... so the tensor pipe works on
// chunk c+1's first GEMM while the 8 epilogue warps convert chunk c; the producer warp streams the operand stages in
// exactly that order through a 4-stage ring of 32 KB stages.  3-product FP16 split everywhere, as in the GEMM engine.
#pragma once
#include "tip_umma.cuh"

namespace tip {

constexpr int FF_BK = 32;                          // fp16 elements per k-block (64-byte rows, SWIZZLE_64B)
constexpr int FF_STAGES = 4;
constexpr int FF_STAGE_BYTES = 32768;              // ff1: A hi|lo (2 x 8 KB) + W1 chunk hi|lo (2 x 8 KB); ff2: W2 hi|lo (2 x 16 KB)
constexpr int FF_HC = 128;                         // hidden units per chunk
constexpr int FF_NCHUNK = F / FF_HC;               // 8
constexpr int FF_KB1 = E / FF_BK;                  // 8 k-blocks per ff1 chunk
constexpr int FF_KB2 = FF_HC / FF_BK;              // 4 k-blocks per ff2 chunk
constexpr int FF_PLANE = UM_BM * FF_BK * 2;        // 8 KB: one plane of a 128-row k-block
constexpr int FF_HIDA_BYTES = FF_KB2 * 2 * FF_PLANE;   // 64 KB: [k-block][hi | lo]
constexpr int FF_RING_BYTES = FF_STAGES * FF_STAGE_BYTES;   // 128 KB (also: the residual tile of the LayerNorm epilogue)
constexpr int FF_CONST_FLOATS = 3 * E + F;         // b2, gamma*16, beta*16, b1
constexpr int FF_SMEM_BYTES = FF_RING_BYTES + FF_HIDA_BYTES + FF_CONST_FLOATS * 4 + 2048 /*row stats*/ + 256 /*barriers*/;
constexpr int FF_TMEM_COLS = 512;                  // [0,256) out accumulator, [256,384) / [384,512) hidden-chunk accumulators
static_assert(FF_SMEM_BYTES <= 232448, "fused FFN kernel exceeds the 227 KB of shared memory per CTA");

struct FfnArgs {
    const float* b1;          // [1024]
    const float* b2;          // [256]
    const float* gamma;       // [256]
    const float* beta;        // [256]
    const float* sc1;         // device scalars: 1 / (s_w1 * 16), 1 / (s_w2 * 16)
    const float* sc2;
    int M;                    // rows of the batch (b * L + t)
    int m_tile0, m_tiles;     // row tiles [m_tile0, m_tile0 + m_tiles)
    uint32_t drop_thr;        // encoder dropout (0 = off): threshold / scale of BOTH sites
    float drop_inv;
    const uint64_t* seed_ptr;
    uint64_t seed_ff1, seed_ff2;
    int pdl_early;
};

template <bool DROP>
__global__ void __launch_bounds__(UM_THREADS, 1)
ffn_ln_kernel(const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,      // x (xb), 32-col boxes x 128 rows
              const __grid_constant__ CUtensorMap mapW1_hi, const __grid_constant__ CUtensorMap mapW1_lo,    // W1 [1024][256], 32-col boxes x 128 rows
              const __grid_constant__ CUtensorMap mapW2_hi, const __grid_constant__ CUtensorMap mapW2_lo,    // W2 [256][1024], 32-col boxes x 256 rows
              const __grid_constant__ CUtensorMap mapR_hi, const __grid_constant__ CUtensorMap mapR_lo,      // x again as the residual: 64-col boxes (SWIZZLE_128B)
              const __grid_constant__ CUtensorMap mapC0, const __grid_constant__ CUtensorMap mapC1,          // output planes, 32 x 32 store boxes
              FfnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* hidA = smem + FF_RING_BYTES;
    float* cvec = reinterpret_cast<float*>(smem + FF_RING_BYTES + FF_HIDA_BYTES);   // [0,256) b2, [256,512) gamma*16, [512,768) beta*16, [768,1792) b1
    float* row_stat = cvec + FF_CONST_FLOATS;                                       // [512] LayerNorm partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(row_stat + 512);
    uint64_t* full_bar = bars;               // [4] TMA -> MMA
    uint64_t* empty_bar = bars + 4;          // [4] MMA -> TMA
    uint64_t* hfull_bar = bars + 8;          // [2] ff1(c) accumulated -> epilogue
    uint64_t* hempty_bar = bars + 10;        // [2] epilogue has read the hidden accumulator -> MMA (count 8)
    uint64_t* hready_bar = bars + 12;        //     hidA(c) written -> MMA (count 8)
    uint64_t* hfree_bar = bars + 13;         //     ff2(c) done with hidA -> epilogue
    uint64_t* ofull_bar = bars + 14;         //     all ff2 of the tile accumulated -> LayerNorm epilogue
    uint64_t* oempty_bar = bars + 15;        //     LayerNorm epilogue has read the out accumulator -> MMA (count 8)
    uint64_t* rfull_bar = bars + 16;         //     residual tile landed in the ring
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    volatile uint64_t* seed_slot = reinterpret_cast<volatile uint64_t*>(bars + 18);     // [2]: ff1 site, ff2 site

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mapX_hi); ptx::prefetch_tmap(&mapX_lo);
        ptx::prefetch_tmap(&mapW1_hi); ptx::prefetch_tmap(&mapW1_lo);
        ptx::prefetch_tmap(&mapW2_hi); ptx::prefetch_tmap(&mapW2_lo);
        for (int s = 0; s < FF_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&hfull_bar[s], 1); ptx::mbar_init(&hempty_bar[s], UM_EPI_WARPS); }
        ptx::mbar_init(hready_bar, UM_EPI_WARPS);
        ptx::mbar_init(hfree_bar, 1);
        ptx::mbar_init(ofull_bar, 1);
        ptx::mbar_init(oempty_bar, UM_EPI_WARPS);
        ptx::mbar_init(rfull_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, FF_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();
    if (a.pdl_early) griddep_launch();

    if (warp == 0) {
        // ================= TMA producer: stages in the MMA warp's consumption order =================
        if (lane == 0) {
            int stage = 0;
            uint32_t uses[FF_STAGES] = {0u, 0u, 0u, 0u};
            ptx::prefetch_tmap(&mapR_hi); ptx::prefetch_tmap(&mapR_lo);
            for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x) {
                const int m0 = (a.m_tile0 + tile) * UM_BM;
                for (int s = 0; s <= FF_NCHUNK; ++s) {
                    if (s < FF_NCHUNK) {                       // ff1(s): x k-block + W1 rows [128 s, 128 s + 128)
                        for (int kb = 0; kb < FF_KB1; ++kb) {
                            ptx::mbar_wait(&empty_bar[stage], (uses[stage] & 1u) ^ 1u);
                            uses[stage]++;
                            uint8_t* st = smem + stage * FF_STAGE_BYTES;
                            ptx::mbar_expect_tx(&full_bar[stage], 4 * FF_PLANE);
                            ptx::tma_load_2d(st, &mapX_hi, &full_bar[stage], kb * FF_BK, m0);
                            ptx::tma_load_2d(st + FF_PLANE, &mapX_lo, &full_bar[stage], kb * FF_BK, m0);
                            ptx::tma_load_2d(st + 2 * FF_PLANE, &mapW1_hi, &full_bar[stage], kb * FF_BK, s * FF_HC);
                            ptx::tma_load_2d(st + 3 * FF_PLANE, &mapW1_lo, &full_bar[stage], kb * FF_BK, s * FF_HC);
                            if (++stage == FF_STAGES) stage = 0;
                        }
                    }
                    if (s >= 1) {                              // ff2(s-1): W2 columns [128 (s-1), +128) of all 256 rows
                        for (int kb = 0; kb < FF_KB2; ++kb) {
                            ptx::mbar_wait(&empty_bar[stage], (uses[stage] & 1u) ^ 1u);
                            uses[stage]++;
                            uint8_t* st = smem + stage * FF_STAGE_BYTES;
                            ptx::mbar_expect_tx(&full_bar[stage], 4 * FF_PLANE);
                            ptx::tma_load_2d(st, &mapW2_hi, &full_bar[stage], (s - 1) * FF_HC + kb * FF_BK, 0);
                            ptx::tma_load_2d(st + 2 * FF_PLANE, &mapW2_lo, &full_bar[stage], (s - 1) * FF_HC + kb * FF_BK, 0);
                            if (++stage == FF_STAGES) stage = 0;
                        }
                    }
                }
                // residual tile [128 rows x 256 cols] hi + lo -> ring bytes [0, 128 KB) as eight 64-column boxes; needs the whole
                // ring: every stage must have been consumed
#pragma unroll
                for (int s2 = 0; s2 < FF_STAGES; ++s2) { ptx::mbar_wait(&empty_bar[s2], (uses[s2] & 1u) ^ 1u); uses[s2]++; }
                ptx::mbar_expect_tx(rfull_bar, 8 * 16384);
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    ptx::tma_load_2d(smem + cb * 16384, &mapR_hi, rfull_bar, cb * UM_BK, m0);
                    ptx::tma_load_2d(smem + (4 + cb) * 16384, &mapR_lo, rfull_bar, cb * UM_BK, m0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc_f16(UM_BM, FF_HC);      // 128 x 128 (hidden chunk)
            constexpr uint32_t idesc2 = umma_idesc_f16(UM_BM, E);          // 128 x 256 (output)
            int stage = 0; uint32_t phase = 0;
            uint32_t n_h[2] = {0u, 0u};          // uses of the two hidden accumulators
            uint32_t n_ready = 0, n_tiles = 0;
            const uint32_t hid_s = ptx::smem_u32(hidA);
            for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++n_tiles) {
                for (int s = 0; s <= FF_NCHUNK; ++s) {
                    if (s < FF_NCHUNK) {
                        const int hb = s & 1;
                        ptx::mbar_wait(&hempty_bar[hb], (n_h[hb] & 1u) ^ 1u);      // the epilogue has read this buffer's previous chunk
                        n_h[hb]++;
                        ptx::tc_fence_after();
                        const uint32_t d_h = tmem_base + 256u + (uint32_t)(hb * FF_HC);
                        for (int kb = 0; kb < FF_KB1; ++kb) {
                            ptx::mbar_wait(&full_bar[stage], phase);
                            ptx::tc_fence_after();
                            const uint32_t sa = ptx::smem_u32(smem + stage * FF_STAGE_BYTES);
                            const uint64_t a_hi = umma_smem_desc_bk<FF_BK>(sa), a_lo = umma_smem_desc_bk<FF_BK>(sa + FF_PLANE);
                            const uint64_t b_hi = umma_smem_desc_bk<FF_BK>(sa + 2 * FF_PLANE), b_lo = umma_smem_desc_bk<FF_BK>(sa + 3 * FF_PLANE);
#pragma unroll
                            for (int k = 0; k < FF_BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                ptx::umma_f16(d_h, a_lo + adv, b_hi + adv, idesc1, (kb | k) ? 1u : 0u);
                                ptx::umma_f16(d_h, a_hi + adv, b_lo + adv, idesc1, 1u);
                                ptx::umma_f16(d_h, a_hi + adv, b_hi + adv, idesc1, 1u);
                            }
                            ptx::umma_commit(&empty_bar[stage]);
                            if (++stage == FF_STAGES) { stage = 0; phase ^= 1; }
                        }
                        ptx::umma_commit(&hfull_bar[hb]);
                    }
                    if (s >= 1) {
                        const int c = s - 1;
                        if (c == 0) {                               // the previous tile's LayerNorm epilogue has read the out accumulator
                            ptx::mbar_wait(oempty_bar, (n_tiles & 1u) ^ 1u);
                        }
                        ptx::mbar_wait(hready_bar, n_ready & 1u);   // the epilogue has written hidA(c)
                        n_ready++;
                        ptx::tc_fence_after();
                        for (int kb = 0; kb < FF_KB2; ++kb) {
                            ptx::mbar_wait(&full_bar[stage], phase);
                            ptx::tc_fence_after();
                            const uint32_t sa = ptx::smem_u32(smem + stage * FF_STAGE_BYTES);
                            const uint64_t a_hi = umma_smem_desc_bk<FF_BK>(hid_s + kb * 2 * FF_PLANE);
                            const uint64_t a_lo = umma_smem_desc_bk<FF_BK>(hid_s + kb * 2 * FF_PLANE + FF_PLANE);
                            const uint64_t b_hi = umma_smem_desc_bk<FF_BK>(sa), b_lo = umma_smem_desc_bk<FF_BK>(sa + 2 * FF_PLANE);
#pragma unroll
                            for (int k = 0; k < FF_BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                ptx::umma_f16(tmem_base, a_lo + adv, b_hi + adv, idesc2, (c | kb | k) ? 1u : 0u);
                                ptx::umma_f16(tmem_base, a_hi + adv, b_lo + adv, idesc2, 1u);
                                ptx::umma_f16(tmem_base, a_hi + adv, b_hi + adv, idesc2, 1u);
                            }
                            ptx::umma_commit(&empty_bar[stage]);
                            if (++stage == FF_STAGES) { stage = 0; phase ^= 1; }
                        }
                        ptx::umma_commit(hfree_bar);                // hidA may be overwritten once these MMAs retire
                    }
                }
                ptx::umma_commit(ofull_bar);
            }
        }
    } else {
        // ================= epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int trow = quarter * 32 + lane;                 // row within the tile
        const float asc1 = __ldg(a.sc1), asc2 = __ldg(a.sc2);
        {   // constants -> shared memory (the kernel leaves ~3 KB of L1: every __ldg in the loops would be an L2 round trip)
            const int t = (int)threadIdx.x - 64;              // 0..255
            cvec[t] = __ldg(a.b2 + t) * (DROP ? a.drop_inv : 1.f);
            cvec[256 + t] = __ldg(a.gamma + t) * ACT_SCALE;
            cvec[512 + t] = __ldg(a.beta + t) * ACT_SCALE;
#pragma unroll
            for (int i = 0; i < 4; ++i) cvec[768 + 4 * t + i] = __ldg(a.b1 + 4 * t + i) * ACT_SCALE * (DROP ? a.drop_inv : 1.f);
            if (t == 0) {
                seed_slot[0] = DROP ? site_seed(a.seed_ptr, a.seed_ff1) : 0ull;
                seed_slot[1] = DROP ? site_seed(a.seed_ptr, a.seed_ff2) : 0ull;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const float sc1 = asc1 * ACT_SCALE * (DROP ? a.drop_inv : 1.f);    // hidden planes hold 16 * h (kept elements: x 1/(1-p))
        const uint32_t thr_hi = a.drop_thr << 16;
        uint32_t n_h[2] = {0u, 0u};
        uint32_t n_free = 0, n_tiles = 0;
        for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++n_tiles) {
            const int m0 = (a.m_tile0 + tile) * UM_BM;
            const int rbase = m0 + quarter * 32;
            // ---------------- hidden chunks: TMEM -> relu / dropout / split -> hidA ----------------
#pragma unroll 1
            for (int c = 0; c < FF_NCHUNK; ++c) {
                const int hb = c & 1;
                ptx::mbar_wait(&hfull_bar[hb], n_h[hb] & 1u);
                n_h[hb]++;
                ptx::tc_fence_after();
                const uint32_t t_h = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)(hb * FF_HC + half * 64);
                float v[64];
                ptx::tmem_ld32_nowait(t_h, *reinterpret_cast<float(*)[32]>(&v[0]));
                ptx::tmem_ld32_nowait(t_h + 32, *reinterpret_cast<float(*)[32]>(&v[32]));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&hempty_bar[hb]);           // ff1(c + 2) may overwrite this accumulator
                const float* b1s = cvec + 768 + c * FF_HC + half * 64;
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4) {
                    const float4 b = *reinterpret_cast<const float4*>(b1s + 4 * j4);     // broadcast
                    v[4 * j4 + 0] = fmaxf(fmaf(v[4 * j4 + 0], sc1, b.x), 0.f);
                    v[4 * j4 + 1] = fmaxf(fmaf(v[4 * j4 + 1], sc1, b.y), 0.f);
                    v[4 * j4 + 2] = fmaxf(fmaf(v[4 * j4 + 2], sc1, b.z), 0.f);
                    v[4 * j4 + 3] = fmaxf(fmaf(v[4 * j4 + 3], sc1, b.w), 0.f);
                }
                if constexpr (DROP) {                                       // dropout(relu(linear1(x))): element index = row * 1024 + hidden unit
                    const uint64_t g0 = ((uint64_t)(m0 + trow) * F + c * FF_HC + half * 64) >> 2;
                    const uint64_t sd = seed_slot[0];
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4) {
                        const uint64_t h = hash_u64(sd, g0 + j4);
                        const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
                        if ((hl << 16) < thr_hi) v[4 * j4 + 0] = 0.f;
                        if (hl < thr_hi) v[4 * j4 + 1] = 0.f;
                        if ((hh << 16) < thr_hi) v[4 * j4 + 2] = 0.f;
                        if (hh < thr_hi) v[4 * j4 + 3] = 0.f;
                    }
                }
                // hidA may be overwritten once ff2(c - 1) has retired (first chunk of the first tile: nothing to wait for)
                if (c > 0 || n_tiles > 0) ptx::mbar_wait(hfree_bar, n_free & 1u);
                if (c > 0 || n_tiles > 0) n_free++;
                // this thread's 64 hidden units = k-blocks 2 half, 2 half + 1 of the chunk; row trow of each [128 x 64 B] plane,
                // 16-byte chunks XOR-ed with (row >> 1) & 3 (SWIZZLE_64B, what the tcgen05 descriptor of hidA expects)
                const int sw = (trow >> 1) & 3;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    uint8_t* ph = hidA + (2 * half + kb) * 2 * FF_PLANE + trow * 64;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                           // 8 hidden units per 16-byte chunk
                        uint32_t uh[4], ul[4];
#pragma unroll
                        for (int p2 = 0; p2 < 4; ++p2) {
                            float h0, h1, l0, l1;
                            veltkamp11(v[32 * kb + 8 * j + 2 * p2], h0, l0); veltkamp11(v[32 * kb + 8 * j + 2 * p2 + 1], h1, l1);
                            __half2 t2 = __floats2half2_rn(h0, h1); uh[p2] = *reinterpret_cast<uint32_t*>(&t2);
                            t2 = __floats2half2_rn(l0, l1); ul[p2] = *reinterpret_cast<uint32_t*>(&t2);
                        }
                        *reinterpret_cast<uint4*>(ph + ((j ^ sw) << 4)) = make_uint4(uh[0], uh[1], uh[2], uh[3]);
                        *reinterpret_cast<uint4*>(ph + FF_PLANE + ((j ^ sw) << 4)) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
                    }
                }
                ptx::fence_async_smem();                                    // generic writes -> async proxy (tcgen05 operand reads)
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(hready_bar);
            }
            // ---------------- LayerNorm epilogue (the GEMM engine's ring path; BN = 256) ----------------
            ptx::mbar_wait(ofull_bar, n_tiles & 1u);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 128);
            const int col0 = half * 128;
            float x[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) ptx::tmem_ld32_nowait(t_acc + c * 32, *reinterpret_cast<float(*)[32]>(&x[c * 32]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(oempty_bar);                    // the next tile's ff2 may start accumulating
            ptx::mbar_wait(rfull_bar, n_tiles & 1u);                        // residual tile landed
            float rsum = 0.f;
            const uint8_t* rrow = smem + trow * 128;                        // this row inside every 16 KB box
            const int rsw = trow & 7;
            const float ascd = asc2 * (DROP ? a.drop_inv : 1.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cb = half * 2 + (c >> 1);                         // 64-column box of this chunk
                const uint8_t* bh = rrow + cb * 16384;
                const uint8_t* bl = rrow + (4 + cb) * 16384;
#pragma unroll
                for (int i = 0; i < 4; ++i) {                               // 16-byte chunk = 8 columns
                    const int pos = (((c & 1) * 4 + i) ^ rsw) << 4;
                    const uint4 h4 = *reinterpret_cast<const uint4*>(bh + pos);
                    const uint4 l4 = *reinterpret_cast<const uint4*>(bl + pos);
                    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w};
                    const uint32_t lw[4] = {l4.x, l4.y, l4.z, l4.w};
                    const float4 b0 = *reinterpret_cast<const float4*>(cvec + col0 + c * 32 + i * 8);
                    const float4 b1v = *reinterpret_cast<const float4*>(cvec + col0 + c * 32 + i * 8 + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
                    uint32_t hw2[4] = {0u, 0u, 0u, 0u};
                    if constexpr (DROP) {                                   // dropout2: element index = row * 256 + column
                        const uint64_t g0 = ((uint64_t)(m0 + trow) * E + col0 + c * 32 + i * 8) >> 2;
                        const uint64_t sd = seed_slot[1];
                        const uint64_t ha = hash_u64(sd, g0), hb2 = hash_u64(sd, g0 + 1);
                        hw2[0] = (uint32_t)ha; hw2[1] = (uint32_t)(ha >> 32); hw2[2] = (uint32_t)hb2; hw2[3] = (uint32_t)(hb2 >> 32);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) {
                        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[q2]));
                        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[q2]));
                        const int j = c * 32 + i * 8 + q2 * 2;
                        float s0 = fmaf(x[j], ascd, bb[q2 * 2]), s1 = fmaf(x[j + 1], ascd, bb[q2 * 2 + 1]);
                        if constexpr (DROP) {
                            if ((hw2[q2] << 16) < thr_hi) s0 = 0.f;
                            if (hw2[q2] < thr_hi) s1 = 0.f;
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
            const float mean = (row_stat[trow] + row_stat[128 + trow]) * (1.f / E);
            float q2s = 0.f;
#pragma unroll
            for (int j = 0; j < 128; ++j) { const float d = x[j] - mean; q2s = fmaf(d, d, q2s); }
            row_stat[256 + half * 128 + trow] = q2s;
            asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
            const float var = (row_stat[256 + trow] + row_stat[256 + 128 + trow]) * (1.f / E);
            const float ca = rsqrtf(var + 1e-5f), cb2 = -mean * ca;
            uint8_t* obuf = hidA + (warp - 2) * 8192;                       // two 4 KB buffers (hi 2 KB | lo 2 KB); hidA is idle now
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int colb = col0 + c * 32;
                uint8_t* sbuf = obuf + (c & 1) * 4096;
                if (c >= 2) {
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
            // hand the ring (residual) and hidA (output boxes) back once every warp's boxes have been read out
            if (lane == 0) ptx::bulk_wait_read0();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
#pragma unroll
                for (int s2 = 0; s2 < FF_STAGES; ++s2) ptx::mbar_arrive(&empty_bar[s2]);
            }
        }
    }
    if (!a.pdl_early) griddep_launch();
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, FF_TMEM_COLS);
    }
}

}  // namespace tip
