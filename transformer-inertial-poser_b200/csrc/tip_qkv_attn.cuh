// Fused QKV projection + causal self-attention (reference :85-91 -> nn.MultiheadAttention inside
// nn.TransformerEncoderLayer): qkv = x W_in^T + b, P = softmax(q k^T + causal mask) [dropout], o = P v.
//
// Round 1 ran this as two kernels: the QKV GEMM wrote q|k|v as FP16 hi/lo planes (31.5 MB per layer at B = 256) and
// the attention kernel read them back -- 41 us per layer together, both occupying every SM.  Here the projection's
// accumulator never leaves the SM:
//
//   * work unit = (the windows of one row tile) x (4 heads): a 128 x 192 tcgen05 tile whose columns are q | k | v of
//     those 4 heads (W_in's rows are re-ordered at pack time so they are contiguous).  Row tiles are cut on WINDOW
//     boundaries: wpt = 128 / L windows (3 at L = 40 -> 120 live rows), so a tile holds every key of its queries.
//   * mainloop: the GEMM engine's warp-specialised pipeline (TMA producer warp, one MMA-issuing thread, 3-product
//     FP16 split, fp32 accumulation in a double-buffered TMEM accumulator) with 32-wide k-blocks (SWIZZLE_64B) in a
//     4-stage ring, so 64 KB of shared memory stay free for ...
//   * ... the epilogue = the attention itself, in plain fp32 FFMA (no split needed: nothing is re-quantised).  Thread =
//     accumulator row = one query.  Each of the 8 epilogue warps reads q, k, v of its 2 heads from tensor memory, adds
//     the biases, parks k and v (fp32) in shared memory, and after a block barrier walks the keys of its window:
//     scores into registers (<= 40), max, exp2, sum, P v -- key / value rows are warp-wide broadcasts.  The output
//     goes straight to the FP16 hi/lo planes the out-projection GEMM reads.
//
// The mma.sync attention kernel (tip_attn_mma.cuh) and the plain QKV GEMM remain for L-independent cross-checks
// (TIP_FUSED_ATTN=0) and the FFMA engine.
#pragma once
#include "tip_umma.cuh"

namespace tip {

constexpr int QA_HG = 4;                         // heads per work unit
constexpr int QA_GROUPS = NH / QA_HG;            // 4 head groups
constexpr int QA_BN = 3 * QA_HG * HD;            // 192 accumulator columns: q(64) | k(64) | v(64)
constexpr int QA_BK = 32;                        // fp16 elements per k-block (64-byte rows, SWIZZLE_64B)
constexpr int QA_STAGES = 4;
constexpr int QA_A_BYTES = UM_BM * QA_BK * 2;    // 8 KB per plane per stage
constexpr int QA_B_BYTES = QA_BN * QA_BK * 2;    // 12 KB
constexpr int QA_STAGE_BYTES = 2 * (QA_A_BYTES + QA_B_BYTES);       // 40 KB
constexpr int QA_KV_FLOATS = UM_BM * QA_HG * 2 * HD;                // [128 rows][4 heads][k16 | v16] fp32 = 64 KB
constexpr int QA_SMEM_BYTES = QA_STAGES * QA_STAGE_BYTES + QA_KV_FLOATS * 4 + QA_BN * 4 /*bias*/ + 256 /*barriers*/;
constexpr int QA_TMEM_COLS = 512;                // 2 x 192 accumulator columns -> next power of two
static_assert(QA_SMEM_BYTES <= 232448, "fused QKV + attention kernel exceeds the 227 KB of shared memory per CTA");

// kv[row][8 chunks: k0..k3, v0..v3][4 heads] in 16-byte units, index XOR-ed with a per-row swizzle: the writes (thread = row)
// are bank-conflict free, the reads are broadcasts with at most four distinct addresses per warp (two heads x two
// windows) that fall into different bank groups (adjacent heads = adjacent 16-byte units; rows 40 apart get different swizzles)
__device__ __forceinline__ float4* qa_kv_ptr(float* kv, int row, int head, int chunk) {
    return reinterpret_cast<float4*>(kv) + row * (QA_HG * 8) + ((chunk * QA_HG + head) ^ ((row ^ (row >> 3)) & 7));
}

__global__ void __launch_bounds__(UM_THREADS, 1)
qkv_attn_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                const float* __restrict__ bias_r,           // [768] re-ordered like the weight rows
                __half* __restrict__ out_hi, __half* __restrict__ out_lo,     // att planes [rows][256] of 16 * o
                const float* __restrict__ acc_scale,        // 1 / (s_w * 16)
                int w0, int nw, int L,                      // windows [w0, w0 + nw) of the batch, L rows each
                float drop_p, uint32_t drop_thr, const uint64_t* __restrict__ seed_ptr, uint64_t seed_off, int pdl_early) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    float* kv = reinterpret_cast<float*>(smem + QA_STAGES * QA_STAGE_BYTES);
    float* sbias = kv + QA_KV_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + QA_BN);
    uint64_t* full_bar = bars;                        // [4] TMA -> MMA
    uint64_t* empty_bar = bars + QA_STAGES;           // [4] MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * QA_STAGES;       // [2] MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * QA_STAGES + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * QA_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpt = UM_BM / L;                        // windows per row tile
    const int rows_used = wpt * L;                    // live rows of a tile (120 at L = 40)
    const int m_units = (nw + wpt - 1) / wpt;
    const int total_units = m_units * QA_GROUPS;
    constexpr int num_kb = E / QA_BK;                 // 8 k-blocks

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mapA_hi); ptx::prefetch_tmap(&mapA_lo);
        ptx::prefetch_tmap(&mapW_hi); ptx::prefetch_tmap(&mapW_lo);
        for (int s = 0; s < QA_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], UM_EPI_WARPS); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, QA_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();
    if (pdl_early) griddep_launch();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t uses[QA_STAGES] = {0u, 0u, 0u, 0u};
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
                const int m0 = (w0 + (unit / QA_GROUPS) * wpt) * L;          // first row of the tile (rows beyond the tensor: zero fill)
                const int n0 = (unit % QA_GROUPS) * QA_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], (uses[stage] & 1u) ^ 1u);
                    uses[stage]++;
                    uint8_t* s = smem + stage * QA_STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], QA_STAGE_BYTES);
                    ptx::tma_load_2d(s, &mapA_hi, &full_bar[stage], kb * QA_BK, m0);
                    ptx::tma_load_2d(s + QA_A_BYTES, &mapA_lo, &full_bar[stage], kb * QA_BK, m0);
                    ptx::tma_load_2d(s + 2 * QA_A_BYTES, &mapW_hi, &full_bar[stage], kb * QA_BK, n0);
                    ptx::tma_load_2d(s + 2 * QA_A_BYTES + QA_B_BYTES, &mapW_lo, &full_bar[stage], kb * QA_BK, n0);
                    if (++stage == QA_STAGES) stage = 0;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(UM_BM, QA_BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
                const int as = it & 1;
                ptx::mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * QA_BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * QA_STAGE_BYTES);
                    const uint64_t a_hi = umma_smem_desc_bk<QA_BK>(sa), a_lo = umma_smem_desc_bk<QA_BK>(sa + QA_A_BYTES);
                    const uint64_t b_hi = umma_smem_desc_bk<QA_BK>(sa + 2 * QA_A_BYTES);
                    const uint64_t b_lo = umma_smem_desc_bk<QA_BK>(sa + 2 * QA_A_BYTES + QA_B_BYTES);
#pragma unroll
                    for (int k = 0; k < QA_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        ptx::umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                        ptx::umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                        ptx::umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == QA_STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        // ================= epilogue = attention: warps 2..9; TMEM lane quarter = warp % 4, head pair = (warp - 2) / 4 ====
        // A thread reads the accumulator row of ITS query (tcgen05.ld: lane = row) for the warp's two heads, but then
        // works on ONE head for TWO adjacent queries (its own row and its lane-pair partner's; the partner's q comes by
        // shuffle): every key / value row fetched from shared memory feeds two dot products.  Shared-memory bandwidth
        // (one 16-byte broadcast load = 4 wavefronts) is what bounds this phase, not the FFMAs.
        const int quarter = warp & 3;
        const int ch = (warp - 2) >> 2;                   // this warp's heads within the group: 2 ch, 2 ch + 1
        const int trow = quarter * 32 + lane;             // row within the tile = query this thread loads
        const int hsel = lane & 1;                        // head this thread computes: 2 ch + hsel
        const int hl = 2 * ch + hsel;
        const float asc = __ldg(acc_scale);
        const float inv_keep = drop_inv_keep(drop_p);
        const uint64_t seed = drop_thr ? site_seed(seed_ptr, seed_off) : 0ull;
        // the query pair (even row, odd row): same window (L is even), positions pos_e and pos_e + 1
        const int trow_e = trow & ~1;
        const int wloc = trow_e / L;
        const int pos_e = trow_e - wloc * L;
        const int wrow0 = (trow_e < rows_used) ? wloc * L : 0;      // (rows past the tile's last window never index beyond the k / v tile)
        int it = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
            const int as = it & 1;
            const int mt = unit / QA_GROUPS, g = unit % QA_GROUPS;
            const int win = w0 + mt * wpt + wloc;                          // global window of the query pair
            const bool live = trow_e < rows_used && (mt * wpt + wloc) < nw;
            {   // this unit's bias slice -> shared memory
                const int et = (int)threadIdx.x - 64;
                if (et < QA_BN) sbias[et] = __ldg(bias_r + g * QA_BN + et);
            }
            ptx::mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            ptx::tc_fence_after();
            asm volatile("bar.sync 1, 256;" ::: "memory");                  // bias visible; every warp is done with the previous unit's k / v
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * QA_BN);
            float qe[HD], qo[HD];
            {
                float q[32], kk[32], vv[32];
                ptx::tmem_ld32_nowait(t_acc + 32 * ch, q);                  // q of heads 2ch, 2ch+1 (this thread's row)
                ptx::tmem_ld32_nowait(t_acc + 64 + 32 * ch, kk);            // k
                ptx::tmem_ld32_nowait(t_acc + 128 + 32 * ch, vv);           // v
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);           // the accumulator is in registers: the next unit's MMAs may start
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 bk = *reinterpret_cast<const float4*>(sbias + 64 + 32 * ch + 16 * hh + 4 * c4);
                        const float4 bv = *reinterpret_cast<const float4*>(sbias + 128 + 32 * ch + 16 * hh + 4 * c4);
                        const int j = 16 * hh + 4 * c4;
                        *qa_kv_ptr(kv, trow, 2 * ch + hh, c4) = make_float4(fmaf(kk[j], asc, bk.x), fmaf(kk[j + 1], asc, bk.y),
                                                                            fmaf(kk[j + 2], asc, bk.z), fmaf(kk[j + 3], asc, bk.w));
                        *qa_kv_ptr(kv, trow, 2 * ch + hh, 4 + c4) = make_float4(fmaf(vv[j], asc, bv.x), fmaf(vv[j + 1], asc, bv.y),
                                                                                fmaf(vv[j + 2], asc, bv.z), fmaf(vv[j + 3], asc, bv.w));
                    }
                }
                // q (+ bias; 1/sqrt(d) is folded into W_q, b_q) of head hl for both queries of the pair: the even lane keeps its
                // head-2ch slice and receives the odd row's, the odd lane keeps its head-(2ch+1) slice and receives the even row's
#pragma unroll
                for (int i = 0; i < HD; ++i) {
                    const float own0 = fmaf(q[i], asc, sbias[32 * ch + i]);             // this row, head 2ch
                    const float own1 = fmaf(q[16 + i], asc, sbias[32 * ch + 16 + i]);   // this row, head 2ch + 1
                    const float recv = __shfl_xor_sync(0xffffffffu, hsel ? own0 : own1, 1);
                    qe[i] = hsel ? recv : own0;
                    qo[i] = hsel ? own1 : recv;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");                  // k / v of all 4 heads of all rows are in shared memory
            const int mypos_e = live ? pos_e : -1, mypos_o = live ? pos_e + 1 : -1;
            int jmax = mypos_o;                                             // keys this warp has to walk
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
            // online softmax over blocks of 8 keys: 8 scores per query in registers, ONE rescale of (l, o) per block, then P v
            float me = -INFINITY, mo = -INFINITY, le = 0.f, lo = 0.f, oe[HD], oo[HD];
#pragma unroll
            for (int i = 0; i < HD; ++i) { oe[i] = 0.f; oo[i] = 0.f; }
#pragma unroll 1
            for (int j0 = 0; j0 <= jmax; j0 += 8) {                         // warp-uniform trip count
                float se[8], so[8];
                float bme = -INFINITY, bmo = -INFINITY;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u;
                    const int kr = wrow0 + min(j, L - 1);
                    float ae = 0.f, ao = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 k4 = *qa_kv_ptr(kv, kr, hl, c4);
                        ae = fmaf(qe[4 * c4], k4.x, ae); ao = fmaf(qo[4 * c4], k4.x, ao);
                        ae = fmaf(qe[4 * c4 + 1], k4.y, ae); ao = fmaf(qo[4 * c4 + 1], k4.y, ao);
                        ae = fmaf(qe[4 * c4 + 2], k4.z, ae); ao = fmaf(qo[4 * c4 + 2], k4.z, ao);
                        ae = fmaf(qe[4 * c4 + 3], k4.w, ae); ao = fmaf(qo[4 * c4 + 3], k4.w, ao);
                    }
                    se[u] = (j <= mypos_e) ? ae * 1.4426950408889634f : -INFINITY;       // log2 domain; keys after the query: masked
                    so[u] = (j <= mypos_o) ? ao * 1.4426950408889634f : -INFINITY;
                    bme = fmaxf(bme, se[u]);
                    bmo = fmaxf(bmo, so[u]);
                }
                const float mne = fmaxf(me, bme), mno = fmaxf(mo, bmo);
                const float mse = (mne == -INFINITY) ? 0.f : mne, mso = (mno == -INFINITY) ? 0.f : mno;   // (rows that are not live never see a key)
                const float ale = exp2f(me - mse), alo = exp2f(mo - mso);   // exp2(-inf) = 0 on the first block
                le *= ale; lo *= alo;
#pragma unroll
                for (int i = 0; i < HD; ++i) { oe[i] *= ale; oo[i] *= alo; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u;
                    float pe = exp2f(se[u] - mse), po = exp2f(so[u] - mso);             // 0 for masked keys
                    le += pe; lo += po;
                    if (drop_thr) {                                         // attention-probability dropout (train mode): warp-uniform
                        const uint64_t bh = (uint64_t)win * NH + g * QA_HG + hl;
                        const uint64_t ie = attn_drop_index(bh, max(mypos_e, 0), j), io = attn_drop_index(bh, max(mypos_o, 0), j);
                        const uint64_t he = hash_u64(seed, ie >> 2), ho = hash_u64(seed, io >> 2);
                        pe *= ((uint32_t)(he >> (16 * (ie & 3))) & 0xFFFFu) < drop_thr ? 0.f : inv_keep;
                        po *= ((uint32_t)(ho >> (16 * (io & 3))) & 0xFFFFu) < drop_thr ? 0.f : inv_keep;
                    }
                    const int kr = wrow0 + min(j, L - 1);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 v4 = *qa_kv_ptr(kv, kr, hl, 4 + c4);
                        oe[4 * c4] = fmaf(pe, v4.x, oe[4 * c4]); oo[4 * c4] = fmaf(po, v4.x, oo[4 * c4]);
                        oe[4 * c4 + 1] = fmaf(pe, v4.y, oe[4 * c4 + 1]); oo[4 * c4 + 1] = fmaf(po, v4.y, oo[4 * c4 + 1]);
                        oe[4 * c4 + 2] = fmaf(pe, v4.z, oe[4 * c4 + 2]); oo[4 * c4 + 2] = fmaf(po, v4.z, oo[4 * c4 + 2]);
                        oe[4 * c4 + 3] = fmaf(pe, v4.w, oe[4 * c4 + 3]); oo[4 * c4 + 3] = fmaf(po, v4.w, oo[4 * c4 + 3]);
                    }
                }
                me = mne; mo = mno;
            }
            if (live) {
                // the out-projection's A operand: FP16 hi/lo planes of 16 * o, feature index = head * 16 + d
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const float sc = ACT_SCALE / (rr ? lo : le);
                    const float* ov = rr ? oo : oe;
                    const size_t off = ((size_t)win * L + pos_e + rr) * E + (size_t)(g * QA_HG + hl) * HD;
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        uint32_t uh[4], ul[4];
#pragma unroll
                        for (int p2 = 0; p2 < 4; ++p2) {
                            float h0, h1, l0, l1;
                            veltkamp11(ov[8 * c8 + 2 * p2] * sc, h0, l0); veltkamp11(ov[8 * c8 + 2 * p2 + 1] * sc, h1, l1);
                            __half2 t2 = __floats2half2_rn(h0, h1); uh[p2] = *reinterpret_cast<uint32_t*>(&t2);
                            t2 = __floats2half2_rn(l0, l1); ul[p2] = *reinterpret_cast<uint32_t*>(&t2);
                        }
                        *reinterpret_cast<uint4*>(out_hi + off + 8 * c8) = make_uint4(uh[0], uh[1], uh[2], uh[3]);
                        *reinterpret_cast<uint4*>(out_lo + off + 8 * c8) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
                    }
                }
            }
        }
    }
    if (!pdl_early) griddep_launch();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, QA_TMEM_COLS);
    }
}

// weight rows / bias entries of one layer in the fused kernel's order: row g*192 + part*64 + hl*16 + d  <-  in_proj row
// part*256 + (4g + hl)*16 + d   (part: 0 q, 1 k, 2 v)
__device__ __forceinline__ int qa_src_row(int n) {
    const int g = n / QA_BN, r = n - g * QA_BN, part = r / (QA_HG * HD), hd = r - part * (QA_HG * HD);
    return part * E + g * (QA_HG * HD) + hd;
}
__global__ void pack_qkv_reorder_kernel(const __half* __restrict__ src_hi, const __half* __restrict__ src_lo,
                                        const float* __restrict__ src_bias, __half* __restrict__ dst_hi,
                                        __half* __restrict__ dst_lo, float* __restrict__ dst_bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (8 halves) of a row
    if (i >= 3 * E * (E / 8)) return;
    const int n = i / (E / 8), c = i - n * (E / 8);
    const int so = qa_src_row(n);
    reinterpret_cast<uint4*>(dst_hi)[(size_t)n * (E / 8) + c] = reinterpret_cast<const uint4*>(src_hi)[(size_t)so * (E / 8) + c];
    reinterpret_cast<uint4*>(dst_lo)[(size_t)n * (E / 8) + c] = reinterpret_cast<const uint4*>(src_lo)[(size_t)so * (E / 8) + c];
    if (c == 0) dst_bias[n] = src_bias[so];
}

}  // namespace tip
