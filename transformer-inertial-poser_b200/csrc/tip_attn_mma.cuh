// Causal multi-head self-attention core on the tensor cores (reference :85-91 ->
// nn.MultiheadAttention): P = softmax(q k^T + causal mask) [dropout], o = P v per (window, head),
// d = 16, L <= 40.  Used by the tcgen05 engine; the FFMA kernel in tip_simt.cuh is the cross-check.
//
// One warp per (window, head); 8 heads of a window per CTA.  Inputs are the FP16 hi/lo planes of
// 16*q, 16*k, 16*v written by the QKV GEMM epilogue (q already carries 1/sqrt(d)); both GEMMs of the
// head run as warp-level mma.sync.m16n8k16 (f16 in, f32 accumulate) with the same 3-product
// error-compensated split as the big GEMMs (hi*hi + hi*lo + lo*hi, ~22 significant bits):
//   S  (48 x 40) = Q K^T : 11 causal 16x8 tiles x 3 products, K = 16 = head dim -> one k-step
//   O  (48 x 16) = P V   : probabilities re-split to FP16 hi/lo in registers (the S accumulator
//                          fragment of two adjacent key tiles IS the A fragment of the next MMA);
//                          V fragments come from ldmatrix.trans on the row-major tile; Q fragments come straight
//                          from global memory
// Softmax statistics are fp32 on the accumulator fragments (quad shuffles).  The contraction is 2 %
// of the path's MACs; the kernel is bound by moving qkv in and o out, so everything goes through
// shared memory in 256-byte-per-row coalesced pieces.
#pragma once
#include "tip_common.cuh"

namespace tip {

constexpr int AM_VROWS = 48;                 // V rows (keys) padded to 3 k-steps of 16; rows >= L are zero
template <int HPB> struct AttnCfg {          // HPB heads per CTA (one warp each)
    static constexpr int ROWB = HPB * HD * 2 + 16;          // bytes per row of a plane tile (+16 B pad: bank spread)
    static constexpr int QK_BYTES = MAXL * ROWB;
    static constexpr int V_BYTES = AM_VROWS * ROWB;
    // K hi/lo + V hi/lo tiles; Q never touches shared memory (its fragments are 4-byte global loads that use every
    // byte of their 32-byte sectors), which keeps a CTA of 8 heads at 47.9 KB: FOUR CTAs per SM, so the 512 CTAs of a
    // 256-window batch are co-resident in one wave (with Q staged too it was 69.6 KB -> 3 per SM -> two waves).
    static constexpr int SMEM_BYTES = 2 * QK_BYTES + 2 * V_BYTES;
};

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> packed FP16 hi pair and lo pair (error-compensated split)
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int AM_HPB>
__global__ void __launch_bounds__(AM_HPB * 32, 1024 / (AM_HPB * 32) < 4 ? 1024 / (AM_HPB * 32) : 4)
attention_mma_kernel(const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo,
                     __half* __restrict__ out_hi, __half* __restrict__ out_lo, int L, float drop_p,
                     const uint64_t* __restrict__ seed_ptr, uint64_t seed_off, int b0, unsigned long long* tbuf = nullptr) {
#define AM_TS(i) do { if (tbuf && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tbuf[i] = t_; } } while (0)
    constexpr int AM_ROWB = AttnCfg<AM_HPB>::ROWB, AM_QK_BYTES = AttnCfg<AM_HPB>::QK_BYTES, AM_V_BYTES = AttnCfg<AM_HPB>::V_BYTES;
    constexpr int CPR = AM_HPB * 2;                  // 16-byte chunks per row of a plane tile
    extern __shared__ __align__(16) uint8_t am_smem[];
    griddep_wait();
    griddep_launch();
    AM_TS(0);
    uint8_t* sKh = am_smem;                          // [L][8 heads][16] halves
    uint8_t* sKl = sKh + AM_QK_BYTES;
    uint8_t* sVh = sKl + AM_QK_BYTES;                // [48 keys][8 heads][16] halves, rows >= L zero
    uint8_t* sVl = sVh + AM_V_BYTES;
    float* sO = reinterpret_cast<float*>(am_smem);   // [L][8 heads][16] fp32, aliases the K planes after they are consumed

    const int b = blockIdx.x, h0 = blockIdx.y * AM_HPB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const size_t rowbase = (size_t)b * L;

    // ---- stage K, V rows with cp.async (16 B = 8 dims of one head); V rows L..47 are zeroed ----
    for (int i = tid; i < L * CPR * 4; i += AM_HPB * 32) {         // (row, {Kh,Kl,Vh,Vl}, chunk)
        const int row = i / (4 * CPR), r = i - row * (4 * CPR), which = r / CPR, c = r - which * CPR;
        const __half* src = ((which & 1) ? qkv_lo : qkv_hi) + (rowbase + row) * (3 * E) + (1 + (which >> 1)) * E + h0 * HD + c * 8;
        uint8_t* dst = (which == 0 ? sKh : which == 1 ? sKl : which == 2 ? sVh : sVl) + row * AM_ROWB + c * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < (AM_VROWS - L) * CPR * 2; i += AM_HPB * 32) {
        const int plane = i & 1, c = (i >> 1) % CPR, row = L + (i >> 1) / CPR;
        *reinterpret_cast<uint4*>((plane ? sVl : sVh) + row * AM_ROWB + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- per warp: head hl ----
    const int hl = warp;
    // A fragments of Q (3 row tiles of 16) straight from the global planes; rows >= L are don't-care (clamped)
    uint32_t qa_hi[3][4], qa_lo[3][4];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        const int r0 = min(16 * mt + g, L - 1), r1 = min(16 * mt + g + 8, L - 1);
        const size_t o0 = (rowbase + r0) * (3 * E) + (h0 + hl) * HD + t4 * 2, o1 = (rowbase + r1) * (3 * E) + (h0 + hl) * HD + t4 * 2;
        qa_hi[mt][0] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o0));     qa_lo[mt][0] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o0));
        qa_hi[mt][1] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o1));     qa_lo[mt][1] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o1));
        qa_hi[mt][2] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o0 + 8)); qa_lo[mt][2] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o0 + 8));
        qa_hi[mt][3] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o1 + 8)); qa_lo[mt][3] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o1 + 8));
    }
    AM_TS(1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    AM_TS(2);
    uint32_t kb_hi[5][2], kb_lo[5][2];               // B fragments of K^T (5 key tiles of 8)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
        const int o = (8 * nt + g) * AM_ROWB + hl * 32 + t4 * 4;
        kb_hi[nt][0] = *reinterpret_cast<const uint32_t*>(sKh + o);      kb_lo[nt][0] = *reinterpret_cast<const uint32_t*>(sKl + o);
        kb_hi[nt][1] = *reinterpret_cast<const uint32_t*>(sKh + o + 16); kb_lo[nt][1] = *reinterpret_cast<const uint32_t*>(sKl + o + 16);
    }
    __syncthreads();                                 // every warp holds its K fragments: the K planes may become sO
    AM_TS(3);

    const float inv_keep = drop_inv_keep(drop_p);
    const uint32_t dthr = drop_threshold(drop_p);
    const uint64_t seed = drop_p > 0.f ? site_seed(seed_ptr, seed_off) : 0ull;
    // ldmatrix.trans row addresses of this lane: lanes 0-7 -> keys +0..7, lanes 8-15 -> keys +8..15
    const uint32_t vaddr_h = (uint32_t)__cvta_generic_to_shared(sVh) + (uint32_t)((lane & 15) * AM_ROWB + hl * 32);
    const uint32_t vaddr_l = (uint32_t)__cvta_generic_to_shared(sVl) + (uint32_t)((lane & 15) * AM_ROWB + hl * 32);
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        if (16 * mt >= L) break;                     // warp-uniform
        const int ra = 16 * mt + g, rb = ra + 8;     // this lane's two query rows
        // S tiles of the causal range: key tiles nt <= 2*mt + 1 (and < 5)
        float s[6][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            if (nt <= 2 * mt + 1 && nt < 5) {
                mma_f16_16816(s[nt], qa_lo[mt], kb_hi[nt][0], kb_hi[nt][1]);
                mma_f16_16816(s[nt], qa_hi[mt], kb_lo[nt][0], kb_lo[nt][1]);
                mma_f16_16816(s[nt], qa_hi[mt], kb_hi[nt][0], kb_hi[nt][1]);
            }
        }
        // mask (key <= query, key < L), un-scale (planes carry 16*q and 16*k), row max over the quad
        // (key tiles beyond the causal range of this row tile are skipped at compile time: their probabilities are 0)
        float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            if (!(nt <= 2 * mt + 1 && nt < 5)) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = 8 * nt + 2 * t4 + e;
                const bool live = key < L;
                s[nt][e] = (live && key <= ra) ? s[nt][e] * (1.f / (ACT_SCALE * ACT_SCALE)) : -INFINITY;
                s[nt][2 + e] = (live && key <= rb) ? s[nt][2 + e] * (1.f / (ACT_SCALE * ACT_SCALE)) : -INFINITY;
                ma = fmaxf(ma, s[nt][e]);
                mb = fmaxf(mb, s[nt][2 + e]);
            }
        }
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
        // exp, row sums (fp32, before dropout), optional attention dropout (element index: attn_drop_index)
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            if (!(nt <= 2 * mt + 1 && nt < 5)) continue;         // s[nt] stays 0 there: contributes nothing to P V
            // this lane's four probabilities of the tile (rows ra, rb = ra + 8; keys key0, key0 + 1) = one hash group: lanes
            // 0..3 of the hash = (ra, key0), (ra, key0 + 1), (rb, key0), (rb, key0 + 1).  A dropped probability becomes 0
            // (one compare + select on the hash word, the 16-bit lane in the upper half); the kept ones' 1 / (1 - p) is
            // applied once per row with the softmax normalisation below
            uint32_t hl32 = 0xFFFFFFFFu, hh32 = 0xFFFFFFFFu;
            if (drop_p > 0.f) {
                const uint64_t grp = attn_drop_index((uint64_t)(b0 + b) * NH + h0 + hl, ra, 8 * nt + 2 * t4) >> 2;
                const uint64_t hsh = hash_u64(seed, grp);
                hl32 = (uint32_t)hsh; hh32 = (uint32_t)(hsh >> 32);
            }
            const uint32_t thr_hi = dthr << 16;                          // (dthr <= 65535 whenever drop_p < 1)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float pa = exp2f((s[nt][e] - ma) * 1.4426950408889634f);        // exp(-inf) = 0 for masked slots
                float pb = exp2f((s[nt][2 + e] - mb) * 1.4426950408889634f);
                la += pa;
                lb += pb;
                if (drop_p > 0.f) {
                    if ((e ? hl32 : (hl32 << 16)) < thr_hi) pa = 0.f;
                    if ((e ? hh32 : (hh32 << 16)) < thr_hi) pb = 0.f;
                }
                s[nt][e] = pa;
                s[nt][2 + e] = pb;
            }
        }
        la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
        lb += __shfl_xor_sync(0xffffffffu, lb, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
        // O tile (16 x 16) = P V over the key steps ks <= mt (16 keys each)
        float o[2][4];
#pragma unroll
        for (int dn = 0; dn < 2; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            if (ks <= mt) {
                uint32_t pa_hi[4], pa_lo[4];
                split_pair(s[2 * ks][0], s[2 * ks][1], pa_hi[0], pa_lo[0]);
                split_pair(s[2 * ks][2], s[2 * ks][3], pa_hi[1], pa_lo[1]);
                split_pair(s[2 * ks + 1][0], s[2 * ks + 1][1], pa_hi[2], pa_lo[2]);
                split_pair(s[2 * ks + 1][2], s[2 * ks + 1][3], pa_hi[3], pa_lo[3]);
#pragma unroll
                for (int dn = 0; dn < 2; ++dn) {
                    // B fragment of V (k = 16 keys, n = 8 dims) straight from the row-major tile
                    uint32_t vh0, vh1, vl0, vl1;
                    const uint32_t off = (uint32_t)(16 * ks * AM_ROWB + dn * 16);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(vh0), "=r"(vh1) : "r"(vaddr_h + off));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(vl0), "=r"(vl1) : "r"(vaddr_l + off));
                    mma_f16_16816(o[dn], pa_lo, vh0, vh1);
                    mma_f16_16816(o[dn], pa_hi, vl0, vl1);
                    mma_f16_16816(o[dn], pa_hi, vh0, vh1);
                }
            }
        }
        // o_true = acc / (16 * l); the output planes carry 16 * o_true = acc / l
        const float ia = inv_keep / la, ib = inv_keep / lb;
#pragma unroll
        for (int dn = 0; dn < 2; ++dn) {
            if (ra < L) *reinterpret_cast<float2*>(sO + ((size_t)ra * AM_HPB + hl) * HD + 8 * dn + 2 * t4) = make_float2(o[dn][0] * ia, o[dn][1] * ia);
            if (rb < L) *reinterpret_cast<float2*>(sO + ((size_t)rb * AM_HPB + hl) * HD + 8 * dn + 2 * t4) = make_float2(o[dn][2] * ib, o[dn][3] * ib);
        }
    }
    __syncthreads();
    AM_TS(4);
    // ---- coalesced store: FP16 hi/lo planes of 16*o (the A operand of the out-projection GEMM) ----
    __half* oh = out_hi + rowbase * E + h0 * HD;
    __half* ol = out_lo + rowbase * E + h0 * HD;
    for (int i = tid; i < L * (AM_HPB * HD / 4); i += AM_HPB * 32) {
        const int row = i / (AM_HPB * HD / 4), c4 = i - row * (AM_HPB * HD / 4);
        const float4 v = reinterpret_cast<const float4*>(sO + (size_t)row * AM_HPB * HD)[c4];
        half_split_store4(oh + (size_t)row * E + c4 * 4, ol + (size_t)row * E + c4 * 4, v);
    }
    AM_TS(5);
#undef AM_TS
}

// Persistent, software-pipelined variant: a CTA walks the (window, head group) units u = blockIdx.x, += gridDim.x with TWO
// K / V tile sets in shared memory: the cp.async loads of the next unit are in flight while the current one is computed
// and stored, so an SM no longer runs load -> compute -> store in lock-step for all its resident CTAs, and the launch can
// be NARROW (a few CTAs per lane-share of the GPU) without losing per-SM efficiency.  Same arithmetic, same results.
template <int AM_HPB>
__global__ void __launch_bounds__(AM_HPB * 32, 2)
attention_mma_pipe_kernel(const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo,
                          __half* __restrict__ out_hi, __half* __restrict__ out_lo, int L, float drop_p,
                          const uint64_t* __restrict__ seed_ptr, uint64_t seed_off, int b0, int n_windows) {
    constexpr int AM_ROWB = AttnCfg<AM_HPB>::ROWB, AM_QK_BYTES = AttnCfg<AM_HPB>::QK_BYTES, AM_V_BYTES = AttnCfg<AM_HPB>::V_BYTES;
    constexpr int CPR = AM_HPB * 2;                  // 16-byte chunks per row of a plane tile
    extern __shared__ __align__(16) uint8_t am_smem[];
    griddep_wait();
    griddep_launch();
    constexpr int SET_BYTES = AttnCfg<AM_HPB>::SMEM_BYTES;        // one K / V tile set: [Kh | Kl | Vh | Vl]
    constexpr int GROUPS = NH / AM_HPB;                           // head groups per window
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int units = n_windows * GROUPS;

    // K, V rows of unit u -> tile set `base` with cp.async (16 B = 8 dims of one head)
    auto issue_loads = [&](int u, uint8_t* base) {
        const int ub = u / GROUPS, uh0 = (u - ub * GROUPS) * AM_HPB;
        const size_t rb = (size_t)ub * L;
        for (int i = tid; i < L * CPR * 4; i += AM_HPB * 32) {         // (row, {Kh,Kl,Vh,Vl}, chunk)
            const int row = i / (4 * CPR), r = i - row * (4 * CPR), which = r / CPR, c = r - which * CPR;
            const __half* src = ((which & 1) ? qkv_lo : qkv_hi) + (rb + row) * (3 * E) + (1 + (which >> 1)) * E + uh0 * HD + c * 8;
            uint8_t* dst = base + (which == 0 ? 0 : which == 1 ? AM_QK_BYTES : which == 2 ? 2 * AM_QK_BYTES : 2 * AM_QK_BYTES + AM_V_BYTES)
                           + row * AM_ROWB + c * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
    };
    // V rows L..47 of both tile sets are zero for the whole launch (the loads only write rows < L; sO aliases the K planes)
    for (int i = tid; i < 2 * (AM_VROWS - L) * CPR * 2; i += AM_HPB * 32) {
        const int set = i / ((AM_VROWS - L) * CPR * 2), j = i - set * ((AM_VROWS - L) * CPR * 2);
        const int plane = j & 1, c = (j >> 1) % CPR, row = L + (j >> 1) / CPR;
        *reinterpret_cast<uint4*>(am_smem + set * SET_BYTES + 2 * AM_QK_BYTES + plane * AM_V_BYTES + row * AM_ROWB + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if ((int)blockIdx.x < units) issue_loads(blockIdx.x, am_smem);
    asm volatile("cp.async.commit_group;" ::: "memory");

    int iter = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++iter) {
    uint8_t* sKh = am_smem + (iter & 1) * SET_BYTES;  // [L][8 heads][16] halves
    uint8_t* sKl = sKh + AM_QK_BYTES;
    uint8_t* sVh = sKl + AM_QK_BYTES;                // [48 keys][8 heads][16] halves, rows >= L zero
    uint8_t* sVl = sVh + AM_V_BYTES;
    float* sO = reinterpret_cast<float*>(sKh);       // [L][8 heads][16] fp32, aliases the K planes after they are consumed
    const int b = u / GROUPS, h0 = (u - b * GROUPS) * AM_HPB;
    const size_t rowbase = (size_t)b * L;
    // the next unit's tiles -> the other set (free since the barrier that ended the previous iteration)
    if (u + (int)gridDim.x < units) issue_loads(u + gridDim.x, am_smem + ((iter + 1) & 1) * SET_BYTES);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- per warp: head hl ----
    const int hl = warp;
    // A fragments of Q (3 row tiles of 16) straight from the global planes; rows >= L are don't-care (clamped)
    uint32_t qa_hi[3][4], qa_lo[3][4];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        const int r0 = min(16 * mt + g, L - 1), r1 = min(16 * mt + g + 8, L - 1);
        const size_t o0 = (rowbase + r0) * (3 * E) + (h0 + hl) * HD + t4 * 2, o1 = (rowbase + r1) * (3 * E) + (h0 + hl) * HD + t4 * 2;
        qa_hi[mt][0] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o0));     qa_lo[mt][0] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o0));
        qa_hi[mt][1] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o1));     qa_lo[mt][1] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o1));
        qa_hi[mt][2] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o0 + 8)); qa_lo[mt][2] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o0 + 8));
        qa_hi[mt][3] = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + o1 + 8)); qa_lo[mt][3] = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + o1 + 8));
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // this unit's tiles have landed (the next unit's may be in flight)
    __syncthreads();
    uint32_t kb_hi[5][2], kb_lo[5][2];               // B fragments of K^T (5 key tiles of 8)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
        const int o = (8 * nt + g) * AM_ROWB + hl * 32 + t4 * 4;
        kb_hi[nt][0] = *reinterpret_cast<const uint32_t*>(sKh + o);      kb_lo[nt][0] = *reinterpret_cast<const uint32_t*>(sKl + o);
        kb_hi[nt][1] = *reinterpret_cast<const uint32_t*>(sKh + o + 16); kb_lo[nt][1] = *reinterpret_cast<const uint32_t*>(sKl + o + 16);
    }
    __syncthreads();                                 // every warp holds its K fragments: the K planes may become sO

    const float inv_keep = drop_inv_keep(drop_p);
    const uint32_t dthr = drop_threshold(drop_p);
    const uint64_t seed = drop_p > 0.f ? site_seed(seed_ptr, seed_off) : 0ull;
    // ldmatrix.trans row addresses of this lane: lanes 0-7 -> keys +0..7, lanes 8-15 -> keys +8..15
    const uint32_t vaddr_h = (uint32_t)__cvta_generic_to_shared(sVh) + (uint32_t)((lane & 15) * AM_ROWB + hl * 32);
    const uint32_t vaddr_l = (uint32_t)__cvta_generic_to_shared(sVl) + (uint32_t)((lane & 15) * AM_ROWB + hl * 32);
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        if (16 * mt >= L) break;                     // warp-uniform
        const int ra = 16 * mt + g, rb = ra + 8;     // this lane's two query rows
        // S tiles of the causal range: key tiles nt <= 2*mt + 1 (and < 5)
        float s[6][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            if (nt <= 2 * mt + 1 && nt < 5) {
                mma_f16_16816(s[nt], qa_lo[mt], kb_hi[nt][0], kb_hi[nt][1]);
                mma_f16_16816(s[nt], qa_hi[mt], kb_lo[nt][0], kb_lo[nt][1]);
                mma_f16_16816(s[nt], qa_hi[mt], kb_hi[nt][0], kb_hi[nt][1]);
            }
        }
        // mask (key <= query, key < L), un-scale (planes carry 16*q and 16*k), row max over the quad
        // (key tiles beyond the causal range of this row tile are skipped at compile time: their probabilities are 0)
        float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            if (!(nt <= 2 * mt + 1 && nt < 5)) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = 8 * nt + 2 * t4 + e;
                const bool live = key < L;
                s[nt][e] = (live && key <= ra) ? s[nt][e] * (1.f / (ACT_SCALE * ACT_SCALE)) : -INFINITY;
                s[nt][2 + e] = (live && key <= rb) ? s[nt][2 + e] * (1.f / (ACT_SCALE * ACT_SCALE)) : -INFINITY;
                ma = fmaxf(ma, s[nt][e]);
                mb = fmaxf(mb, s[nt][2 + e]);
            }
        }
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
        // exp, row sums (fp32, before dropout), optional attention dropout (element index: attn_drop_index)
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            if (!(nt <= 2 * mt + 1 && nt < 5)) continue;         // s[nt] stays 0 there: contributes nothing to P V
            // this lane's four probabilities of the tile (rows ra, rb = ra + 8; keys key0, key0 + 1) = one hash group: lanes
            // 0..3 of the hash = (ra, key0), (ra, key0 + 1), (rb, key0), (rb, key0 + 1).  A dropped probability becomes 0
            // (one compare + select on the hash word, the 16-bit lane in the upper half); the kept ones' 1 / (1 - p) is
            // applied once per row with the softmax normalisation below
            uint32_t hl32 = 0xFFFFFFFFu, hh32 = 0xFFFFFFFFu;
            if (drop_p > 0.f) {
                const uint64_t grp = attn_drop_index((uint64_t)(b0 + b) * NH + h0 + hl, ra, 8 * nt + 2 * t4) >> 2;
                const uint64_t hsh = hash_u64(seed, grp);
                hl32 = (uint32_t)hsh; hh32 = (uint32_t)(hsh >> 32);
            }
            const uint32_t thr_hi = dthr << 16;                          // (dthr <= 65535 whenever drop_p < 1)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float pa = exp2f((s[nt][e] - ma) * 1.4426950408889634f);        // exp(-inf) = 0 for masked slots
                float pb = exp2f((s[nt][2 + e] - mb) * 1.4426950408889634f);
                la += pa;
                lb += pb;
                if (drop_p > 0.f) {
                    if ((e ? hl32 : (hl32 << 16)) < thr_hi) pa = 0.f;
                    if ((e ? hh32 : (hh32 << 16)) < thr_hi) pb = 0.f;
                }
                s[nt][e] = pa;
                s[nt][2 + e] = pb;
            }
        }
        la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
        lb += __shfl_xor_sync(0xffffffffu, lb, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
        // O tile (16 x 16) = P V over the key steps ks <= mt (16 keys each)
        float o[2][4];
#pragma unroll
        for (int dn = 0; dn < 2; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            if (ks <= mt) {
                uint32_t pa_hi[4], pa_lo[4];
                split_pair(s[2 * ks][0], s[2 * ks][1], pa_hi[0], pa_lo[0]);
                split_pair(s[2 * ks][2], s[2 * ks][3], pa_hi[1], pa_lo[1]);
                split_pair(s[2 * ks + 1][0], s[2 * ks + 1][1], pa_hi[2], pa_lo[2]);
                split_pair(s[2 * ks + 1][2], s[2 * ks + 1][3], pa_hi[3], pa_lo[3]);
#pragma unroll
                for (int dn = 0; dn < 2; ++dn) {
                    // B fragment of V (k = 16 keys, n = 8 dims) straight from the row-major tile
                    uint32_t vh0, vh1, vl0, vl1;
                    const uint32_t off = (uint32_t)(16 * ks * AM_ROWB + dn * 16);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(vh0), "=r"(vh1) : "r"(vaddr_h + off));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(vl0), "=r"(vl1) : "r"(vaddr_l + off));
                    mma_f16_16816(o[dn], pa_lo, vh0, vh1);
                    mma_f16_16816(o[dn], pa_hi, vl0, vl1);
                    mma_f16_16816(o[dn], pa_hi, vh0, vh1);
                }
            }
        }
        // o_true = acc / (16 * l); the output planes carry 16 * o_true = acc / l
        const float ia = inv_keep / la, ib = inv_keep / lb;
#pragma unroll
        for (int dn = 0; dn < 2; ++dn) {
            if (ra < L) *reinterpret_cast<float2*>(sO + ((size_t)ra * AM_HPB + hl) * HD + 8 * dn + 2 * t4) = make_float2(o[dn][0] * ia, o[dn][1] * ia);
            if (rb < L) *reinterpret_cast<float2*>(sO + ((size_t)rb * AM_HPB + hl) * HD + 8 * dn + 2 * t4) = make_float2(o[dn][2] * ib, o[dn][3] * ib);
        }
    }
    __syncthreads();
    // ---- coalesced store: FP16 hi/lo planes of 16*o (the A operand of the out-projection GEMM) ----
    __half* oh = out_hi + rowbase * E + h0 * HD;
    __half* ol = out_lo + rowbase * E + h0 * HD;
    for (int i = tid; i < L * (AM_HPB * HD / 4); i += AM_HPB * 32) {
        const int row = i / (AM_HPB * HD / 4), c4 = i - row * (AM_HPB * HD / 4);
        const float4 v = reinterpret_cast<const float4*>(sO + (size_t)row * AM_HPB * HD)[c4];
        half_split_store4(oh + (size_t)row * E + c4 * 4, ol + (size_t)row * E + c4 * 4, v);
    }
    __syncthreads();                                 // sO (this tile set) is free for the loads of the unit after next
    }
}


}  // namespace tip
