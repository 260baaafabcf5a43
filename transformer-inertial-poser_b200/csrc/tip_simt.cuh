// FP32 (FFMA) kernels of the TIP hot path: input conditioning, a register-tiled NT GEMM with fused
// bias / ReLU / dropout / residual+LayerNorm epilogues, the warp-cooperative causal attention, the
// tanh-RNN recurrence and the window shift.  These serve small M = B*L (the per-frame B=1 path) and
// every stage that is not a big GEMM; large-M GEMMs go to the tcgen05 kernels in tip_umma.cuh.
#pragma once
#include "tip_common.cuh"

namespace tip {

// ------------------------------------------------------------------------------------------------
// Input conditioning (reference simple_transformer_with_state.py:63-78): clone, NaN->0 on x_s,
// dropout(in_dropout) on x_imu, zero root velocity (also folded into the packed weight),
// dropout(past_state_dropout) on x_s -- or an explicit keep-mask -- and the concat, written as
// one (M, kin_pad) matrix (zero padded) in fp32 or FP16 hi/lo planes.
__device__ __forceinline__ float condition_value(const float* __restrict__ x_imu, const float* __restrict__ x_s,
                                                 const float* __restrict__ keep_mask, float past_scale, int r, int c,
                                                 int n_imu, int size_s, float f_in, float f_past) {
    float v = 0.f;
    if (c < n_imu) {
        v = __ldg(x_imu + (int64_t)r * n_imu + c) * f_in;                            // :73
    } else if (c < n_imu + size_s) {
        const int cs = c - n_imu;
        v = __ldg(x_s + (int64_t)r * size_s + cs);
        if (v != v) v = 0.f;                               // :65  x_s[isnan] = 0
        if (cs >= 108 && cs < 111) v = 0.f;                // :75  root velocity removed
        if (keep_mask != nullptr) v *= __ldg(keep_mask + (int64_t)r * size_s + cs) * past_scale;
        else v *= f_past;                                  // :77
    }
    return v;
}
// one thread produces 8 consecutive columns of a row (kin_pad is a multiple of 64): 16-byte stores into
// the FP16 hi/lo planes (scale 1: raw model input) of the tcgen05 engine, or fp32 for the FFMA engine.
// Dropout element index = row * kin_pad + column of the concatenated (padded) input row (row0 = first global row).
__global__ void condition_kernel(const float* __restrict__ x_imu, const float* __restrict__ x_s,
                                 const float* __restrict__ keep_mask, float past_scale,
                                 float* __restrict__ out, float* __restrict__ out_lo,
                                 int M, int n_imu, int size_s, int kin_pad,
                                 float p_in, float p_past, const uint64_t* __restrict__ seed_ptr, uint64_t seed_off, int row0) {
    griddep_wait();
    griddep_launch();
    const int groups = kin_pad >> 3;
    const int64_t total = (int64_t)M * groups;
    const uint32_t thr_in = drop_threshold(p_in), thr_past = keep_mask ? 0u : drop_threshold(p_past);
    const float inv_in = drop_inv_keep(p_in), inv_past = drop_inv_keep(p_past);
    const uint64_t seed = (thr_in | thr_past) ? site_seed(seed_ptr, seed_off) : 0ull;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(gi / groups);
        const int c0 = (int)(gi - (int64_t)r * groups) << 3;
        float fin[8], fpast[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { fin[j] = 1.f; fpast[j] = 1.f; }
        const uint64_t g0 = ((uint64_t)(row0 + r) * kin_pad + c0) >> 2;              // two groups of four elements
        if (thr_in && c0 < n_imu) {
            const float4 a = dropout_factor4(seed + SEED_IN, g0, thr_in, inv_in), b = dropout_factor4(seed + SEED_IN, g0 + 1, thr_in, inv_in);
            fin[0] = a.x; fin[1] = a.y; fin[2] = a.z; fin[3] = a.w; fin[4] = b.x; fin[5] = b.y; fin[6] = b.z; fin[7] = b.w;
        }
        if (thr_past && c0 + 8 > n_imu && c0 < n_imu + size_s) {
            const float4 a = dropout_factor4(seed + SEED_PAST, g0, thr_past, inv_past), b = dropout_factor4(seed + SEED_PAST, g0 + 1, thr_past, inv_past);
            fpast[0] = a.x; fpast[1] = a.y; fpast[2] = a.z; fpast[3] = a.w; fpast[4] = b.x; fpast[5] = b.y; fpast[6] = b.z; fpast[7] = b.w;
        }
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            v[j] = condition_value(x_imu, x_s, keep_mask, past_scale, r, c0 + j, n_imu, size_s, fin[j], fpast[j]);
        const int64_t o = (int64_t)r * kin_pad + c0;
        if (out_lo != nullptr) {
            half_split_store4(reinterpret_cast<__half*>(out) + o, reinterpret_cast<__half*>(out_lo) + o,
                              make_float4(v[0], v[1], v[2], v[3]));
            half_split_store4(reinterpret_cast<__half*>(out) + o + 4, reinterpret_cast<__half*>(out_lo) + o + 4,
                              make_float4(v[4], v[5], v[6], v[7]));
        } else {
            *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(out + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// one thread: the base seed of this call's dropout masks -> device memory (outside any captured graph)
__global__ void seed_set_kernel(uint64_t* __restrict__ dst, uint64_t seed) { *dst = seed; }

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).  A and W are K-contiguous (nn.Linear layout).
struct Epi {
    const float* bias;      // [N]
    const float* resid;     // LN variant: residual rows [M][ldr]: fp32, or (resid_lo != null) the __half hi plane
    const float* resid_lo;  // __half lo plane of the residual (planes hold ACT_SCALE * x)
    int ldr;
    const float* gamma;     // LN affine
    const float* beta;
    float* out;             // fp32 output, or (out_lo != null) the __half hi plane of ACT_SCALE * output
    float* out_lo;          // __half lo plane
    const float* acc_scale; // tcgen05 engine: device scalar multiplying the accumulator (1/(s_a*s_w))
    int ldc;
    int relu;
    float drop_p;           // dropout on the GEMM output (after ReLU; before the residual add); element index = row * N + col
    uint32_t drop_thr;      // = drop_threshold(drop_p), drop_inv = drop_inv_keep(drop_p): set by the host (tcgen05 engine reads them from the constant bank)
    float drop_inv;
    const uint64_t* seed_ptr;   // base seed of the call (device memory; null = 0)
    uint64_t seed;              // + site offset (tip_common.cuh seed_out / seed_ff1 / seed_ff2 [+ chunk]); rows are batch-global
    int dbg;                // experiment flags (TIP_DBG env): 1 skip stores, 2 skip bias loads
    unsigned long long* tbuf;   // optional phase timestamps of CTA 0 (TIP_DBG & 4)
    int tma_out;            // tcgen05 engine: the output goes through TMA store boxes
    int pdl_early;          // programmatic dependent launch: trigger the next kernel at the start (small forwards) or at the end
    int* sched;             // tcgen05 engine, plain tiles: {next tile, CTAs done} counters of the dynamic tile scheduler (null = static round-robin)
};

constexpr int SG_BK = 16;

template <int BM, int BN, int TM, bool LN>
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                int M, int N, int K, Epi ep) {
    constexpr int TN = 8;
    constexpr int NT = 256;
    constexpr int TX = BN / TN;            // threads along N
    static_assert((BM / TM) * TX == NT, "tile/thread mismatch");
    static_assert(!LN || TX == 32, "LN epilogue needs one warp per row group");
    constexpr int RG = TM / 4;             // row groups of 4 per thread (1 or 2)
    __shared__ __align__(16) float As[2][SG_BK][BM + 4];
    __shared__ __align__(16) float Ws[2][SG_BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    constexpr int A_F4 = BM * SG_BK / 4, W_F4 = BN * SG_BK / 4;
    constexpr int A_PER = (A_F4 + NT - 1) / NT, W_PER = (W_F4 + NT - 1) / NT;
    float4 ra[A_PER], rw[W_PER];

    auto gload = [&](int kt) {
        const int k0 = kt * SG_BK;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * NT;
            const int row = idx >> 2, kq = idx & 3;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < A_F4 && m0 + row < M)
                ra[i] = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * lda + k0 + kq * 4));
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            const int idx = tid + i * NT;
            const int row = idx >> 2, kq = idx & 3;
            rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < W_F4 && n0 + row < N)
                rw[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + row) * ldw + k0 + kq * 4));
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * NT;
            if (idx < A_F4) {
                const int row = idx >> 2, kq = idx & 3;
                As[buf][kq * 4 + 0][row] = ra[i].x;
                As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z;
                As[buf][kq * 4 + 3][row] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            const int idx = tid + i * NT;
            if (idx < W_F4) {
                const int row = idx >> 2, kq = idx & 3;
                Ws[buf][kq * 4 + 0][row] = rw[i].x;
                Ws[buf][kq * 4 + 1][row] = rw[i].y;
                Ws[buf][kq * 4 + 2][row] = rw[i].z;
                Ws[buf][kq * 4 + 3][row] = rw[i].w;
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = K / SG_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < RG; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / 2) + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][g * (BN / 2) + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    const float inv_keep = drop_inv_keep(ep.drop_p);
    const uint64_t dseed = ep.drop_p > 0.f ? site_seed(ep.seed_ptr, ep.seed) : 0ull;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
        const bool row_ok = row < M;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + (j / 4) * (BN / 2) + tx * 4 + (j & 3);
            float x = acc[i][j] + ((col < N) ? __ldg(ep.bias + col) : 0.f);
            if (ep.relu) x = fmaxf(x, 0.f);
            if (ep.drop_p > 0.f) x *= dropout_factor(ep.drop_p, inv_keep, dseed, (uint64_t)row * N + col);
            v[j] = x;
        }
        if constexpr (LN) {
            // residual add + LayerNorm over the full 256-wide row held by this warp
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int col = (j / 4) * (BN / 2) + tx * 4 + (j & 3);
                float r = 0.f;
                if (row_ok) {
                    r = ep.resid[(size_t)row * ep.ldr + col];
                }
                v[j] += r;
                s += v[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.f / BN);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * (1.f / BN) + 1e-5f);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int col = (j / 4) * (BN / 2) + tx * 4 + (j & 3);
                v[j] = (v[j] - mean) * rstd * __ldg(ep.gamma + col) + __ldg(ep.beta + col);
            }
        }
        if (row_ok) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int col = n0 + g * (BN / 2) + tx * 4;
                float* o = ep.out + (size_t)row * ep.ldc + col;
                if (col + 3 < N && (ep.ldc & 3) == 0) {
                    *reinterpret_cast<float4*>(o) = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < N) o[j] = v[g * 4 + j];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Causal multi-head self-attention core (reference :85-91 -> nn.MultiheadAttention):
//   P = softmax(q k^T + causal mask) [dropout], o = P v, per (window b, head h); 1/sqrt(d) is folded
//   into W_q at pack time.  The contraction is tiny (d=16, L<=40): FFMA + warp shuffles, no tensor
//   cores; the kernel is bound by moving qkv in and o out, so both go through shared memory in
//   512-byte-per-row coalesced pieces (HPB=8 heads of one window per CTA).
//   One thread group of KS lanes owns the query-row pair (p, L-1-p), so every group does the same
//   L+1 keys of causal work; the KS lanes split the keys and merge (max, sum, o) with shuffles.
//   The lane's scores stay in registers between the max and the exp/accumulate sweeps.
template <int KS, int HPB>
__global__ void __launch_bounds__(HPB * 20 * KS)
attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ out_lo,
                 int L, float drop_p, const uint64_t* __restrict__ seed_ptr, uint64_t seed_off, int b0) {
    constexpr int NT = HPB * 20 * KS;
    __shared__ __align__(16) float Qs[MAXL][HPB][HD];      // reused for the output tile
    __shared__ __align__(16) float Ks[MAXL][HPB][HD];
    __shared__ __align__(16) float Vs[MAXL][HPB][HD];
    const int b = blockIdx.x;
    const int h0 = blockIdx.y * HPB;
    const int tid = threadIdx.x;
    const float* base = qkv + (size_t)b * L * (3 * E) + h0 * HD;

    // stage Q, K, V of HPB heads: per row 3 pieces of HPB*16 contiguous floats, global -> shared with
    // cp.async (no register staging, all of a thread's 16-byte copies in flight at once)
    constexpr int F4_ROW = HPB * HD / 4;
    for (int i = tid; i < L * 3 * F4_ROW; i += NT) {
        const int row = i / (3 * F4_ROW);
        const int rem = i - row * (3 * F4_ROW);
        const int which = rem / F4_ROW, c4 = rem - which * F4_ROW;
        const float* src = base + (size_t)row * (3 * E) + which * E + c4 * 4;
        float* dst = (which == 0 ? &Qs[row][0][0] : (which == 1 ? &Ks[row][0][0] : &Vs[row][0][0])) + c4 * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int hl = tid / (20 * KS);
    const int rem = tid - hl * (20 * KS);
    const int p = rem / KS, ks = rem % KS;
    const int npairs = (L + 1) >> 1;
    const bool active = p < npairs;
    const int ra = active ? p : 0;                 // shorter row
    const int rb = active ? L - 1 - p : 0;         // longer row (ra <= rb)
    const bool two = active && (ra != rb);

    float qa[HD], qb[HD];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 va = *reinterpret_cast<const float4*>(&Qs[ra][hl][i * 4]);
        const float4 vb = *reinterpret_cast<const float4*>(&Qs[rb][hl][i * 4]);
        qa[i * 4 + 0] = va.x; qa[i * 4 + 1] = va.y; qa[i * 4 + 2] = va.z; qa[i * 4 + 3] = va.w;
        qb[i * 4 + 0] = vb.x; qb[i * 4 + 1] = vb.y; qb[i * 4 + 2] = vb.z; qb[i * 4 + 3] = vb.w;
    }
    const int jend = active ? rb : -1;
    // scores of this lane's keys (j = jj*KS + ks) stay in registers; masked / unused slots hold -inf
    constexpr int NJ = (MAXL + KS - 1) / KS;
    float sa[NJ], sb[NJ];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int j = jj * KS + ks;
        sa[jj] = -INFINITY;
        sb[jj] = -INFINITY;
        if (j <= jend) {
            float da = 0.f, db = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 kv = *reinterpret_cast<const float4*>(&Ks[j][hl][i * 4]);
                da = fmaf(qa[i * 4 + 0], kv.x, da); db = fmaf(qb[i * 4 + 0], kv.x, db);
                da = fmaf(qa[i * 4 + 1], kv.y, da); db = fmaf(qb[i * 4 + 1], kv.y, db);
                da = fmaf(qa[i * 4 + 2], kv.z, da); db = fmaf(qb[i * 4 + 2], kv.z, db);
                da = fmaf(qa[i * 4 + 3], kv.w, da); db = fmaf(qb[i * 4 + 3], kv.w, db);
            }
            sb[jj] = db;
            mb = fmaxf(mb, db);
            if (j <= ra) { sa[jj] = da; ma = fmaxf(ma, da); }
        }
    }
#pragma unroll
    for (int o = KS >> 1; o > 0; o >>= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    // exp, row sums, P v   (expf(-inf - m) = 0 for the masked slots)
    float oa[HD], ob[HD], la = 0.f, lb = 0.f;
#pragma unroll
    for (int i = 0; i < HD; ++i) { oa[i] = 0.f; ob[i] = 0.f; }
    const float inv_keep = drop_inv_keep(drop_p);
    const uint64_t seed = drop_p > 0.f ? site_seed(seed_ptr, seed_off) : 0ull;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int j = jj * KS + ks;
        if (j <= jend) {
            float pa = expf(sa[jj] - ma);
            float pb = expf(sb[jj] - mb);
            la += pa;
            lb += pb;
            if (drop_p > 0.f) {   // attention-probability dropout (train mode only)
                const uint64_t ida = attn_drop_index((uint64_t)(b0 + b) * NH + h0 + hl, ra, j);
                const uint64_t idb = attn_drop_index((uint64_t)(b0 + b) * NH + h0 + hl, rb, j);
                pa *= dropout_factor(drop_p, inv_keep, seed, ida);
                pb *= dropout_factor(drop_p, inv_keep, seed, idb);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][hl][i * 4]);
                oa[i * 4 + 0] = fmaf(pa, vv.x, oa[i * 4 + 0]); ob[i * 4 + 0] = fmaf(pb, vv.x, ob[i * 4 + 0]);
                oa[i * 4 + 1] = fmaf(pa, vv.y, oa[i * 4 + 1]); ob[i * 4 + 1] = fmaf(pb, vv.y, ob[i * 4 + 1]);
                oa[i * 4 + 2] = fmaf(pa, vv.z, oa[i * 4 + 2]); ob[i * 4 + 2] = fmaf(pb, vv.z, ob[i * 4 + 2]);
                oa[i * 4 + 3] = fmaf(pa, vv.w, oa[i * 4 + 3]); ob[i * 4 + 3] = fmaf(pb, vv.w, ob[i * 4 + 3]);
            }
        }
    }
#pragma unroll
    for (int o = KS >> 1; o > 0; o >>= 1) {
        la += __shfl_xor_sync(0xffffffffu, la, o);
        lb += __shfl_xor_sync(0xffffffffu, lb, o);
#pragma unroll
        for (int i = 0; i < HD; ++i) {
            oa[i] += __shfl_xor_sync(0xffffffffu, oa[i], o);
            ob[i] += __shfl_xor_sync(0xffffffffu, ob[i], o);
        }
    }
    // normalised rows -> the Q tile (each row of Qs is read and written by its own lane group only)
    if (active && ks == 0) {
        const float ia = 1.f / la, ib = 1.f / lb;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(&Qs[ra][hl][i * 4]) =
                make_float4(oa[i * 4 + 0] * ia, oa[i * 4 + 1] * ia, oa[i * 4 + 2] * ia, oa[i * 4 + 3] * ia);
            if (two)
                *reinterpret_cast<float4*>(&Qs[rb][hl][i * 4]) =
                    make_float4(ob[i * 4 + 0] * ib, ob[i * 4 + 1] * ib, ob[i * 4 + 2] * ib, ob[i * 4 + 3] * ib);
        }
    }
    __syncthreads();
    // coalesced store (fp32, or FP16 hi/lo planes of ACT_SCALE*o for the tcgen05 out-projection)
    float* ob_ = out + (size_t)b * L * E + h0 * HD;
    __half* oh_ = reinterpret_cast<__half*>(out) + (size_t)b * L * E + h0 * HD;
    __half* ol_ = out_lo ? reinterpret_cast<__half*>(out_lo) + (size_t)b * L * E + h0 * HD : nullptr;
    for (int i = tid; i < L * F4_ROW; i += NT) {
        const int row = i / F4_ROW, c4 = i - row * F4_ROW;
        const float4 v = reinterpret_cast<const float4*>(&Qs[row][0][0])[c4];
        if (ol_) {
            const size_t e = (size_t)row * E + c4 * 4;
            half_split_store4(oh_ + e, ol_ + e,
                              make_float4(v.x * ACT_SCALE, v.y * ACT_SCALE, v.z * ACT_SCALE, v.w * ACT_SCALE));
        } else {
            reinterpret_cast<float4*>(ob_ + (size_t)row * E)[c4] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tanh RNN recurrence (reference :95-99): h_t = tanh(gi[b,t,:] + W_hh h_{t-1}), h_0 = 0, where
// gi = x W_ih^T + (b_ih + b_hh) comes from the preceding GEMM.  Baseline kernel: one CTA carries RB
// windows through all L steps; thread n owns hidden unit n and streams column n of W_hh^T from L2
// (coalesced across the CTA); h_{t-1} lives in shared memory.  The serial chain is the only
// inherently sequential stage of the path.
template <int RB>
__global__ void __launch_bounds__(R)
rnn_stream_kernel(const float* __restrict__ gi, const float* __restrict__ whh_t,
                  float* __restrict__ hs, float* __restrict__ hs_lo, int B, int L) {
    __shared__ __align__(16) float h[2][RB][R];
    const int n = threadIdx.x;
    const int b0 = blockIdx.x * RB;
#pragma unroll
    for (int r = 0; r < RB; ++r) h[0][r][n] = 0.f;
    __syncthreads();
    for (int t = 0; t < L; ++t) {
        const int cur = t & 1;
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r)
            acc[r] = (b0 + r < B) ? gi[((size_t)(b0 + r) * L + t) * R + n] : 0.f;
        if (t > 0) {
#pragma unroll 4
            for (int k = 0; k < R; k += 4) {
                const float w0 = __ldg(whh_t + (size_t)(k + 0) * R + n);
                const float w1 = __ldg(whh_t + (size_t)(k + 1) * R + n);
                const float w2 = __ldg(whh_t + (size_t)(k + 2) * R + n);
                const float w3 = __ldg(whh_t + (size_t)(k + 3) * R + n);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float4 hv = *reinterpret_cast<const float4*>(&h[cur][r][k]);
                    acc[r] = fmaf(hv.x, w0, acc[r]);
                    acc[r] = fmaf(hv.y, w1, acc[r]);
                    acc[r] = fmaf(hv.z, w2, acc[r]);
                    acc[r] = fmaf(hv.w, w3, acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const float v = tanhf(acc[r]);
            h[cur ^ 1][r][n] = v;
            if (b0 + r < B) {
                const size_t o = ((size_t)(b0 + r) * L + t) * R + n;
                if (hs_lo) {
                    __half hi, lo;
                    half_split(v * ACT_SCALE, hi, lo);
                    reinterpret_cast<__half*>(hs)[o] = hi;
                    reinterpret_cast<__half*>(hs_lo)[o] = lo;
                } else {
                    hs[o] = v;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// tanh RNN recurrence, cluster version (the production kernel).  The recurrence is the only serial
// chain of the path (40 dependent 512x512 mat-vecs per window), so the 1 MB W_hh must not be
// re-streamed from L2 every step: a thread-block cluster of 8 CTAs keeps W_hh resident IN REGISTERS
// (CTA c owns hidden units [64c, 64c+64); thread (u, kq) holds the 128 weights of row u for the
// k-slices kq and kq+4), the hidden state lives in every CTA's shared memory as 8 slices of 64
// units, and after each step a CTA pushes its slice to the 7 peers with one asynchronous bulk
// DSMEM copy each (cp.async.bulk.shared::cluster, 2 KB), completion counted on an mbarrier in the
// destination CTA -- no cluster-wide barrier inside the time loop.  A cluster carries up to 16
// windows at once as two groups of 8: while group A's slices are in flight the cluster computes
// group B (windows are independent chains), which hides the exchange latency.  FFMA fp32
// throughout (parity with the reference's fp32 RNN).
constexpr int RC_CTAS = 8;                       // CTAs per cluster
constexpr int RC_UNITS = R / RC_CTAS;            // 64 hidden units per CTA
constexpr int RC_GROUP = 8;                      // windows per group
constexpr int RC_MAXG = 3;                       // groups a cluster interleaves
constexpr int RC_ROWS = RC_MAXG * RC_GROUP;      // max windows per cluster pass
constexpr int RC_RPITCH = 72;                    // floats per (slice, window): 2 x (32 + 4 pad)
constexpr int RC_SLICE = RC_ROWS * RC_RPITCH + 8; // floats per CTA-slice (+32 B: bank spread)
constexpr int RC_BUF = RC_CTAS * RC_SLICE;       // floats per h buffer
constexpr int RC_GROUP_BYTES = RC_GROUP * RC_RPITCH * (int)sizeof(float);     // 2304 B per push
constexpr int RC_SMEM_BYTES = 2 * RC_BUF * (int)sizeof(float) + 128;

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_bulk_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void rc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rc_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rc_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();        // protocol bug: fail loudly, never hang
    }
}

// 32 partial sums per lane, 16 lanes (lane & 15) hold different k-slices: butterfly with halving.
// Afterwards lane l holds the two complete sums with index 2*(l&15) and 2*(l&15)+1.
__device__ __forceinline__ void rc_reduce16(const float (&v)[32], int lane, float& o0, float& o1) {
    float a[16], b[8], c[4];
    const bool u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float send = u8 ? v[i] : v[i + 16];
        a[i] = (u8 ? v[i + 16] : v[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = u4 ? a[i] : a[i + 8];
        b[i] = (u4 ? a[i + 8] : a[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = u2 ? b[i] : b[i + 4];
        c[i] = (u2 ? b[i + 4] : b[i]) + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const float s0 = u1 ? c[0] : c[2], s1 = u1 ? c[1] : c[3];
        o0 = (u1 ? c[2] : c[0]) + __shfl_xor_sync(0xffffffffu, s0, 1);
        o1 = (u1 ? c[3] : c[1]) + __shfl_xor_sync(0xffffffffu, s1, 1);
    }
}

// Thread (ug, s): unit group ug = tid/16 (4 hidden units), k-slice s = tid%16 (32 k).  Its 128 weights
// stay in registers for the whole launch; per k-chunk of 4 it loads one float4 of h per window and
// issues 16 FMAs with it (4 units), so shared-memory bandwidth (a 128-bit load costs 4 crossbar
// cycles) and the FMA pipes are balanced.
__global__ void __cluster_dims__(RC_CTAS, 1, 1) __launch_bounds__(256, 1)
rnn_cluster_kernel(const float* __restrict__ gi, const float* __restrict__ whh,
                   float* __restrict__ hs, float* __restrict__ hs_lo, int B, int L, int rows_per_pass) {
    extern __shared__ __align__(128) float hbuf[];          // [2 buffers][8 slices][24 windows][72] (+pad)
    const int tid = threadIdx.x, lane = tid & 31;
    const int ug = tid >> 4, s = tid & 15;
    const uint32_t rank = cluster_ctarank();
    const int unit0 = (int)rank * RC_UNITS + ug * 4;         // first of this thread's 4 units
    const int n_clusters = gridDim.x / RC_CTAS, cluster_id = blockIdx.x / RC_CTAS;
    const uint32_t hbuf_s = (uint32_t)__cvta_generic_to_shared(hbuf);
    const uint32_t bar_s = hbuf_s + 2u * RC_BUF * 4u;        // 2*RC_MAXG mbarriers: [group][buffer]

    float w[4][32];                                          // W_hh[unit0 + u][32 s + kk]
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4* wp = reinterpret_cast<const float4*>(whh + (size_t)(unit0 + u) * R + s * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(wp + j);
            w[u][4 * j] = t.x; w[u][4 * j + 1] = t.y; w[u][4 * j + 2] = t.z; w[u][4 * j + 3] = t.w;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 2 * RC_MAXG; ++i) rc_mbar_init(bar_s + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_arrive();
    cluster_wait();
    uint32_t fills = 0;                                      // bit (g*2+p): parity of the next fill to wait for

    // after the reduction this lane owns window fr (of the group) and units fu, fu+1 (of the CTA)
    const int fr = (lane & 15) >> 1;
    const int fu = ug * 4 + 2 * (lane & 1);
    // k position 64*rank + fu inside a window's h row -> slice `rank`, half fu/32, offset fu%32
    const int h_local = (fu >> 5) * 36 + (fu & 31);
    // where this thread reads h: k-slice s lives in CTA-slice s/2, half s%2
    const int h_read = (s >> 1) * RC_SLICE + (s & 1) * 36;

    // windows are dealt to clusters in blocks of `rows_per_pass` (host-chosen so that one pass covers the
    // batch whenever B <= clusters * RC_ROWS)
    for (int rb = cluster_id; rb * rows_per_pass < B; rb += n_clusters) {
        const int b0 = rb * rows_per_pass;
        const int nrows = min(rows_per_pass, B - b0);
        const int ngroups = (nrows + RC_GROUP - 1) / RC_GROUP;   // uniform over the cluster
        for (int t = 0; t < L; ++t) {
            const int cur = t & 1, nxt = cur ^ 1;
            for (int g = 0; g < ngroups; ++g) {
                const int nr = min(RC_GROUP, nrows - g * RC_GROUP);
                // arm the barrier that will collect the peers' h_g(t) (its previous fill, h_g(t-2), was
                // consumed during step t-1; early complete_tx from fast peers is legal)
                if (tid == 0 && t + 1 < L) rc_mbar_expect(bar_s + 8u * (g * 2 + nxt), (RC_CTAS - 1) * RC_GROUP_BYTES);
                const int wrow = g * RC_GROUP + fr;          // window (within the block) this lane finishes
                float2 giv = make_float2(0.f, 0.f);
                if (fr < nr) giv = __ldg(reinterpret_cast<const float2*>(gi + ((size_t)(b0 + wrow) * L + t) * R + rank * RC_UNITS + fu));
                float o0 = 0.f, o1 = 0.f;
                if (t > 0) {
                    // h_g(t-1): 7 peer slices (mbarrier) + the local slice (the __syncthreads below)
                    const int bi = g * 2 + cur;
                    rc_mbar_wait(bar_s + 8u * bi, (fills >> bi) & 1u);
                    fills ^= 1u << bi;
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    const float* hb = hbuf + cur * RC_BUF + h_read + (g * RC_GROUP) * RC_RPITCH;
#pragma unroll
                    for (int kc = 0; kc < 8; ++kc) {
#pragma unroll
                        for (int r = 0; r < RC_GROUP; ++r) {
                            const float4 hv = *reinterpret_cast<const float4*>(hb + r * RC_RPITCH + 4 * kc);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                float a = acc[r * 4 + u];
                                a = fmaf(w[u][4 * kc], hv.x, a);
                                a = fmaf(w[u][4 * kc + 1], hv.y, a);
                                a = fmaf(w[u][4 * kc + 2], hv.z, a);
                                a = fmaf(w[u][4 * kc + 3], hv.w, a);
                                acc[r * 4 + u] = a;
                            }
                        }
                    }
                    rc_reduce16(acc, lane, o0, o1);          // -> window fr, units fu, fu+1
                }
                const float va = tanhf(o0 + giv.x), vb = tanhf(o1 + giv.y);
                float* mine = hbuf + nxt * RC_BUF + (int)rank * RC_SLICE + wrow * RC_RPITCH + h_local;
                mine[0] = va;                                // windows beyond nr carry don't-care values
                mine[1] = vb;
                if (fr < nr) {
                    const size_t o = ((size_t)(b0 + wrow) * L + t) * R + rank * RC_UNITS + fu;
                    if (hs_lo) {
                        __half h0, l0, h1, l1;
                        half_split(va * ACT_SCALE, h0, l0); half_split(vb * ACT_SCALE, h1, l1);
                        *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(hs) + o) = __halves2half2(h0, h1);
                        *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(hs_lo) + o) = __halves2half2(l0, l1);
                    } else {
                        *reinterpret_cast<float2*>(hs + o) = make_float2(va, vb);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
                __syncthreads();
                if (t + 1 < L && tid < RC_CTAS && (uint32_t)tid != rank) {
                    const uint32_t off = (uint32_t)(nxt * RC_BUF + (int)rank * RC_SLICE + (g * RC_GROUP) * RC_RPITCH) * 4u;
                    dsmem_bulk_push(map_to_cta(hbuf_s + off, (uint32_t)tid), hbuf_s + off, RC_GROUP_BYTES,
                                    map_to_cta(bar_s + 8u * (g * 2 + nxt), (uint32_t)tid));
                }
            }
        }
        // next row block reuses the buffers: everyone must be done reading first
        cluster_arrive();
        cluster_wait();
    }
}

// ------------------------------------------------------------------------------------------------
// Small-batch recurrence (streaming: a handful of windows).  Same decomposition as rnn_cluster_kernel -- a cluster
// of 8 CTAs, W_hh in registers (thread = 4 hidden units x one 32-wide k-slice) -- but built for the latency of ONE
// step, which is all that matters when NW <= 8 windows share a cluster:
//   * every CTA keeps the FULL h vector of its NW windows in shared memory (double-buffered over the step parity);
//   * the 16 lanes that finish 4 units each send them straight from registers to all 8 CTAs with
//     st.async.shared::cluster (16 bytes per store, bytes counted on the destination's mbarrier): no staging tile,
//     no proxy fence, no CTA barrier and no bulk-copy engine between tanh and the peers' next step.
template <int NW>
__global__ void __cluster_dims__(RC_CTAS, 1, 1) __launch_bounds__(256, 1)
rnn_small_kernel(const float* __restrict__ gi, const float* __restrict__ whh,
                 float* __restrict__ hs, float* __restrict__ hs_lo, int B, int L, unsigned long long* tbuf) {
    __shared__ __align__(16) float hbuf[2][NW][R];
    __shared__ __align__(8) uint64_t full_bar[2];
    const int tid = threadIdx.x;
    const int ug = tid >> 4, s = tid & 15;
    const uint32_t rank = cluster_ctarank();
    const int unit0 = (int)rank * RC_UNITS + ug * 4;
    const int n_clusters = gridDim.x / RC_CTAS, cluster_id = blockIdx.x / RC_CTAS;
    // k-slice s = the float4 chunks 16 j + s, j = 0..7 (interleaved, so that the 16 lanes of a half warp read 16
    // CONSECUTIVE float4 of h: conflict-free; contiguous 32-float slices would make every load a 16-way conflict)
    float w[4][32];                                          // W_hh[unit0 + u][4 (16 j + s) + e]  (weights: before the PDL wait)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4* wp = reinterpret_cast<const float4*>(whh + (size_t)(unit0 + u) * R) + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t4 = __ldg(wp + 16 * j);
            w[u][4 * j] = t4.x; w[u][4 * j + 1] = t4.y; w[u][4 * j + 2] = t4.z; w[u][4 * j + 3] = t4.w;
        }
    }
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&full_bar[0]);
    const uint32_t hbuf_s = (uint32_t)__cvta_generic_to_shared(&hbuf[0][0][0]);
    if (tid == 0) {
        rc_mbar_init(bar_s, 1);
        rc_mbar_init(bar_s + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_arrive();
    cluster_wait();
    griddep_wait();
    griddep_launch();
    uint32_t par0 = 0, par1 = 0;
    for (int b0 = cluster_id * NW; b0 < B; b0 += n_clusters * NW) {
        for (int t = 0; t < L; ++t) {
            const int cur = t & 1, nxt = cur ^ 1;
            // this step's h_t lands in buffer nxt of every CTA: 512 floats per window from the 8 CTAs (self included)
            if (tid == 0 && t + 1 < L) rc_mbar_expect(bar_s + 8u * nxt, (uint32_t)(NW * R * sizeof(float)));
            float4 g4[NW];
            if (s == 0) {
#pragma unroll
                for (int wi = 0; wi < NW; ++wi)
                    g4[wi] = __ldg(reinterpret_cast<const float4*>(gi + ((size_t)min(b0 + wi, B - 1) * L + t) * R + unit0));
            }
            float acc[NW][4];
#pragma unroll
            for (int wi = 0; wi < NW; ++wi) acc[wi][0] = acc[wi][1] = acc[wi][2] = acc[wi][3] = 0.f;
            if (t > 0) {
                if (cur) { rc_mbar_wait(bar_s + 8u, par1); par1 ^= 1u; }
                else     { rc_mbar_wait(bar_s, par0); par0 ^= 1u; }
                if (tbuf && blockIdx.x == 0 && tid == 0 && (t == 20 || t == 21)) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tbuf[(t - 20) * 4]));
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {
#pragma unroll
                    for (int wi = 0; wi < NW; ++wi) {
                        const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][wi][4 * (16 * kc + s)]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float a = acc[wi][u];
                            a = fmaf(w[u][4 * kc], hv.x, a);
                            a = fmaf(w[u][4 * kc + 1], hv.y, a);
                            a = fmaf(w[u][4 * kc + 2], hv.z, a);
                            a = fmaf(w[u][4 * kc + 3], hv.w, a);
                            acc[wi][u] = a;
                        }
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {                 // sum over the 16 k-slices (lanes of a half warp)
#pragma unroll
                    for (int wi = 0; wi < NW; ++wi)
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc[wi][u] += __shfl_xor_sync(0xffffffffu, acc[wi][u], o);
                }
            }
            if (tbuf && blockIdx.x == 0 && tid == 0 && t == 20) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tbuf[1]));
            if (s == 0) {
#pragma unroll
                for (int wi = 0; wi < NW; ++wi) {
                    float4 h4;
                    h4.x = tanhf(acc[wi][0] + g4[wi].x); h4.y = tanhf(acc[wi][1] + g4[wi].y);
                    h4.z = tanhf(acc[wi][2] + g4[wi].z); h4.w = tanhf(acc[wi][3] + g4[wi].w);
                    if (t + 1 < L) {
                        const uint32_t dst = hbuf_s + (uint32_t)(((nxt * NW + wi) * R + unit0) * sizeof(float));
#pragma unroll
                        for (uint32_t c = 0; c < RC_CTAS; ++c) {
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                         ::"r"(map_to_cta(dst, c)), "r"(__float_as_uint(h4.x)), "r"(__float_as_uint(h4.y)),
                                           "r"(__float_as_uint(h4.z)), "r"(__float_as_uint(h4.w)),
                                           "r"(map_to_cta(bar_s + 8u * nxt, c)) : "memory");
                        }
                    }
                    if (b0 + wi < B) {
                        const size_t o = ((size_t)(b0 + wi) * L + t) * R + unit0;
                        if (hs_lo) {
                            half_split_store4(reinterpret_cast<__half*>(hs) + o, reinterpret_cast<__half*>(hs_lo) + o,
                                              make_float4(h4.x * ACT_SCALE, h4.y * ACT_SCALE, h4.z * ACT_SCALE, h4.w * ACT_SCALE));
                        } else {
                            *reinterpret_cast<float4*>(hs + o) = h4;
                        }
                    }
                }
            }
            if (tbuf && blockIdx.x == 0 && tid == 0 && t == 20) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tbuf[2]));
        }
        // the next block of windows restarts at parity 0 with fresh barriers' phases in step; make sure every CTA has
        // consumed its last buffer before anyone overwrites it
        cluster_arrive();
        cluster_wait();
    }
}

// ------------------------------------------------------------------------------------------------
// Per-frame window update of the streaming path (replaces the runner's per-frame re-assembly,
// real_time_runner_minimal.py:131-147): when a stream's window is full, slide it up by one row
// (in place: every thread first reads its elements, the CTA synchronises, then writes them one row
// earlier -- coalesced both ways), then append the new row.  One CTA per stream and window.
__global__ void __launch_bounds__(256)
window_push_kernel(float* __restrict__ win0, const float* __restrict__ new_row0, int width0,
                   float* __restrict__ win1, const float* __restrict__ new_row1, int width1, int len) {
    griddep_wait();
    griddep_launch();
    // blockIdx.y selects the window set (0: IMU rows, 1: state rows) -- both windows of a frame in one launch.
    // win: (S, MAXL, width); len = rows currently held (same for every stream)
    float* win = blockIdx.y ? win1 : win0;
    const float* new_row = blockIdx.y ? new_row1 : new_row0;
    const int width = blockIdx.y ? width1 : width0;
    float* w = win + (size_t)blockIdx.x * MAXL * width;
    const float* nr = new_row + (size_t)blockIdx.x * width;
    constexpr int PER = (MAXL * 160 + 255) / 256;   // width <= 160
    if (len == MAXL) {
        float v[PER];
        const int n = (MAXL - 1) * width;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = threadIdx.x + i * 256;
            v[i] = idx < n ? w[width + idx] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = threadIdx.x + i * 256;
            if (idx < n) w[idx] = v[i];
        }
        for (int c = threadIdx.x; c < width; c += 256) w[(MAXL - 1) * width + c] = nr[c];
    } else {
        for (int c = threadIdx.x; c < width; c += 256) w[len * width + c] = nr[c];
    }
}

// ------------------------------------------------------------------------------------------------
// Row N1 of SURVEY.md 8f -- the runner's per-frame IMU pre-processing on the device
// (real_time_runner_minimal.py:59-76 record_raw_imu, :131-141 window features; data_utils.py:190-219
// imu_rotate_to_local; constants.py:15-18):
//   raw (72,) = 6 global rotations (54) + 6 global accelerations (18), appended to an 11-frame ring
//   (the very first frame is replicated 5 extra times, :60-63).  Once 11 raw frames exist, the new
//   model row is  [R_root | R_root^-1 R_i (5) | a_root | R_root^-1 a_i (5) | acc-sum / 15]  with the
//   rotations taken from 5 frames ago, the accelerations averaged over the 11 raw frames, and acc-sum =
//   sum of the rotated accelerations of the last <= 40 rows.  The reference keeps these buffers in
//   float64 and casts the window to float32; the arithmetic here is done in double for the same
//   rounding.  One CTA (one warp) per stream; rows already written never change, so only the newest
//   row is produced (the reference recomputes the whole window every frame).
constexpr int IMU_RAW = 72, IMU_RING = 11, IMU_DELAY = 5, ACC_WIN = 40;
__global__ void __launch_bounds__(32)
imu_push_kernel(const float* __restrict__ raw_new, float* __restrict__ raw_ring, double* __restrict__ acc_ring,
                float* __restrict__ row_out, int n_imu, int n_raw_before, int n_rows_before) {
    griddep_wait();
    griddep_launch();
    const int s = blockIdx.x, lane = threadIdx.x;
    float* ring = raw_ring + (size_t)s * IMU_RING * IMU_RAW;
    double* aring = acc_ring + (size_t)s * ACC_WIN * 18;
    const float* nr = raw_new + (size_t)s * IMU_RAW;
    // append (first frame: 5 pads + itself)
    const int n_app = n_raw_before == 0 ? IMU_DELAY + 1 : 1;
    for (int a = 0; a < n_app; ++a)
        for (int c = lane; c < IMU_RAW; c += 32) ring[((n_raw_before + a) % IMU_RING) * IMU_RAW + c] = nr[c];
    __syncwarp();
    const int n_raw = n_raw_before + n_app;
    if (n_raw < IMU_RING) return;                              // no smoothed frame yet (runner :125-128)
    __shared__ double sm[IMU_RAW];                             // smoothed frame: rotations (54) + mean accelerations (18)
    const float* rot_src = ring + ((n_raw - 1 - IMU_DELAY) % IMU_RING) * IMU_RAW;     // raw[-IMU_n_smooth - 1]
    for (int c = lane; c < 54; c += 32) sm[c] = (double)rot_src[c];
    if (lane < 18) {
        double acc = 0.0;
        for (int k = 0; k < IMU_RING; ++k)                     // chronological, like np.mean over axis 0
            acc += (double)ring[((n_raw - IMU_RING + k) % IMU_RING) * IMU_RAW + 54 + lane];
        sm[54 + lane] = acc / (double)IMU_RING;
    }
    __syncwarp();
    // inverse of the root rotation (np.linalg.inv, data_utils.py:196): adjugate / determinant
    double r[9], inv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = sm[i];
    inv[0] = r[4] * r[8] - r[5] * r[7]; inv[1] = r[2] * r[7] - r[1] * r[8]; inv[2] = r[1] * r[5] - r[2] * r[4];
    inv[3] = r[5] * r[6] - r[3] * r[8]; inv[4] = r[0] * r[8] - r[2] * r[6]; inv[5] = r[2] * r[3] - r[0] * r[5];
    inv[6] = r[3] * r[7] - r[4] * r[6]; inv[7] = r[1] * r[6] - r[0] * r[7]; inv[8] = r[0] * r[4] - r[1] * r[3];
    const double det = r[0] * inv[0] + r[1] * inv[3] + r[2] * inv[6];
#pragma unroll
    for (int i = 0; i < 9; ++i) inv[i] /= det;
    float* out = row_out + (size_t)s * n_imu;
    double* anew = aring + (n_rows_before % ACC_WIN) * 18;
    // 72 outputs: [0:9] root R, [9:54] inv*R_i, [54:57] root acc, [57:72] inv*a_i
    for (int c = lane; c < IMU_RAW; c += 32) {
        double v;
        if (c < 9) v = sm[c];
        else if (c < 54) {
            const int i = (c - 9) / 9, e = (c - 9) % 9, rr = e / 3, cc = e % 3;
            const double* Ri = sm + 9 + 9 * i;
            v = inv[rr * 3 + 0] * Ri[0 * 3 + cc] + inv[rr * 3 + 1] * Ri[1 * 3 + cc] + inv[rr * 3 + 2] * Ri[2 * 3 + cc];
        } else if (c < 57) v = sm[c];
        else {
            const int i = (c - 57) / 3, rr = (c - 57) % 3;
            const double* ai = sm + 57 + 3 * i;
            v = inv[rr * 3 + 0] * ai[0] + inv[rr * 3 + 1] * ai[1] + inv[rr * 3 + 2] * ai[2];
        }
        out[c] = (float)v;
        if (c >= 54) anew[c - 54] = v;
    }
    __syncwarp();
    if (n_imu > IMU_RAW && lane < 18) {                        // acc-sum feature (:134-141), /ACC_SUM_DOWN_SCALE
        const int cnt = min(n_rows_before + 1, ACC_WIN);
        double acc = 0.0;
        for (int k = 0; k < cnt; ++k)                          // chronological over the last <= 40 rows
            acc += aring[((n_rows_before + 1 - cnt + k) % ACC_WIN) * 18 + lane];
        out[IMU_RAW + lane] = (float)(acc / 15.0);
    }
}

// ------------------------------------------------------------------------------------------------
// Un-fused residual + LayerNorm of the skinny-M path (tcgen05 engine, M <= a few row tiles): the GEMM wrote
// pre = acc + bias as fp32 [rows][256]; out = LN(pre + residual) * gamma + beta, residual and output being FP16 hi/lo
// planes of ACT_SCALE * x.  One warp per row, 8 columns per lane, fully coalesced.
__global__ void __launch_bounds__(256)
resid_ln_kernel(const float* __restrict__ pre, const __half* __restrict__ res_hi, const __half* __restrict__ res_lo,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                __half* __restrict__ out_hi, __half* __restrict__ out_lo, int row0, int rows) {
    griddep_wait();
    griddep_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= rows) return;
    const size_t base = (size_t)(row0 + r) * E + lane * 8;
    const float4 p0 = *reinterpret_cast<const float4*>(pre + base), p1 = *reinterpret_cast<const float4*>(pre + base + 4);
    const uint4 h4 = *reinterpret_cast<const uint4*>(res_hi + base), l4 = *reinterpret_cast<const uint4*>(res_lo + base);
    float x[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
        x[2 * i] = fmaf(hf.x + lf.x, 1.f / ACT_SCALE, x[2 * i]);
        x[2 * i + 1] = fmaf(hf.y + lf.y, 1.f / ACT_SCALE, x[2 * i + 1]);
        sum += x[2 * i] + x[2 * i + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / E) + 1e-5f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    __half hh[8], ll[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) half_split(fmaf((x[i] - mean) * rstd, gg[i], bb[i]) * ACT_SCALE, hh[i], ll[i]);
    *reinterpret_cast<uint4*>(out_hi + base) = *reinterpret_cast<const uint4*>(hh);
    *reinterpret_cast<uint4*>(out_lo + base) = *reinterpret_cast<const uint4*>(ll);
}

// ------------------------------------------------------------------------------------------------
// Row N3: the model-visible part of RTRunnerMin.step AFTER the model call, on the device
// (real_time_runner_minimal.py:87-112 smooth_and_split_s_c, :150-167 state assembly, :78-85/:196
// record_state_aa_and_c; data_utils.py:164-187; fairmotion A2R / R2A = scipy Rotation).  One warp per
// stream, double arithmetic like the reference's float64 buffers -- except during the first five model
// calls, where the reference works on the raw float32 output row itself (:98 returns it un-copied, so
// the normalisation / cross product of data_utils.py:170-172 run in float32 and :107-110 edit the
// buffered row in place); both quirks are kept.  PyBullet FK and the SBP root correction (:169-194)
// only produce the root translation, which is never fed back, and stay on the CPU.
//   out_state[s] = [ s_t[3:60] (root aa from the IMU, 17 joint aa, root velocity) | c_t | root_v ]   (60 + n_c doubles;
//                  root_v = the filtered, not yet averaged velocity that :159 integrates into the root position)
//   fb_row[s]    = the (size_s,) float row record_state_aa_and_c appends for the next call.
constexpr int PP_TAPS = 6, PP_NJ = 18, PP_TAIL = 54;

// nearest rotation (orthogonal Procrustes, what scipy's Rotation.from_matrix applies to non-orthogonal
// input): Newton iteration X <- (X + X^-T) / 2 on the polar factor, then Markley's matrix->quaternion
// and the rotation-vector log map of scipy's as_rotvec.
__device__ inline void pp_rotmat_to_aa(const double* M, double* aa) {
    double X[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) X[i] = M[i];
    for (int it = 0; it < 60; ++it) {
        double C[9];
        C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
        C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
        C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
        const double det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
        if (!(fabs(det) > 1e-300)) break;
        double delta = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double nx = 0.5 * (X[i] + C[i] / det);     // cofactor / det = X^-T
            delta = fmax(delta, fabs(nx - X[i]));
            X[i] = nx;
        }
        if (delta < 1e-15) break;
    }
    const double d0 = X[0], d1 = X[4], d2 = X[8], tr = d0 + d1 + d2;
    double q[4];
    int c = 0;
    double best = d0;
    if (d1 > best) { best = d1; c = 1; }
    if (d2 > best) { best = d2; c = 2; }
    if (tr > best) c = 3;
    if (c != 3) {
        const int i = c, j = (c + 1) % 3, k = (c + 2) % 3;
        q[i] = 1.0 - tr + 2.0 * X[i * 3 + i];
        q[j] = X[j * 3 + i] + X[i * 3 + j];
        q[k] = X[k * 3 + i] + X[i * 3 + k];
        q[3] = X[k * 3 + j] - X[j * 3 + k];
    } else {
        q[0] = X[7] - X[5]; q[1] = X[2] - X[6]; q[2] = X[3] - X[1]; q[3] = 1.0 + tr;
    }
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double sgn = (q[3] < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = sgn * q[i] / qn;
    const double vn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double angle = 2.0 * atan2(vn, q[3]);
    double scale;
    if (angle <= 1e-3) { const double a2 = angle * angle; scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0; }
    else scale = angle / sin(angle / 2.0);
    aa[0] = q[0] * scale; aa[1] = q[1] * scale; aa[2] = q[2] * scale;
}

// first two columns of the rotation matrix of a rotation vector (scipy from_rotvec().as_matrix())
__device__ inline void pp_aa_to_rot6(const double* a, double* r6) {
    const double angle = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double scale;
    if (angle <= 1e-3) { const double a2 = angle * angle; scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0; }
    else scale = sin(angle / 2.0) / angle;
    const double x = a[0] * scale, y = a[1] * scale, z = a[2] * scale, w = cos(angle / 2.0);
    r6[0] = 1.0 - 2.0 * (y * y + z * z); r6[1] = 2.0 * (x * y - z * w);
    r6[2] = 2.0 * (x * y + z * w);       r6[3] = 1.0 - 2.0 * (x * x + z * z);
    r6[4] = 2.0 * (x * z - y * w);       r6[5] = 2.0 * (y * z + x * w);
}

__global__ void __launch_bounds__(32)
post_step_kernel(const float* __restrict__ y_last, const float* __restrict__ imu_rows, int n_imu,
                 double* __restrict__ ring, double* __restrict__ last_tail, float* __restrict__ fb_row,
                 double* __restrict__ out_state, int size_s, int n_before) {
    griddep_wait();
    griddep_launch();
    const int s = blockIdx.x, lane = threadIdx.x;
    const int n_c = size_s - 111, out_w = 60 + n_c;
    double* rg = ring + (size_t)s * PP_TAPS * size_s;
    double* lt = last_tail + (size_t)s * PP_TAIL;
    const float* y = y_last + (size_t)s * size_s;
    double* out = out_state + (size_t)s * out_w;
    float* fb = fb_row + (size_t)s * size_s;
    __shared__ double sm[160];          // smoothed row
    __shared__ double st[57];           // s_t[3:60]
    const bool mode32 = (n_before + 1) < PP_TAPS;
    const int slot = n_before % PP_TAPS;
    for (int c = lane; c < size_s; c += 32) rg[slot * size_s + c] = (double)y[c];
    __syncwarp();
    if (!mode32) {
        // np.sum(buf[-6:] * coeff[:, None], axis=0) / np.sum(coeff), coeff = 0.6 ** [5..0]      (:93-96)
        const double cf[PP_TAPS] = {0.07776, 0.1296, 0.216, 0.36, 0.6, 1.0};
        double csum = 0.0;
#pragma unroll
        for (int k = 0; k < PP_TAPS; ++k) csum += pow(0.6, (double)(PP_TAPS - 1 - k));
        for (int c = lane; c < size_s; c += 32) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < PP_TAPS; ++k)
                acc += rg[((n_before + 1 - PP_TAPS + k) % PP_TAPS) * size_s + c] * pow(0.6, (double)(PP_TAPS - 1 - k));
            (void)cf;
            sm[c] = acc / csum;
        }
    } else {
        for (int c = lane; c < size_s; c += 32) sm[c] = (double)y[c];
    }
    __syncwarp();
    // SBP block: flag = logit > 0, offsets / 5                                                    (:103-110)
    for (int c = 111 + lane; c < size_s; c += 32) {
        const int k = (c - 111) & 3;
        double v = sm[c];
        if (k == 0) v = v > 0.0 ? 1.0 : 0.0;
        else v = mode32 ? (double)__fdiv_rn((float)v, 5.0f) : v / 5.0;
        sm[c] = v;
        if (mode32) rg[slot * size_s + c] = v;            // the reference edits the buffered row in place
    }
    __syncwarp();
    // 2-axis -> rotation vector per joint                                        (data_utils.py:164-179)
    if (lane < PP_NJ) {
        double M[9];
        const double* v = sm + 6 * lane;                  // (3, 2) row-major: columns a1 = v[0,2,4], a2 = v[1,3,5]
        if (mode32) {
            const float x0 = (float)v[0], x1 = (float)v[2], x2 = (float)v[4];
            const float y0 = (float)v[1], y1 = (float)v[3], y2 = (float)v[5];
            const float n1 = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2))), 1e-6f);
            const float n2 = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1)), __fmul_rn(y2, y2))), 1e-6f);
            const float a0 = __fdiv_rn(x0, n1), a1 = __fdiv_rn(x1, n1), a2 = __fdiv_rn(x2, n1);
            const float b0 = __fdiv_rn(y0, n2), b1 = __fdiv_rn(y1, n2), b2 = __fdiv_rn(y2, n2);
            const float c0 = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
            const float c1 = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
            const float c2 = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
            M[0] = a0; M[1] = b0; M[2] = c0; M[3] = a1; M[4] = b1; M[5] = c1; M[6] = a2; M[7] = b2; M[8] = c2;
        } else {
            const double n1 = sqrt(v[0] * v[0] + v[2] * v[2] + v[4] * v[4]) + 1e-6;
            const double n2 = sqrt(v[1] * v[1] + v[3] * v[3] + v[5] * v[5]) + 1e-6;
            const double a0 = v[0] / n1, a1 = v[2] / n1, a2 = v[4] / n1;
            const double b0 = v[1] / n2, b1 = v[3] / n2, b2 = v[5] / n2;
            M[0] = a0; M[1] = b0; M[2] = a1 * b2 - a2 * b1;
            M[3] = a1; M[4] = b1; M[5] = a2 * b0 - a0 * b2;
            M[6] = a2; M[7] = b2; M[8] = a0 * b1 - a1 * b0;
        }
        double aa[3];
        if (lane == 0) {                                  // root: taken from the IMU, not the prediction (:161-163)
            const float* rr = imu_rows + (size_t)s * n_imu;
#pragma unroll
            for (int i = 0; i < 9; ++i) M[i] = (double)rr[i];
        }
        pp_rotmat_to_aa(M, aa);
        st[3 * lane] = aa[0]; st[3 * lane + 1] = aa[1]; st[3 * lane + 2] = aa[2];
    }
    if (lane >= 29) { st[54 + lane - 29] = sm[108 + lane - 29]; out[57 + n_c + lane - 29] = sm[108 + lane - 29]; }   // root velocity (:154-159)
    __syncwarp();
    // "To make motion a bit smoother": average s_t[6:] with the previous frame's                  (:165-167)
    for (int c = lane; c < PP_TAIL; c += 32) {
        double v = st[3 + c];
        if (n_before > 0) v = (v + lt[c]) / 2.0;
        lt[c] = v;
        st[3 + c] = v;
    }
    __syncwarp();
    for (int c = lane; c < 57; c += 32) out[c] = st[c];
    for (int c = lane; c < n_c; c += 32) { out[57 + c] = sm[111 + c]; fb[111 + c] = (float)sm[111 + c]; }
    if (lane < PP_NJ) {                                   // record_state_aa_and_c                   (:78-85)
        double r6[6];
        pp_aa_to_rot6(st + 3 * lane, r6);
#pragma unroll
        for (int i = 0; i < 6; ++i) fb[6 * lane + i] = (float)r6[i];
    }
    if (lane >= 29) fb[108 + lane - 29] = (float)st[54 + lane - 29];
}

// gather compacted (S, L, width) windows out of the (S, MAXL, width) storage when L < MAXL
__global__ void window_compact_kernel(const float* __restrict__ win, float* __restrict__ out,
                                      int S, int L, int width) {
    griddep_wait();
    griddep_launch();
    const int64_t total = (int64_t)S * L * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / ((int64_t)L * width));
        const int64_t rem = i - (int64_t)s * L * width;
        out[i] = win[(size_t)s * MAXL * width + rem];
    }
}

// y_last[b, :] = y[b, L-1, :]
__global__ void last_row_kernel(const float* __restrict__ y, float* __restrict__ y_last,
                                int B, int L, int size_s) {
    griddep_wait();
    griddep_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * size_s) {
        const int b = i / size_s, c = i - b * size_s;
        y_last[i] = y[((size_t)b * L + (L - 1)) * size_s + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Weight packing (pack time only).
__global__ void pack_in_linear_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ wp, float* __restrict__ bp,
                                      int d_in, int kin_pad, int n_imu) {
    // new row j*NH + h  <-  old row h*HD + j   (reference :88-89 folded into the weight);
    // columns of the root velocity (x_s[108:111]) zeroed (:75); zero pad to kin_pad.
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * kin_pad) return;
    const int rn = i / kin_pad, c = i - rn * kin_pad;
    const int j = rn / NH, h = rn % NH;
    const int ro = h * HD + j;
    float v = 0.f;
    if (c < d_in && !(c >= n_imu + 108 && c < n_imu + 111)) v = w[(size_t)ro * d_in + c];
    wp[i] = v;
    if (c == 0) bp[rn] = b[ro];
}
__global__ void pack_scale_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                       int64_t n, int64_t n_scaled, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (i < n_scaled) ? src[i] * scale : src[i];
}
__global__ void pack_add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i] + b[i];
}
__global__ void pack_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    // dst[k][j] = src[j][k], n x n
    __shared__ float t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int x = threadIdx.x, y = threadIdx.y;
    t[y][x] = src[(size_t)(by + y) * n + bx + x];
    __syncthreads();
    dst[(size_t)(bx + y) * n + by + x] = t[x][y];
}
__global__ void pack_pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                     int rows, int rows_pad, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows_pad * cols) dst[i] = (i / cols < rows) ? src[i] : 0.f;
}
// max |w| of a matrix -> power-of-two scale s_w with s_w * max|w| in [16384, 32768); writes
// scales[idx] = 1 / (s_w * act_scale) for the GEMM epilogue and scales_w[idx] = s_w for the split.
__global__ void __launch_bounds__(1024)
pack_scale_kernel(const float* __restrict__ src, int64_t n, float act_scale,
                  float* __restrict__ inv_scale, float* __restrict__ w_scale) {
    // one block of 1024 threads, 16-byte loads (every packed matrix is 256-byte aligned with n % 4 == 0)
    __shared__ float red[32];
    float m = 0.f;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int64_t i = threadIdx.x; i < (n >> 2); i += 1024) {
        const float4 v = __ldg(s4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int64_t i = (n & ~(int64_t)3) + threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(src[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            float sw = 1.f;
            if (m > 0.f && m < INFINITY) sw = exp2f(floorf(log2f(32768.f / m)));
            *w_scale = sw;
            *inv_scale = 1.f / (sw * act_scale);
        }
    }
}
__global__ void pack_split_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                  __half* __restrict__ lo, int64_t n, const float* __restrict__ w_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        __half h, l;
        half_split(src[i] * *w_scale, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

}  // namespace tip
