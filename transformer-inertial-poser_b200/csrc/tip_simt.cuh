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
// one (M, kin_pad) matrix (zero padded) in fp32 or TF32 hi/lo planes.
__global__ void condition_kernel(const float* __restrict__ x_imu, const float* __restrict__ x_s,
                                 const float* __restrict__ keep_mask, float past_scale,
                                 float* __restrict__ out, float* __restrict__ out_lo,
                                 int M, int n_imu, int size_s, int kin_pad,
                                 float p_in, float p_past, uint64_t seed) {
    const int64_t total = (int64_t)M * kin_pad;
    const float inv_in = p_in > 0.f ? 1.f / (1.f - p_in) : 1.f;
    const float inv_past = p_past > 0.f ? (p_past < 1.f ? 1.f / (1.f - p_past) : 0.f) : 1.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kin_pad);
        const int c = (int)(i - (int64_t)r * kin_pad);
        float v = 0.f;
        if (c < n_imu) {
            v = x_imu[(int64_t)r * n_imu + c];
            if (p_in > 0.f) v *= dropout_factor(p_in, inv_in, seed ^ 0x1111, i);
        } else if (c < n_imu + size_s) {
            const int cs = c - n_imu;
            v = x_s[(int64_t)r * size_s + cs];
            if (v != v) v = 0.f;                               // :65  x_s[isnan] = 0
            if (cs >= 108 && cs < 111) v = 0.f;                // :75  root velocity removed
            if (keep_mask != nullptr) v *= keep_mask[(int64_t)r * size_s + cs] * past_scale;
            else if (p_past > 0.f) v *= dropout_factor(p_past, inv_past, seed ^ 0x2222, i);   // :77
        }
        if (out_lo != nullptr) {
            float hi, lo;
            tf32_split(v, hi, lo);
            out[i] = hi;
            out_lo[i] = lo;
        } else {
            out[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).  A and W are K-contiguous (nn.Linear layout).
struct Epi {
    const float* bias;      // [N]
    const float* resid;     // LN variant: residual rows [M][ldr] (fp32, or hi plane when resid_lo)
    const float* resid_lo;  // optional lo plane of the residual
    int ldr;
    const float* gamma;     // LN affine
    const float* beta;
    float* out;             // fp32 output, or hi plane when out_lo != nullptr
    float* out_lo;
    int ldc;
    int relu;
    float drop_p;           // dropout on the GEMM output (after ReLU; before the residual add)
    uint64_t seed;
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
    const float inv_keep = ep.drop_p > 0.f ? 1.f / (1.f - ep.drop_p) : 1.f;
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
            if (ep.drop_p > 0.f) x *= dropout_factor(ep.drop_p, inv_keep, ep.seed, (uint64_t)row * N + col);
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
                    if (ep.resid_lo) r += ep.resid_lo[(size_t)row * ep.ldr + col];
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
                    if (ep.out_lo) {
                        float4 hi, lo;
                        tf32_split(v[g * 4 + 0], hi.x, lo.x);
                        tf32_split(v[g * 4 + 1], hi.y, lo.y);
                        tf32_split(v[g * 4 + 2], hi.z, lo.z);
                        tf32_split(v[g * 4 + 3], hi.w, lo.w);
                        *reinterpret_cast<float4*>(o) = hi;
                        *reinterpret_cast<float4*>(ep.out_lo + (size_t)row * ep.ldc + col) = lo;
                    } else {
                        *reinterpret_cast<float4*>(o) =
                            make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < N) {
                            if (ep.out_lo) {
                                float hi, lo;
                                tf32_split(v[g * 4 + j], hi, lo);
                                o[j] = hi;
                                ep.out_lo[(size_t)row * ep.ldc + col + j] = lo;
                            } else {
                                o[j] = v[g * 4 + j];
                            }
                        }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Causal multi-head self-attention core (reference :85-91 -> nn.MultiheadAttention):
//   P = softmax(q k^T + causal mask) [dropout], o = P v, per (window b, head h); 1/sqrt(d) is folded
//   into W_q at pack time.  One thread owns the query-row pair (p, L-1-p) so every lane does the
//   same ~L+1 keys of causal work; KS lanes split the keys of a pair and merge (max, sum, o) with
//   warp shuffles.  K and V of the block's heads are staged once in shared memory and read as
//   128-bit broadcasts.  The contraction is tiny (d=16, L<=40): FFMA, no tensor cores.
template <int KS, int HPB>
__global__ void __launch_bounds__(HPB * 20 * KS)
attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ out_lo,
                 int L, float drop_p, uint64_t seed) {
    constexpr int NJ = (MAXL + KS - 1) / KS;
    __shared__ __align__(16) float Ks[HPB][MAXL][HD];
    __shared__ __align__(16) float Vs[HPB][MAXL][HD];
    const int b = blockIdx.x;
    const int h0 = blockIdx.y * HPB;
    const int tid = threadIdx.x;
    const float* base = qkv + (size_t)b * L * (3 * E);

    // stage K, V: HPB heads x L rows x 16 floats (4 float4 per row)
    for (int i = tid; i < HPB * L * 4; i += blockDim.x) {
        const int hl = i / (L * 4);
        const int rem = i - hl * (L * 4);
        const int row = rem >> 2, q4 = rem & 3;
        const float* src = base + (size_t)row * (3 * E) + (h0 + hl) * HD + q4 * 4;
        *reinterpret_cast<float4*>(&Ks[hl][row][q4 * 4]) = __ldg(reinterpret_cast<const float4*>(src + E));
        *reinterpret_cast<float4*>(&Vs[hl][row][q4 * 4]) = __ldg(reinterpret_cast<const float4*>(src + 2 * E));
    }
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
    {
        const float* qpa = base + (size_t)ra * (3 * E) + (h0 + hl) * HD;
        const float* qpb = base + (size_t)rb * (3 * E) + (h0 + hl) * HD;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 va = __ldg(reinterpret_cast<const float4*>(qpa) + i);
            const float4 vb = __ldg(reinterpret_cast<const float4*>(qpb) + i);
            qa[i * 4 + 0] = va.x; qa[i * 4 + 1] = va.y; qa[i * 4 + 2] = va.z; qa[i * 4 + 3] = va.w;
            qb[i * 4 + 0] = vb.x; qb[i * 4 + 1] = vb.y; qb[i * 4 + 2] = vb.z; qb[i * 4 + 3] = vb.w;
        }
    }
    float sa[NJ], sb[NJ];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int j = jj * KS + ks;
        sa[jj] = -INFINITY;
        sb[jj] = -INFINITY;
        if (active && j <= rb) {
            float da = 0.f, db = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 kv = *reinterpret_cast<const float4*>(&Ks[hl][j][i * 4]);
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
    float oa[HD], ob[HD], la = 0.f, lb = 0.f;
#pragma unroll
    for (int i = 0; i < HD; ++i) { oa[i] = 0.f; ob[i] = 0.f; }
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int j = jj * KS + ks;
        if (active && j <= rb) {
            float pa = (j <= ra) ? expf(sa[jj] - ma) : 0.f;
            float pb = expf(sb[jj] - mb);
            la += pa;
            lb += pb;
            if (drop_p > 0.f) {   // attention-probability dropout (train mode only)
                const uint64_t ida = (((uint64_t)b * NH + h0 + hl) * MAXL + ra) * MAXL + j;
                const uint64_t idb = (((uint64_t)b * NH + h0 + hl) * MAXL + rb) * MAXL + j;
                pa *= dropout_factor(drop_p, inv_keep, seed, ida);
                pb *= dropout_factor(drop_p, inv_keep, seed, idb);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 vv = *reinterpret_cast<const float4*>(&Vs[hl][j][i * 4]);
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
    if (active && ks == 0) {
        const float ia = 1.f / la, ib = 1.f / lb;
        float* da = out + ((size_t)b * L + ra) * E + (h0 + hl) * HD;
        float* db = out + ((size_t)b * L + rb) * E + (h0 + hl) * HD;
        float* dal = out_lo ? out_lo + ((size_t)b * L + ra) * E + (h0 + hl) * HD : nullptr;
        float* dbl = out_lo ? out_lo + ((size_t)b * L + rb) * E + (h0 + hl) * HD : nullptr;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 va = make_float4(oa[i * 4 + 0] * ia, oa[i * 4 + 1] * ia, oa[i * 4 + 2] * ia, oa[i * 4 + 3] * ia);
            float4 vb = make_float4(ob[i * 4 + 0] * ib, ob[i * 4 + 1] * ib, ob[i * 4 + 2] * ib, ob[i * 4 + 3] * ib);
            if (out_lo) {
                float4 ha, lla, hb, llb;
                tf32_split(va.x, ha.x, lla.x); tf32_split(va.y, ha.y, lla.y);
                tf32_split(va.z, ha.z, lla.z); tf32_split(va.w, ha.w, lla.w);
                tf32_split(vb.x, hb.x, llb.x); tf32_split(vb.y, hb.y, llb.y);
                tf32_split(vb.z, hb.z, llb.z); tf32_split(vb.w, hb.w, llb.w);
                reinterpret_cast<float4*>(da)[i] = ha;
                reinterpret_cast<float4*>(dal)[i] = lla;
                if (two) { reinterpret_cast<float4*>(db)[i] = hb; reinterpret_cast<float4*>(dbl)[i] = llb; }
            } else {
                reinterpret_cast<float4*>(da)[i] = va;
                if (two) reinterpret_cast<float4*>(db)[i] = vb;
            }
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
                    float hi, lo;
                    tf32_split(v, hi, lo);
                    hs[o] = hi;
                    hs_lo[o] = lo;
                } else {
                    hs[o] = v;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Per-frame window update of the streaming path (replaces the runner's per-frame re-assembly,
// real_time_runner_minimal.py:131-147): when a stream's window is full, slide it up by one row
// (in place: every thread first reads its elements, the CTA synchronises, then writes them one row
// earlier -- coalesced both ways), then append the new row.  One CTA per stream and window.
__global__ void __launch_bounds__(256)
window_push_kernel(float* __restrict__ win, const float* __restrict__ new_row, int width, int len) {
    // win: (S, MAXL, width); len = rows currently held (same for every stream)
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

// gather compacted (S, L, width) windows out of the (S, MAXL, width) storage when L < MAXL
__global__ void window_compact_kernel(const float* __restrict__ win, float* __restrict__ out,
                                      int S, int L, int width) {
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
__global__ void pack_split_kernel(const float* __restrict__ src, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float h, l;
        tf32_split(src[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

}  // namespace tip
