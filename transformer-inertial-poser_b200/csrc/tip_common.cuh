// Shared definitions for the TIP hot-path kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <utility>

#include "../../include/tip_b200.h"

namespace tip {

// The one architecture the reference ships (offline_testing_simple.py:87-95).
constexpr int E = 256;        // tf_in_dim
constexpr int NH = 16;        // n_heads
constexpr int HD = 16;        // head dim
constexpr int F = 1024;       // tf_hid_size
constexpr int R = 512;        // rnn_hid_size
constexpr int MAXL = 40;      // runner's max_input_l (real_time_runner_minimal.py:131)
constexpr int MAX_LAYERS = 8;
constexpr int HEAD_NPAD = 144;  // size_s (<=144) padded to a legal UMMA N (multiple of 16)

// Offsets (in floats) into the packed weight blob.  Every matrix exists as an fp32 plane (FFMA
// engine) and as an FP16 hi/lo pair of the power-of-two pre-scaled matrix, hi = fp16(s*w),
// lo = fp16(s*w - hi), for the tcgen05 3xFP16 GEMMs (the *_hi/_lo offsets are float offsets of
// __half planes; `scales` holds 1/(s_w * s_a) per matrix, read by the GEMM epilogue).
struct LayerOff {
    size_t wqkv, bqkv, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2;
    size_t wqkv_hi, wqkv_lo, wo_hi, wo_lo, w1_hi, w1_lo, w2_hi, w2_lo;
    size_t wqkvr_hi, wqkvr_lo, bqkvr;      // in_proj rows / bias re-ordered per head group (fused QKV + attention kernel)
};
// index of a matrix in the `scales` table
constexpr int SC_IN = 0, SC_LAYER0 = 1 /* + 4*layer + {qkv,o,1,2} */, SC_IH = 1 + 4 * MAX_LAYERS, SC_HEAD = SC_IH + 1,
              SC_HH = SC_HEAD + 1, SC_COUNT = SC_HH + 1;
// activations are stored as FP16 hi/lo planes of ACT_SCALE * x (|x| < 4096 representable; the
// model's LayerNorm / ReLU / tanh outputs stay below ~20), the raw model input with scale 1.
constexpr float ACT_SCALE = 16.f;
struct PackOff {
    size_t win, bin, win_hi, win_lo;          // [E][kin_pad]
    LayerOff layer[MAX_LAYERS];
    size_t wih, brnn, whh, whh_t, wih_hi, wih_lo;   // whh [R][R] (out,in);  whh_t [R k][R n]
    size_t whh_hi, whh_lo;                          // FP16 planes of s_w * W_hh (tensor-core recurrence)
    size_t wl, bl, wl_hi, wl_lo;              // head [HEAD_NPAD][khead] zero padded rows
    size_t scales;                            // [SC_COUNT] accumulator un-scale factors
    size_t total;
};

struct Dims {
    int n_imu;      // input_size_imu (+18 with acc sum)
    int size_s;
    int d_in;       // n_imu + size_s
    int kin_pad;    // d_in rounded up to 64 (one 128-byte TMA/UMMA k-block of fp16)
    int layers;
    int with_rnn;
    int khead;      // R or E
};

// Counter-based RNG of the reference's dropout sites (statistically equivalent, not bit-equal to torch's Philox
// stream -- DESIGN.md "as-shipped mode").  One 64-bit splitmix hash yields FOUR 16-bit uniforms: element `idx` of a
// site uses lane (idx & 3) of hash(site_seed, idx >> 2) and is DROPPED when its uniform is below
// thr = round(p * 65536); kept elements are scaled by 1/(1-p) like nn.Dropout.  The site seed is
// base_seed + site constant, the base seed lives in DEVICE memory (tip_model::d_seed, written by a one-thread
// kernel before every stochastic forward) so a captured CUDA graph draws fresh masks on every replay.
// oracle/tip_oracle.py restates exactly this generator, so stochastic forwards are parity-tested mask for mask.
__host__ __device__ __forceinline__ uint64_t hash_u64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
    return p <= 0.f ? 0u : (p >= 1.f ? 65536u : (uint32_t)(p * 65536.f + 0.5f));
}
__host__ __device__ __forceinline__ float drop_inv_keep(float p) { return p > 0.f ? (p < 1.f ? 1.f / (1.f - p) : 0.f) : 1.f; }
// factors of the four elements 4*group .. 4*group + 3
__device__ __forceinline__ float4 dropout_factor4(uint64_t seed, uint64_t group, uint32_t thr, float inv_keep) {
    const uint64_t h = hash_u64(seed, group);
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    return make_float4((lo & 0xFFFFu) < thr ? 0.f : inv_keep, (lo >> 16) < thr ? 0.f : inv_keep,
                       (hi & 0xFFFFu) < thr ? 0.f : inv_keep, (hi >> 16) < thr ? 0.f : inv_keep);
}
// single element (scalar sites): 0 with probability p, 1/(1-p) otherwise
__device__ __forceinline__ float dropout_factor(float p, float inv_keep, uint64_t seed, uint64_t idx) {
    const uint64_t h = hash_u64(seed, idx >> 2);
    const uint32_t u = (uint32_t)(h >> (16 * (idx & 3))) & 0xFFFFu;
    return u < drop_threshold(p) ? 0.f : inv_keep;
}
// Element index of the attention-probability dropout, P[bh = b * 16 + h][query][key]: the four probabilities one lane
// of the tensor-core attention kernel holds per key tile -- queries r and r + 8 of a 16-row tile x two adjacent keys --
// form ONE hash group (one hash per four probabilities): group = ((bh * 3 + query / 16) * 8 + query % 8) * 20 + key / 2,
// lane = 2 * ((query % 16) / 8) + key % 2.
__host__ __device__ __forceinline__ uint64_t attn_drop_index(uint64_t bh, int query, int key) {
    const uint64_t grp = ((bh * 3 + (uint64_t)(query >> 4)) * 8 + (uint64_t)(query & 7)) * 20 + (uint64_t)(key >> 1);
    return 4 * grp + (uint64_t)(2 * ((query >> 3) & 1) + (key & 1));
}
// site seed = *base (device memory; null = 0) + per-site offset
__device__ __forceinline__ uint64_t site_seed(const uint64_t* base, uint64_t off) { return (base ? *base : 0ull) + off; }
// per-site offsets (layer l = 0..): also restated in oracle/tip_oracle.py
constexpr uint64_t SEED_IN = 0x1111ull, SEED_PAST = 0x2222ull;
__host__ __device__ constexpr uint64_t seed_attn(int l) { return 101ull * (uint64_t)(l + 1); }
__host__ __device__ constexpr uint64_t seed_out(int l) { return 211ull * (uint64_t)(l + 1); }
__host__ __device__ constexpr uint64_t seed_ff1(int l) { return 307ull * (uint64_t)(l + 1); }
__host__ __device__ constexpr uint64_t seed_ff2(int l) { return 401ull * (uint64_t)(l + 1); }
constexpr uint64_t SEED_CHUNK = 0x632BE59BD9B4E019ull;     // added per 1024-window chunk of a large batch

// Error-compensated FP16 split of an (already scaled) fp32 value: v ~= hi + lo to ~22 bits.
// Both conversions saturate: |v| > 65504 clamps to +-65504 (hi) and the residual to +-65504 (lo), so an
// out-of-range finite input cannot inject inf/NaN (it clamps near +-131008; DESIGN.md "range contract").
__device__ __forceinline__ void half_split(float v, __half& hi, __half& lo) {
    unsigned short h, l;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    hi = __ushort_as_half(h);
    const float r = v - __half2float(hi);
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(r));
    lo = __ushort_as_half(l);
}
// four consecutive values -> 8-byte stores into the hi / lo planes
__device__ __forceinline__ void half_split_store4(__half* hi_p, __half* lo_p, float4 v) {
    __half h[4], l[4];
    half_split(v.x, h[0], l[0]); half_split(v.y, h[1], l[1]);
    half_split(v.z, h[2], l[2]); half_split(v.w, h[3], l[3]);
    uint2 uh, ul;
    uh.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    uh.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    ul.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    ul.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(hi_p) = uh;
    *reinterpret_cast<uint2*>(lo_p) = ul;
}
// four consecutive values of hi + lo (un-scaled by the caller)
__device__ __forceinline__ float4 half_pair_load4(const __half* hi_p, const __half* lo_p) {
    const uint2 uh = __ldg(reinterpret_cast<const uint2*>(hi_p));
    const uint2 ul = __ldg(reinterpret_cast<const uint2*>(lo_p));
    const __half2 h0 = *reinterpret_cast<const __half2*>(&uh.x), h1 = *reinterpret_cast<const __half2*>(&uh.y);
    const __half2 l0 = *reinterpret_cast<const __half2*>(&ul.x), l1 = *reinterpret_cast<const __half2*>(&ul.y);
    const float2 a = __half22float2(h0), b = __half22float2(h1), c = __half22float2(l0), d = __half22float2(l1);
    return make_float4(a.x + c.x, a.y + c.y, b.x + d.x, b.y + d.y);
}

// Programmatic dependent launch (PDL): every kernel of the path is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, calls griddep_wait() before it touches anything an earlier
// kernel wrote (or still reads) and griddep_launch() right after, so that the NEXT kernel's launch latency and
// prologue (barrier init, TMEM allocation, tensor-map prefetch, resident-weight loads) overlap this kernel's
// execution instead of following it.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Measured on B200: B = 1 forward 327 -> 309 us with PDL (33 tiny kernels, the fixed cost per kernel is what it
// hides); B = 256: 553 -> 565 us (the kernels are long, and early-resident dependents only add contention).  So it
// is on for small batches only: TIP_PDL = 0 never, 1 (default) when the forward has <= 1024 rows, 2 always.
inline int pdl_mode() {
    static const int v = getenv("TIP_PDL") ? atoi(getenv("TIP_PDL")) : 1;
    return v;
}
inline int& pdl_rows() { static thread_local int rows = 0; return rows; }     // rows of the forward being launched (set by the host path)
inline int& pdl_kind() { static thread_local int kind = 8; return kind; }     // 1 GEMM, 2 attention, 4 recurrence, 8 everything else
inline int pdl_big_mask() {                                      // kernel kinds that use PDL in large forwards (experiment)
    static const int v = getenv("TIP_PDL_BIG_MASK") ? atoi(getenv("TIP_PDL_BIG_MASK")) : 0;
    return v;
}
inline bool pdl_enabled() {
    if (pdl_mode() == 2) return true;
    if (pdl_mode() != 1) return false;
    return pdl_rows() <= 1024 || (pdl_big_mask() & pdl_kind()) != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tip

#define TIP_CUDA_TRY(m, expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            (m)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));            \
            return TIP_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)
