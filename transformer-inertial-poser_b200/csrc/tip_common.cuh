// Shared definitions for the TIP hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

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

// Offsets (in floats) into the packed weight blob.  Every matrix exists as an fp32 plane and as a
// TF32 hi/lo pair (hi = rna_tf32(w), lo = rna_tf32(w - hi)) for the tcgen05 3xTF32 GEMMs.
struct LayerOff {
    size_t wqkv, bqkv, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2;
    size_t wqkv_hi, wqkv_lo, wo_hi, wo_lo, w1_hi, w1_lo, w2_hi, w2_lo;
};
struct PackOff {
    size_t win, bin, win_hi, win_lo;          // [E][kin_pad]
    LayerOff layer[MAX_LAYERS];
    size_t wih, brnn, whh, whh_t, wih_hi, wih_lo;   // whh [R][R] (out,in);  whh_t [R k][R n]
    size_t wl, bl, wl_hi, wl_lo;              // head [HEAD_NPAD][khead] zero padded rows
    size_t total;
};

struct Dims {
    int n_imu;      // input_size_imu (+18 with acc sum)
    int size_s;
    int d_in;       // n_imu + size_s
    int kin_pad;    // d_in rounded up to 32 (one 128-byte TMA/UMMA k-block of fp32)
    int layers;
    int with_rnn;
    int khead;      // R or E
};

// counter-based RNG for the reference's dropout sites (statistically equivalent, not bit-equal to
// torch's Philox stream -- see DESIGN.md "stochastic mode").
__host__ __device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
// returns the multiplicative factor of nn.Dropout(p): 0 with prob p, 1/(1-p) otherwise
__device__ __forceinline__ float dropout_factor(float p, float inv_keep, uint64_t seed, uint64_t idx) {
    const uint32_t r = hash_u32(seed, idx);
    return ((float)r * 2.3283064365386963e-10f < p) ? 0.f : inv_keep;
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) {
    hi = tf32_rna(v);
    lo = tf32_rna(v - hi);
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
