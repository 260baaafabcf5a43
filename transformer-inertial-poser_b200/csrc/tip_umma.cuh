// tcgen05 3xTF32 GEMM engine (placeholder interface; filled in by tip_umma.cuh proper).
#pragma once
#include "tip_common.cuh"
#include "tip_simt.cuh"

namespace tip {

enum UmmaGemmId { UG_IN = 0, UG_QKV, UG_OUT, UG_FF1, UG_FF2, UG_IH, UG_HEAD_R, UG_HEAD_E, UG_COUNT };

constexpr bool UMMA_AVAILABLE = false;
struct UmmaMaps { int dummy = 0; };

inline int umma_build_maps(UmmaMaps&, const float*, const PackOff&, const Dims&, float*, size_t, float*, float*,
                           float*, size_t, float*, size_t, float*, size_t, int, std::string& err) {
    err = "tcgen05 engine not built";
    return TIP_ERR_INVALID_ARG;
}
inline void umma_gemm(UmmaMaps&, int, int, int, int, int, const Epi&, bool, cudaStream_t) {}

}  // namespace tip
