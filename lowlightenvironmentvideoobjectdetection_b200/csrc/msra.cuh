// msra.cuh -- shared declarations of the most-similar-RoI-align kernels (tafa.cu, msra_gemm.cu).
#pragma once
#include "common.cuh"

namespace vod {

constexpr int kMsraMaxK = 4;      // num_most_similar_points <= 4 (reference default 2)
constexpr int kMsraCand = 16;     // candidates per (row, frame) kept by the tensor-core pass (4 column groups x top-4)
// A candidate is re-scored in fp32 when its bf16-GEMM similarity is within this margin of the 2nd best one.
// bf16 operand rounding gives an error of ~1e-4 (sigma) on unit vectors, the 20-bit key truncation < 3.5e-4.
constexpr float kMsraMargin = 2.5e-3f;

int msra_launch_scan(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm, float *out,
                     int *idx_out, float *val_out, int NP, int C, int T, int HW, int k, cudaStream_t st);
int msra_launch_rescore(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm,
                        const uint32_t *cand, int KC, float *out, int *idx_out, float *val_out, int NP, int C, int T,
                        int HW, int k, cudaStream_t st);
// tensor-core candidate pass (msra_gemm.cu): bf16 unit rows -> cand [NP, T, kMsraCand]
bool msra_gemm_supported(int NP, int C, int T, int HW);
int msra_launch_gemm_topk(const void *roi_unit_bf16, const void *ref_unit_bf16, uint32_t *cand, int NP, int NP_pad, int C,
                          int T, int HW, cudaStream_t st);

}  // namespace vod
