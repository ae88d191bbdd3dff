// selsa.cuh -- internal interface between selsa.cu (C ABI + SIMT kernel) and selsa_tc.cu (tcgen05 kernel).
#pragma once
#include "common.cuh"

namespace vod {

bool selsa_tc_supported(int N, int M, int heads, int d, int dtype, const void *q, const void *k, const void *v,
                        int v_layout, int ldv);
size_t selsa_tc_workspace_bytes(int N, int M, int heads, int d);
int selsa_tc_launch(const void *q, const void *k, const void *v, float *out, int N, int M, int heads, int d, float scale,
                    int dtype, int v_layout, int ldv, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace vod
