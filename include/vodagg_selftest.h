/*
 * vodagg_selftest.h -- C ABI of libvodagg_selftest.so: test-only kernels that validate the tcgen05 / TMA building blocks
 * (csrc/tc.cuh, csrc/tmap.cu) in isolation.  Built by the same build.py next to libvodagg.so; NOT part of the product
 * library and never on the path of any operator.  Same conventions as vodagg.h (device pointers, stream-ordered, 0 = ok;
 * vod_last_error() of THIS library carries the message).
 */
#ifndef VODAGG_SELFTEST_H_
#define VODAGG_SELFTEST_H_

#include "vodagg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Plain tcgen05 GEMM used by tests/test_gpu_tc.py to validate descriptors / pipeline:
 * D[M,N] (fp32) = A[M,K] * B[N,K]^T, A/B row-major (K contiguous); dtype VOD_DTYPE_BF16, VOD_DTYPE_F32 (tf32 MMA) or 2
 * (bf16 with the A tile staged in TMEM, TS-form MMA). */
int vod_test_gemm_nt(const void *a, const void *b, float *d, int M, int N, int K, int dtype,
                     vod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VODAGG_SELFTEST_H_ */
