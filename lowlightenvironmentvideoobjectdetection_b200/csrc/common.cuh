// common.cuh -- shared helpers for libvodagg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vodagg.h"

namespace vod {

// thread-local last-error message (C ABI never throws)
char *last_error_buf();
int fail(int code, const char *fmt, ...);

// Every kernel launch site calls note_launch(); vod_kernel_launch_count() exposes the total so that
// the benchmark can report how many of OUR kernels ran inside its timed region.
void note_launch(int n = 1);

static inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(VOD_E_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return VOD_OK;
}

static inline cudaStream_t as_stream(vod_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// SM count of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200): grids of the
// persistent kernels are sized from it, never from a constant.
int num_sms();

// 128-bit streaming loads/stores (read-only path, do not pollute L1 for write-once data)
__device__ __forceinline__ float4 ldg_f4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void stg_cs_f4(float *p, float4 v) {
    __stcs(reinterpret_cast<float4 *>(p), v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace vod

#define VOD_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return vod::fail(VOD_E_BADARG, __VA_ARGS__); \
    } while (0)
