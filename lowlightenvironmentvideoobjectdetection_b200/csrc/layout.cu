// layout.cu -- NCHW -> NHWC transposition fused with the per-pixel L2 norm and the unit-norm bf16
// copy that feeds the most-similar-location GEMM.  Replaces the permute/contiguous/norm passes of
// mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:127-140,159-161.
// HBM-bound: every input element is read once, every output written once; the [32 px][C] slab
// lives in shared memory (padded row => conflict-free in both access directions).
#include <cuda_bf16.h>

#include "common.cuh"

namespace vod {

constexpr int kTrPix = 32;
constexpr int kTrThreads = 256;

__global__ void __launch_bounds__(kTrThreads)
nchw_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, float *__restrict__ norm_out,
                    __nv_bfloat16 *__restrict__ unit, int C, int HW) {
    extern __shared__ float slab[];  // [kTrPix][C + 1]
    __shared__ float s_nrm[kTrPix];
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * kTrPix;
    const int np = min(kTrPix, HW - p0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = C + 1;
    const float *src = in + (size_t)b * C * HW + p0;
    // 8 channel planes per warp in flight before the first shared-memory store (one load per store left the warps waiting on a
    // single 128-byte request at a time: ncu put 38 % of the stalls on that STS)
#ifndef VOD_TR_BATCH
#define VOD_TR_BATCH 8
#endif
    constexpr int kWarps = kTrThreads / 32, kBatch = VOD_TR_BATCH;
    for (int c = warp; c < C; c += kWarps * kBatch) {
        float v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int cu = c + u * kWarps;
            v[u] = (lane < np && cu < C) ? __ldg(src + (size_t)cu * HW + lane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int cu = c + u * kWarps;
            if (cu < C) slab[lane * ld + cu] = v[u];
        }
    }
    __syncthreads();
    if (norm_out || unit) {
        for (int p = warp; p < np; p += kTrThreads / 32) {
            float s = 0.f;
            for (int c = lane; c < C; c += 32) { float v = slab[p * ld + c]; s = fmaf(v, v, s); }
            s = warp_sum(s);
            float nrm = sqrtf(s);
            if (lane == 0) {
                s_nrm[p] = nrm;
                if (norm_out) norm_out[(size_t)b * HW + p0 + p] = nrm;
            }
        }
        __syncthreads();
    }
    for (int p = warp; p < np; p += kTrThreads / 32) {
        float *dst = out ? out + ((size_t)b * HW + p0 + p) * C : nullptr;
        if (dst)
            for (int c = lane; c < C; c += 32) dst[c] = slab[p * ld + c];
        if (unit) {
            const float rinv = 1.0f / s_nrm[p];  // x * (1/||x||): per-element IEEE division is slow on exact zeros
            __nv_bfloat16 *u = unit + ((size_t)b * HW + p0 + p) * C;
            if ((C & 1) == 0) {
                for (int c = 2 * lane; c < C; c += 64) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(slab[p * ld + c] * rinv, slab[p * ld + c + 1] * rinv);
                    *reinterpret_cast<__nv_bfloat162 *>(u + c) = h;
                }
            } else {
                for (int c = lane; c < C; c += 32) u[c] = __float2bfloat16_rn(slab[p * ld + c] * rinv);
            }
        }
    }
}

// one warp per row
__global__ void __launch_bounds__(256)
rows_l2norm_kernel(const float *__restrict__ rows, float *__restrict__ norm_out,
                   __nv_bfloat16 *__restrict__ unit, int R, int C) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float *src = rows + (size_t)r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { float v = __ldg(src + c); s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    if (lane == 0 && norm_out) norm_out[r] = nrm;
    if (unit) {
        __nv_bfloat16 *u = unit + (size_t)r * C;
        const float rinv = 1.0f / nrm;
        for (int c = lane; c < C; c += 32) u[c] = __float2bfloat16_rn(__ldg(src + c) * rinv);
    }
}

}  // namespace vod

using namespace vod;

extern "C" int vod_nchw_to_nhwc(const float *in_nchw, float *out_nhwc, float *norm_out, void *out_unit_bf16,
                                int B, int C, int H, int W, vod_stream_t stream) {
    if (B == 0) return VOD_OK;
    VOD_REQUIRE(in_nchw && (out_nhwc || norm_out || out_unit_bf16), "vod_nchw_to_nhwc: null pointer");
    VOD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "vod_nchw_to_nhwc: bad dims");
    size_t smem = sizeof(float) * kTrPix * (C + 1);
    VOD_REQUIRE(smem <= 200 * 1024, "vod_nchw_to_nhwc: C=%d too large", C);
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(nchw_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ceil_div(H * W, kTrPix), B);
    nchw_to_nhwc_kernel<<<grid, kTrThreads, smem, as_stream(stream)>>>(
        in_nchw, out_nhwc, norm_out, reinterpret_cast<__nv_bfloat16 *>(out_unit_bf16), C, H * W); note_launch();
    return check_launch("vod_nchw_to_nhwc");
}

extern "C" int vod_rows_l2norm(const float *rows, float *norm_out, void *out_unit_bf16, int R, int C,
                               vod_stream_t stream) {
    if (R == 0) return VOD_OK;
    VOD_REQUIRE(rows && (norm_out || out_unit_bf16) && R > 0 && C > 0, "vod_rows_l2norm: bad args");
    rows_l2norm_kernel<<<ceil_div(R, 8), 256, 0, as_stream(stream)>>>(
        rows, norm_out, reinterpret_cast<__nv_bfloat16 *>(out_unit_bf16), R, C); note_launch();
    return check_launch("vod_rows_l2norm");
}
