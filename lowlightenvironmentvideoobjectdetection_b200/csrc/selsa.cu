// selsa.cu -- (3) SELSA similarity-softmax-weighted-sum: C-ABI entry + the generic fp32 SIMT kernel.
// Restates the bmm/softmax/bmm core of SelsaAggregator.forward,
// mmtracking/mmtrack/models/aggregators/selsa_aggregator.py:51-70, without materialising the
// [heads, N, M] weight tensor.  The tensor-core (tcgen05/TMEM) kernel for d == 64 lives in
// selsa_tc.cu; this file's SIMT kernel covers every other head size (the reference's own unit test
// uses d = 4) and doubles as an on-device cross-check of the tensor-core path.
#include <cuda_bf16.h>

#include "common.cuh"
#include "selsa.cuh"

namespace vod {

template <typename T> __device__ __forceinline__ float ld_as_float(const T *p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

constexpr int kSimtWarps = 4;

// warp = (query row n, head h); lane l owns reference rows l, l+32, ... with a private online softmax.
template <typename T, int DMAX>
__global__ void __launch_bounds__(kSimtWarps * 32)
selsa_simt_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v, float *__restrict__ out,
                  int N, int M, int heads, int d, float scale, int v_layout, int ldv) {
    const int lane = threadIdx.x & 31;
    const long task = (long)blockIdx.x * kSimtWarps + (threadIdx.x >> 5);
    if (task >= (long)N * heads) return;
    const int n = (int)(task / heads), h = (int)(task % heads);
    const int D = heads * d;
    float qr[DMAX], o[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        qr[i] = i < d ? ld_as_float(q + (size_t)n * D + h * d + i) * scale : 0.f;
        o[i] = 0.f;
    }
    float mx = -INFINITY, sum = 0.f;
    for (int m = lane; m < M; m += 32) {
        const T *kr = k + (size_t)m * D + h * d;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) if (i < d) s = fmaf(qr[i], ld_as_float(kr + i), s);
        const float nm = fmaxf(mx, s);
        const float corr = expf(mx - nm), p = expf(s - nm);
        sum = sum * corr + p;
        // V[m, h*d + i]: row-major, or transposed V^T[h*d + i, m] with leading dimension ldv
        const T *vr = v_layout == 0 ? v + (size_t)m * D + h * d : v + (size_t)h * d * ldv + m;
        const size_t vs = v_layout == 0 ? 1 : (size_t)ldv;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) if (i < d) o[i] = fmaf(p, ld_as_float(vr + i * vs), o[i] * corr);
        mx = nm;
    }
    const float gm = warp_max(mx);
    const float c = mx == -INFINITY ? 0.f : expf(mx - gm);
    const float gs = warp_sum(sum * c);
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        if (i < d) {
            float t = warp_sum(o[i] * c);
            if (lane == 0) out[(size_t)n * D + h * d + i] = t / gs;
        }
    }
}

template <typename T>
static int launch_simt(const T *q, const T *k, const T *v, float *out, int N, int M, int heads, int d, float scale,
                       int v_layout, int ldv, cudaStream_t st) {
    const long tasks = (long)N * heads;
    const unsigned grid = (unsigned)ceil_div(tasks, (long)kSimtWarps);
    if (d > 128) return fail(VOD_E_UNSUPPORTED, "vod_selsa_attn: head dim %d > 128 unsupported", d);
    if (d <= 16) selsa_simt_kernel<T, 16><<<grid, kSimtWarps * 32, 0, st>>>(q, k, v, out, N, M, heads, d, scale, v_layout, ldv);
    else if (d <= 64) selsa_simt_kernel<T, 64><<<grid, kSimtWarps * 32, 0, st>>>(q, k, v, out, N, M, heads, d, scale, v_layout, ldv);
    else selsa_simt_kernel<T, 128><<<grid, kSimtWarps * 32, 0, st>>>(q, k, v, out, N, M, heads, d, scale, v_layout, ldv);
    note_launch();
    return check_launch("vod_selsa_attn(simt)");
}


// The tail of one SelsaBBoxHead layer (selsa_bbox_head.py:56-58) in one pass instead of four elementwise launches:
//   x[r, c]  = relu(x[r, c] + y[r, c] + bias[c])   r < rows    (x = x + aggregator(x, ref_x); x = relu(x)  with the
//              aggregator's output bias -- fc.bias + fc.weight . ref_fc.bias, see SelsaAggregator.out_bias -- added here)
//   ref[i]   = relu(ref[i])                        i < ref_elems   (ref_x = relu(ref_x); nullable)
// 128-bit accesses; cols % 4 == 0 and 16-byte aligned pointers are checked by the launcher.
__global__ void __launch_bounds__(256) selsa_residual_relu_kernel(float4 *__restrict__ x, const float4 *__restrict__ y,
                                                                  const float4 *__restrict__ bias, long n4, int cols4,
                                                                  float4 *__restrict__ ref, long ref4) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4 + ref4; i += stride) {
        if (i < n4) {
            float4 a = x[i];
            const float4 b = y[i], c = __ldg(bias + (int)(i % cols4));
            a.x = fmaxf(a.x + b.x + c.x, 0.f); a.y = fmaxf(a.y + b.y + c.y, 0.f);
            a.z = fmaxf(a.z + b.z + c.z, 0.f); a.w = fmaxf(a.w + b.w + c.w, 0.f);
            x[i] = a;
        } else {
            float4 a = ref[i - n4];
            a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
            ref[i - n4] = a;
        }
    }
}

}  // namespace vod

using namespace vod;

extern "C" size_t vod_selsa_attn_workspace_bytes(int N, int M, int heads, int d) {
    return selsa_tc_workspace_bytes(N, M, heads, d);
}

extern "C" int vod_selsa_attn(const void *q, const void *k, const void *v, float *out, int N, int M, int heads, int d,
                              float scale, int dtype, int v_layout, int ldv, int impl, void *ws, size_t ws_bytes,
                              vod_stream_t stream) {
    if (N == 0) return VOD_OK;
    VOD_REQUIRE(q && k && v && out, "vod_selsa_attn: null pointer");
    VOD_REQUIRE(N > 0 && M > 0 && heads > 0 && d > 0, "vod_selsa_attn: bad dims N=%d M=%d heads=%d d=%d", N, M, heads, d);
    VOD_REQUIRE(dtype == VOD_DTYPE_F32 || dtype == VOD_DTYPE_BF16, "vod_selsa_attn: dtype");
    VOD_REQUIRE(impl >= 0 && impl <= 2, "vod_selsa_attn: impl");
    VOD_REQUIRE(v_layout == 0 || (v_layout == 1 && ldv >= M), "vod_selsa_attn: v_layout=%d ldv=%d", v_layout, ldv);
    cudaStream_t st = as_stream(stream);
    const bool tc_ok = selsa_tc_supported(N, M, heads, d, dtype, q, k, v, v_layout, ldv) && vod_device_is_sm100();
    if (impl == 2 && !tc_ok) return fail(VOD_E_UNSUPPORTED, "vod_selsa_attn: tcgen05 path needs d == 64, 16B-aligned rows, sm_100");
    if (impl == 2 || (impl == 0 && tc_ok)) return selsa_tc_launch(q, k, v, out, N, M, heads, d, scale, dtype, v_layout, ldv, ws, ws_bytes, st);
    if (dtype == VOD_DTYPE_F32)
        return launch_simt(reinterpret_cast<const float *>(q), reinterpret_cast<const float *>(k),
                           reinterpret_cast<const float *>(v), out, N, M, heads, d, scale, v_layout, ldv, st);
    return launch_simt(reinterpret_cast<const __nv_bfloat16 *>(q), reinterpret_cast<const __nv_bfloat16 *>(k),
                       reinterpret_cast<const __nv_bfloat16 *>(v), out, N, M, heads, d, scale, v_layout, ldv, st);
}

extern "C" int vod_selsa_residual_relu(float *x, const float *y, const float *bias, int rows, int cols, float *ref,
                                       long ref_elems, vod_stream_t stream) {
    VOD_REQUIRE(rows >= 0 && cols > 0 && ref_elems >= 0, "vod_selsa_residual_relu: bad dims rows=%d cols=%d", rows, cols);
    if ((long)rows * cols + ref_elems == 0) return VOD_OK;
    VOD_REQUIRE(rows == 0 || (x && y && bias), "vod_selsa_residual_relu: null pointer");
    VOD_REQUIRE(ref_elems == 0 || ref, "vod_selsa_residual_relu: null ref pointer");
    VOD_REQUIRE(cols % 4 == 0 && ref_elems % 4 == 0, "vod_selsa_residual_relu: cols and ref_elems must be multiples of 4");
    VOD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) |
                  reinterpret_cast<uintptr_t>(ref)) & 15) == 0, "vod_selsa_residual_relu: pointers must be 16-byte aligned");
    const long n4 = (long)rows * cols / 4, ref4 = ref_elems / 4;
    const long blocks = (n4 + ref4 + 255) / 256;
    const int grid = (int)(blocks < 8L * num_sms() ? blocks : 8L * num_sms());
    selsa_residual_relu_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<float4 *>(x), reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(bias), n4, cols / 4,
        reinterpret_cast<float4 *>(ref), ref4); note_launch();
    return check_launch("vod_selsa_residual_relu");
}
