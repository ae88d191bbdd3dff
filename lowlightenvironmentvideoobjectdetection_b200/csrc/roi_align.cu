// roi_align.cu -- (1) average RoIAlign over NHWC features, key frame + all reference frames in one
// launch.  Semantics: mmcv-full 1.2.x roi_align forward, pool_mode='avg' (SURVEY Appendix A.1), as
// constructed at mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55.
//
// Design (HBM/L1-bound gather, not a GEMM):
//   * features are NHWC, so one bilinear tap is C contiguous floats: every lane issues 128-bit loads
//     and a warp covers 128 channels (512 B) per tap.
//   * the bin average is separable: out[i][j] = sum_r sum_c Wy[i][r] * Wx[j][c] * f[r][c], where
//     Wy / Wx are the per-axis bilinear weights summed over the g sample points of the bin.  The
//     per-axis sparse weight lists (<= 2g entries, duplicates merged) are built once per RoI in
//     shared memory; a small RoI whose 14x14 sample points fall on a handful of pixels then loads
//     each (row, col) once per bin-row instead of 16 taps per bin.
//   * one CTA = one RoI x up to 512 channels; warp w owns bin-row w and each lane accumulates 4 x 128-bit
//     channel groups per tap, so the per-tap bookkeeping is paid once per 2 KB of features.  For the
//     reference's [K,C,ph,pw] layout the [C][P] tile is staged in shared memory and written back as one
//     contiguous, coalesced 128-bit stream; the [K,P,C] layout is written directly (no tile, no barrier).
#include "common.cuh"

namespace vod {

constexpr int kMaxBins1D = 16;   // ph, pw <= 16
constexpr int kMaxEntries = 32;  // 2 * g, g <= 16
constexpr int kRoiWarps = 7;

struct AxisTable {
    int cnt[kMaxBins1D];
    int idx[kMaxBins1D][kMaxEntries];
    float w[kMaxBins1D][kMaxEntries];
};

// Build the sparse weight list of one bin along one axis (start = roi start on this axis in
// feature px, bin = bin size, g samples, size = H or W).  Arithmetic order follows the reference:
//   y = start + i*bin + (iy + 0.5) * bin / g
__device__ void build_axis(AxisTable &tab, int i, float start, float bin, int g, int size, int stride) {
    int cnt = 0;
    const float inv_g = 1.0f / (float)g;
    for (int s = 0; s < g; ++s) {
        float y = __fadd_rn(__fadd_rn(start, __fmul_rn((float)i, bin)),
                            __fdiv_rn(__fmul_rn((float)s + 0.5f, bin), (float)g));
        if (y < -1.0f || y > (float)size) continue;  // sample contributes zero
        if (y <= 0.f) y = 0.f;
        int lo = (int)y, hi;
        if (lo >= size - 1) { hi = lo = size - 1; y = (float)lo; } else { hi = lo + 1; }
        float l = y - (float)lo, h = 1.0f - l;
        float wv[2] = {h * inv_g, l * inv_g};
        int iv[2] = {lo, hi};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (wv[q] == 0.f) continue;
            int e = 0;
            for (; e < cnt; ++e)
                if (tab.idx[i][e] == iv[q]) break;
            if (e == cnt) { tab.idx[i][e] = iv[q]; tab.w[i][e] = 0.f; ++cnt; }
            tab.w[i][e] += wv[q];
        }
    }
    for (int e = 0; e < cnt; ++e) tab.idx[i][e] *= stride;   // pixel index -> element offset
    tab.cnt[i] = cnt;
}

__device__ __forceinline__ void prefetch_l1(const float *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int VEC> struct Vec;
template <> struct Vec<4> {
    float4 v;
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void load(const float *p) { v = ldg_f4(p); }
    __device__ __forceinline__ void fma(float w, const Vec &o) {
        v.x = fmaf(w, o.v.x, v.x); v.y = fmaf(w, o.v.y, v.y);
        v.z = fmaf(w, o.v.z, v.z); v.w = fmaf(w, o.v.w, v.w);
    }
    __device__ __forceinline__ float get(int q) const { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }
    __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = v; }
    __device__ __forceinline__ void store_cs(float *p) const { stg_cs_f4(p, v); }
};
template <> struct Vec<1> {
    float v;
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ void load(const float *p) { v = __ldg(p); }
    __device__ __forceinline__ void fma(float w, const Vec &o) { v = fmaf(w, o.v, v); }
    __device__ __forceinline__ float get(int) const { return v; }
    __device__ __forceinline__ void store(float *p) const { *p = v; }
    __device__ __forceinline__ void store_cs(float *p) const { __stcs(p, v); }
};

// One CTA = one RoI x (NCH * 32 * VEC) channels (all 512 channels of the R-50-DC5 neck for VEC=4, NCH=4), so the
// per-RoI work (RoI decode, weight lists, loop control, address arithmetic) is paid once per tap for 2 KB of
// features instead of once per 512 B.  Warp w owns bin-row w.
// (Forcing more resident CTAs through the register budget was measured in round 2, 4500 RoIs: 5 CTAs/SM (56 registers, a few
// spills) 212 us, 6 (40 registers) 327 us, 8 (32) 572 us against 193 us for the compiler's own allocation.)
// TILE = the reference's [K,C,ph,pw] output through the shared-memory tile (100 KB per CTA: two CTAs per SM at most); the
// channels-last variant writes directly.  Compiling the two as separate instantiations with their own register budgets
// (4500 RoIs, round 2): tile variant for 1 / 2 / 3 CTAs per SM 328 / 315 / 365 us (427 us as one kernel with a run-time
// layout switch), channels-last variant for 2 / 3 / 4 CTAs per SM 172 / 172 / 179 us (193 us before).
template <int VEC, int NCH, bool TILE>
#ifndef VOD_ROI_MINB_TILE
#define VOD_ROI_MINB_TILE 2
#define VOD_ROI_MINB_FLAT 3
#endif
__global__ void __launch_bounds__(kRoiWarps * 32, TILE ? VOD_ROI_MINB_TILE : VOD_ROI_MINB_FLAT)
roi_align_kernel(const float *__restrict__ feat, const float *__restrict__ rois, float *__restrict__ out,
                 int B, int C, int H, int W, int K, int ph, int pw, float spatial_scale,
                 int sampling_ratio, int aligned, int out_layout) {
    constexpr int CH = 32 * VEC;        // channels per warp-wide load
    constexpr int CS = CH * NCH;        // channels per CTA
    extern __shared__ __align__(16) float tile[];  // [cs_eff][P] (layout 0 only)
    __shared__ AxisTable ty, tx;
    __shared__ int s_batch;

    const int k = blockIdx.x;
    const int c0 = blockIdx.y * CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = ph * pw;

    {
        const float *r = rois + 5 * (size_t)k;
        const float off = aligned ? 0.5f : 0.f;
        float rsw = __fsub_rn(__fmul_rn(r[1], spatial_scale), off);
        float rsh = __fsub_rn(__fmul_rn(r[2], spatial_scale), off);
        float rew = __fsub_rn(__fmul_rn(r[3], spatial_scale), off);
        float reh = __fsub_rn(__fmul_rn(r[4], spatial_scale), off);
        float rw = __fsub_rn(rew, rsw), rh = __fsub_rn(reh, rsh);
        if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
        float bh = __fdiv_rn(rh, (float)ph), bw = __fdiv_rn(rw, (float)pw);
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)ph));
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)pw));
        gh = min(max(gh, 0), kMaxEntries / 2);
        gw = min(max(gw, 0), kMaxEntries / 2);
        if (tid < ph) build_axis(ty, tid, rsh, bh, gh, H, W * C);
        else if (tid >= 32 && tid < 32 + pw) build_axis(tx, tid - 32, rsw, bw, gw, W, C);
        if (tid == 0) s_batch = min(max((int)r[0], 0), B - 1);
    }
    __syncthreads();

    const int cl = lane * VEC;   // channel within a 32*VEC chunk
    const float *fb = feat + (size_t)s_batch * H * W * C + c0 + cl;
    asm volatile("" : "+l"(fb));  // keep the base pointer materialised (ptxas otherwise re-derives it per load)
    bool on[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) on[c] = c0 + c * CH + cl < C;   // C % VEC == 0 guaranteed by the launcher

    // (Batching the tap loads -- all loads of 4 taps issued before the first FMA -- was measured in round 2: 112 registers, two
    // CTAs per SM, 234 us instead of 193 us for 4500 RoIs.  The kernel is bound by L2 -> SM traffic, ~1.45 GB per launch at an
    // L1 hit rate of 46 %, not by bytes in flight: resident warps matter more than loads per warp.)
    for (int i = warp; i < ph; i += kRoiWarps) {
        const int ny = ty.cnt[i];
        for (int j = 0; j < pw; ++j) {
            Vec<VEC> acc[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[c].zero();
            const int nx = tx.cnt[j];
            for (int e = 0; e < ny; ++e) {
                const float wy = ty.w[i][e];
                const float *rowp = fb + ty.idx[i][e];
                for (int q = 0; q < nx; ++q) {
                    const float *p = rowp + tx.idx[j][q];
                    const float w = tx.w[j][q] * wy;
                    Vec<VEC> v[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (on[c]) v[c].load(p + c * CH); else v[c].zero();
#pragma unroll
                    for (int c = 0; c < NCH; ++c) acc[c].fma(w, v[c]);
                }
            }
            const int bin = i * pw + j;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (TILE) {
                    if (on[c]) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) tile[(c * CH + cl + q) * P + bin] = acc[c].get(q);
                    }
                } else if (on[c]) {
                    acc[c].store_cs(out + ((size_t)k * P + bin) * C + c0 + c * CH + cl);
                }
            }
        }
    }
    if (!TILE) return;
    __syncthreads();
    // contiguous write-back of the [cs_eff][P] tile: out + (k*C + c0)*P
    const int cs_eff = min(CS, C - c0);
    const int total = cs_eff * P;
    float *dst = out + ((size_t)k * C + c0) * P;
    if (VEC == 4 && ((total & 3) == 0) && ((((size_t)k * C + c0) * P) & 3) == 0) {
        const float4 *src4 = reinterpret_cast<const float4 *>(tile);
        for (int t = tid; t < total / 4; t += kRoiWarps * 32) stg_cs_f4(dst + 4 * t, src4[t]);
    } else {
        for (int t = tid; t < total; t += kRoiWarps * 32) dst[t] = tile[t];
    }
}

template <int VEC, int NCH>
static int launch_roi(const float *feat, const float *rois, float *out, int B, int C, int H, int W, int K,
                      int ph, int pw, float scale, int sr, int aligned, int out_layout, cudaStream_t st) {
    constexpr int CS = 32 * VEC * NCH;
    size_t smem = out_layout == 0 ? sizeof(float) * (size_t)min(C, CS) * ph * pw : 0;
    if (smem > 200 * 1024) return fail(VOD_E_UNSUPPORTED, "vod_roi_align_fwd: output tile %zu B too large", smem);
    auto kern = out_layout == 0 ? roi_align_kernel<VEC, NCH, true> : roi_align_kernel<VEC, NCH, false>;
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(K, ceil_div(C, CS));
    kern<<<grid, kRoiWarps * 32, smem, st>>>(feat, rois, out, B, C, H, W, K, ph, pw, scale, sr, aligned,
                                             out_layout); note_launch();
    return check_launch("vod_roi_align_fwd");
}

}  // namespace vod

using namespace vod;

extern "C" int vod_roi_align_fwd(const float *feat_nhwc, const float *rois, float *out, int B, int C,
                                 int H, int W, int K, int ph, int pw, float spatial_scale,
                                 int sampling_ratio, int aligned, int out_layout, vod_stream_t stream) {
    if (K == 0) return VOD_OK;
    VOD_REQUIRE(feat_nhwc && rois && out, "vod_roi_align_fwd: null pointer");
    VOD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && K > 0, "vod_roi_align_fwd: bad dims");
    VOD_REQUIRE(ph >= 1 && pw >= 1 && ph <= kMaxBins1D && pw <= kMaxBins1D,
                "vod_roi_align_fwd: output size %dx%d unsupported (max %d)", ph, pw, kMaxBins1D);
    VOD_REQUIRE(ph <= 32, "vod_roi_align_fwd: ph too large");
    VOD_REQUIRE(sampling_ratio <= kMaxEntries / 2, "vod_roi_align_fwd: sampling_ratio %d > %d", sampling_ratio,
                kMaxEntries / 2);
    VOD_REQUIRE(out_layout == 0 || out_layout == 1, "vod_roi_align_fwd: out_layout");
    cudaStream_t st = as_stream(stream);
    const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat_nhwc) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#define VOD_ROI_ARGS feat_nhwc, rois, out, B, C, H, W, K, ph, pw, spatial_scale, sampling_ratio, aligned, out_layout, st
    if (vec4) {
        // Few RoIs (a key frame's 300 proposals: 300 CTAs on 148 SMs) leave most of the machine idle behind per-CTA latency
        // (measured 40 us for 300 RoIs against 193 us for 4500): such launches are split into 128-channel slabs, 4x the CTAs
        // with a quarter of the work each.  Large launches keep one CTA per RoI (per-RoI set-up paid once for 2 KB per tap).
        // (slabs for every launch size were measured too: 274 us instead of 193 us for 4500 RoIs)
        const bool few = (long)K * 4 <= 16L * num_sms();
        if (C > 256 && !few) return launch_roi<4, 4>(VOD_ROI_ARGS);
        if (C > 128 && !few) return launch_roi<4, 2>(VOD_ROI_ARGS);
        return launch_roi<4, 1>(VOD_ROI_ARGS);
    }
    if (C > 64) return launch_roi<1, 4>(VOD_ROI_ARGS);
    return launch_roi<1, 1>(VOD_ROI_ARGS);
#undef VOD_ROI_ARGS
}
