// warp.cu -- (2) flow-guided bilinear warp and the FGFA cosine-similarity weighting.
//   flow_warp        : mmtracking/mmtrack/core/motion/flow.py:4-41 (closed form)
//   embed weighting  : mmtracking/mmtrack/models/aggregators/embed_aggregator.py:71-81
//   fused variant    : the weighting re-warps the raw feature memory on the fly, as FGFA uses it at
//                      mmtracking/mmtrack/models/vid/fgfa.py:275-283
// All three are HBM-bound and NCHW-native (the tensors come from / go to cuDNN convs in NCHW):
// thread = pixel (w fastest) so every tap of a warp is a coalesced run of one channel plane.
#include "common.cuh"

namespace vod {

struct Taps {
    int i00, i01, i10, i11;   // plane offsets
    float w00, w01, w10, w11; // nw, ne, sw, se
};

// src index/lerp of ATen upsample_bilinear2d(align_corners=False) with a user scale factor
__device__ __forceinline__ void resize_src(int dst, float inv_scale, int in_size, int &i0, int &i1, float &l1) {
    float src = __fsub_rn(__fmul_rn((float)dst + 0.5f, inv_scale), 0.5f);
    if (src < 0.f) src = 0.f;
    int a = (int)src;
    if (a > in_size - 1) a = in_size - 1;
    i0 = a;
    i1 = a + (a < in_size - 1 ? 1 : 0);
    l1 = src - (float)a;
}

// Where the full-resolution flow comes from.  LOWRES: the flow network's last prediction [2, hl, wl] is upsampled on the fly --
// FlowNetSimple ends with interpolate(scale_factor=up, bilinear, align_corners=False) followed by two scalar multiplies
// (mmtracking/mmtrack/models/motion/flownet_simple.py:229-236) and flow_warp_feats immediately shrinks the result by 1/16,
// touching 4 of every 256 upsampled values: evaluating just those skips writing and re-reading the [N, 2, Hf, Wf] tensor.
struct FlowSrc {
    const float *p;       // full-res [2, Hf, Wf] of this frame, or low-res [2, hl, wl]
    int Hf, Wf;           // full-resolution size (virtual when lowres)
    int hl, wl;           // low-res size (lowres only)
    float up_inv, m1, m2; // 1 / upsample factor; the two post-multipliers, applied in the reference's order
    int lowres;
};
template <bool LOWRES>
__device__ __forceinline__ float flow_at(const FlowSrc &f, int ch, int Y, int X) {
    if (!LOWRES) return __ldg(f.p + ((size_t)ch * f.Hf + Y) * f.Wf + X);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_src(Y, f.up_inv, f.hl, y0, y1, ly);
    resize_src(X, f.up_inv, f.wl, x0, x1, lx);
    const float *q = f.p + (size_t)ch * f.hl * f.wl;
    const float a = __ldg(q + (size_t)y0 * f.wl + x0), b = __ldg(q + (size_t)y0 * f.wl + x1);
    const float c = __ldg(q + (size_t)y1 * f.wl + x0), d = __ldg(q + (size_t)y1 * f.wl + x1);
    // ATen upsample_bilinear2d: h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d)
    const float top = __fadd_rn(__fmul_rn(1.0f - lx, a), __fmul_rn(lx, b));
    const float bot = __fadd_rn(__fmul_rn(1.0f - lx, c), __fmul_rn(lx, d));
    const float v = __fadd_rn(__fmul_rn(1.0f - ly, top), __fmul_rn(ly, bot));
    return __fmul_rn(__fmul_rn(v, f.m1), f.m2);
}

// Sampling position of pixel (h, w) of frame n: resized+scaled flow -> normalised grid -> border-clamped bilinear, as the base
// plane offset, the steps to the east / south neighbours (0 at a clamped border) and the two lerp fractions.
struct TapF {
    int i00, dx, dy;
    float tx, ty;
};
__device__ __forceinline__ FlowSrc full_res_flow(const float *flow_n, int Hf, int Wf) {
    FlowSrc f;
    f.p = flow_n; f.Hf = Hf; f.Wf = Wf; f.hl = 0; f.wl = 0; f.up_inv = 1.f; f.m1 = 1.f; f.m2 = 1.f; f.lowres = 0;
    return f;
}

template <bool LOWRES>
__device__ __forceinline__ TapF make_tap_frac(const FlowSrc &src, int h, int w, int H, int W, float s, float inv_s) {
    const int Hf = src.Hf, Wf = src.Wf;
    int y0, y1, x0, x1;
    float ly, lx;
    resize_src(h, inv_s, Hf, y0, y1, ly);
    resize_src(w, inv_s, Wf, x0, x1, lx);
    float f[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float a = flow_at<LOWRES>(src, ch, y0, x0), b = flow_at<LOWRES>(src, ch, y0, x1);
        float c = flow_at<LOWRES>(src, ch, y1, x0), d = flow_at<LOWRES>(src, ch, y1, x1);
        float top = __fadd_rn(__fmul_rn(1.0f - lx, a), __fmul_rn(lx, b));
        float bot = __fadd_rn(__fmul_rn(1.0f - lx, c), __fmul_rn(lx, d));
        f[ch] = __fmul_rn(__fadd_rn(__fmul_rn(1.0f - ly, top), __fmul_rn(ly, bot)), s);
    }
    // grid = (w + fx) / W * 2 - 1  (flow.py:33-34), then align_corners=True un-normalisation + border clamp
    float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)w, f[0]), (float)W), 2.0f), 1.0f);
    float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)h, f[1]), (float)H), 2.0f), 1.0f);
    float px = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
    float py = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
    px = fminf(fmaxf(px, 0.f), (float)(W - 1));
    py = fminf(fmaxf(py, 0.f), (float)(H - 1));
    const int ix0 = (int)floorf(px), iy0 = (int)floorf(py);
    TapF t;
    t.tx = px - (float)ix0; t.ty = py - (float)iy0;
    // at the east / south border the neighbour is the pixel itself and the fraction is exactly 0 (px == W - 1): the products
    // below then give the +0 weights the reference's clamped grid_sample produces
    t.dx = ix0 + 1 <= W - 1 ? 1 : 0;
    t.dy = iy0 + 1 <= H - 1 ? W : 0;
    t.i00 = iy0 * W + ix0;
    return t;
}

__device__ __forceinline__ TapF make_tap_frac(const float *__restrict__ flow_n, int h, int w, int H, int W, int Hf,
                                              int Wf, float s, float inv_s) {
    return make_tap_frac<false>(full_res_flow(flow_n, Hf, Wf), h, w, H, W, s, inv_s);
}

__device__ __forceinline__ Taps taps_from_frac(const TapF &f) {
    Taps t;
    t.i00 = f.i00; t.i01 = f.i00 + f.dx; t.i10 = f.i00 + f.dy; t.i11 = f.i00 + f.dy + f.dx;
    t.w00 = (1.0f - f.tx) * (1.0f - f.ty);
    t.w01 = f.dx ? f.tx * (1.0f - f.ty) : 0.f;
    t.w10 = f.dy ? (1.0f - f.tx) * f.ty : 0.f;
    t.w11 = (f.dx && f.dy) ? f.tx * f.ty : 0.f;
    return t;
}

__device__ __forceinline__ Taps make_taps(const float *__restrict__ flow_n, int h, int w, int H, int W, int Hf,
                                          int Wf, float s, float inv_s) {
    const TapF f = make_tap_frac(flow_n, h, w, H, W, Hf, Wf, s, inv_s);
    Taps t;
    t.i00 = f.i00; t.i01 = f.i00 + f.dx; t.i10 = f.i00 + f.dy; t.i11 = f.i00 + f.dy + f.dx;
    t.w00 = (1.0f - f.tx) * (1.0f - f.ty);
    t.w01 = f.dx ? f.tx * (1.0f - f.ty) : 0.f;
    t.w10 = f.dy ? (1.0f - f.tx) * f.ty : 0.f;
    t.w11 = (f.dx && f.dy) ? f.tx * f.ty : 0.f;
    return t;
}

__device__ __forceinline__ float apply_taps(const float *__restrict__ plane, const Taps &t) {
    float v = __ldg(plane + t.i00) * t.w00;
    v = fmaf(__ldg(plane + t.i01), t.w01, v);
    v = fmaf(__ldg(plane + t.i10), t.w10, v);
    v = fmaf(__ldg(plane + t.i11), t.w11, v);
    return v;
}

// Tile shapes are overridable at build time for experiments (python -m ...build --probes with VOD_EXTRA_DEFINES).  Measured in
// round 2 at T = 31 (us): warp 128 px x 64 ch 98 | 256 x 32 114 | 512 x 16 127; weighting (cos + apply) 128 px x 4 groups 121 |
// 256 x 2 136 | 512 x 1 135 -- longer contiguous runs per plane do NOT help, the small tiles' extra CTAs do.
#ifndef VOD_WARP_PIX
#define VOD_WARP_PIX 128
#define VOD_WARP_CH 64
#endif
constexpr int kWarpPix = VOD_WARP_PIX;   // pixels per CTA
constexpr int kWarpCh = VOD_WARP_CH;     // channels per CTA

// grid (pixel blocks, channel chunks, N)
// (min 16 CTAs per SM = 32 registers: the kernel lives on its 2048 resident threads; at 48 registers -- 1280 threads -- the
// same code ran 117 us instead of 90 us at T = 31)
template <bool LOWRES>
__global__ void __launch_bounds__(kWarpPix, 2048 / kWarpPix)
flow_warp_kernel(const float *__restrict__ x, const float *__restrict__ flow, float *__restrict__ out, int C,
                 int H, int W, FlowSrc src, float s, float inv_s, int x_frames) {
    const int n = blockIdx.z;
    const int p = blockIdx.x * kWarpPix + threadIdx.x;
    const int HW = H * W;
    if (p >= HW) return;
    src.p = flow + (size_t)n * 2 * (LOWRES ? src.hl * src.wl : src.Hf * src.Wf);
    const Taps t = taps_from_frac(make_tap_frac<LOWRES>(src, p / W, p % W, H, W, s, inv_s));
    const int c0 = blockIdx.y * kWarpCh, c1 = min(C, c0 + kWarpCh);
    // x_frames == 1: every flow warps the SAME map (DFF: the non-key frames of an interval share the key frame's features)
    const float *xp = x + ((size_t)(x_frames == 1 ? 0 : n) * C + c0) * HW;
    float *op = out + ((size_t)n * C + c0) * HW + p;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = apply_taps(xp + (size_t)(c - c0 + q) * HW, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) __stcs(op + (size_t)(c - c0 + q) * HW, v[q]);
    }
    for (; c < c1; ++c) __stcs(op + (size_t)(c - c0) * HW, apply_taps(xp + (size_t)(c - c0) * HW, t));
}

// ---------------------------------------------------------------------------------------------
// Embed weighting.  CTA = 8 pixels x 32 channel lanes (256 threads).
//   phase A: per frame t, dot(e_t, e_k), |e_t|^2, |e_k|^2 over C  -> cos_t -> softmax over t
//   phase B: out[c] = sum_t w_t * ref_x[t][c]   (or the on-the-fly warp of raw_x[t][c])
constexpr int kEwPix = 8;
constexpr int kEwLanes = 32;
constexpr int kEwTch = 8;  // frames per register pass

template <bool FUSED_WARP>
__global__ void __launch_bounds__(kEwPix *kEwLanes)
embed_weighted_sum_kernel(const float *__restrict__ key_emb, const float *__restrict__ ref_emb,
                          const float *__restrict__ ref_x, const float *__restrict__ flow,
                          const float *__restrict__ key_x, int key_slot, float *__restrict__ out, int T,
                          int C, int Cx, int H, int W, int Hf, int Wf, float s, float inv_s) {
    extern __shared__ float sm[];
    // layout: wts[T][8] | red[3][32][8] | (FUSED) taps[T][8] (8 words each)
    float *wts = sm;
    float *red = wts + (size_t)T * kEwPix;
    Taps *taps = reinterpret_cast<Taps *>(red + 3 * kEwLanes * kEwPix);
    const int HW = H * W;
    const int pl = threadIdx.x % kEwPix, cg = threadIdx.x / kEwPix;
    const int p = blockIdx.x * kEwPix + pl;
    const bool pv = p < HW;
    const int pc = pv ? p : HW - 1;

    // |e_k|^2
    float kk = 0.f;
    for (int c = cg; c < C; c += kEwLanes) { float v = __ldg(key_emb + (size_t)c * HW + pc); kk = fmaf(v, v, kk); }
    red[(0 * kEwLanes + cg) * kEwPix + pl] = kk;
    __syncthreads();
    if (cg == 0) {
        float a = 0.f;
        for (int g = 0; g < kEwLanes; ++g) a += red[(0 * kEwLanes + g) * kEwPix + pl];
        red[(2 * kEwLanes + 0) * kEwPix + pl] = a;  // stash |e_k|^2
    }
    __syncthreads();
    const float knorm = sqrtf(red[(2 * kEwLanes + 0) * kEwPix + pl]);
    __syncthreads();

    for (int t0 = 0; t0 < T; t0 += kEwTch) {
        float dot[kEwTch], nn[kEwTch];
#pragma unroll
        for (int q = 0; q < kEwTch; ++q) { dot[q] = 0.f; nn[q] = 0.f; }
        for (int c = cg; c < C; c += kEwLanes) {
            const float kv = __ldg(key_emb + (size_t)c * HW + pc);
#pragma unroll
            for (int q = 0; q < kEwTch; ++q) {
                if (t0 + q < T) {
                    float v = __ldg(ref_emb + ((size_t)(t0 + q) * C + c) * HW + pc);
                    dot[q] = fmaf(v, kv, dot[q]);
                    nn[q] = fmaf(v, v, nn[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kEwTch; ++q) {
            if (t0 + q < T) {  // uniform
                red[(0 * kEwLanes + cg) * kEwPix + pl] = dot[q];
                red[(1 * kEwLanes + cg) * kEwPix + pl] = nn[q];
                __syncthreads();
                if (cg == 0) {
                    float d = 0.f, n2 = 0.f;
                    for (int g = 0; g < kEwLanes; ++g) {
                        d += red[(0 * kEwLanes + g) * kEwPix + pl];
                        n2 += red[(1 * kEwLanes + g) * kEwPix + pl];
                    }
                    // (e_t/|e_t|) . (e_k/|e_k|): no epsilon, as the reference (embed_aggregator.py:71,76)
                    wts[(t0 + q) * kEwPix + pl] = d / (sqrtf(n2) * knorm);
                }
                __syncthreads();
            }
        }
    }
    // softmax over t (thread per pixel), and the warp taps for the fused variant
    if (cg == 0) {
        float m = -INFINITY;
        for (int t = 0; t < T; ++t) m = fmaxf(m, wts[t * kEwPix + pl]);
        float sum = 0.f;
        for (int t = 0; t < T; ++t) { float e = expf(wts[t * kEwPix + pl] - m); wts[t * kEwPix + pl] = e; sum += e; }
        for (int t = 0; t < T; ++t) wts[t * kEwPix + pl] = wts[t * kEwPix + pl] / sum;
    }
    if (FUSED_WARP) {
        for (int t = cg; t < T; t += kEwLanes)
            taps[t * kEwPix + pl] = make_taps(flow + (size_t)t * 2 * Hf * Wf, pc / W, pc % W, H, W, Hf, Wf, s, inv_s);
    }
    __syncthreads();
    if (!pv) return;
    for (int c = cg; c < Cx; c += kEwLanes) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            float v;
            if (FUSED_WARP) {
                if (t == key_slot) v = __ldg(key_x + (size_t)c * HW + p);
                else v = apply_taps(ref_x + ((size_t)t * Cx + c) * HW, taps[t * kEwPix + pl]);
            } else {
                v = __ldg(ref_x + ((size_t)t * Cx + c) * HW + p);
            }
            acc = fmaf(v, wts[t * kEwPix + pl], acc);
        }
        __stcs(out + (size_t)c * HW + p, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Two-kernel form of the (un-fused) embed weighting, used when the shape is large enough to fill the machine:
//   embed_cos_kernel   grid (pixel blocks, T): cos(e_t, e_k) per pixel; thread = pixel (coalesced plane rows),
//                      4 channel groups per CTA reduced through shared memory
//   embed_apply_kernel grid (pixel blocks, channel chunks): softmax over t recomputed per CTA from the T cosines of
//                      its pixels (cheap), then out[c] = sum_t w_t * ref_x[t][c]
// Every input element is read once from HBM (the key embedding is re-read per frame, from L2).
#ifndef VOD_EC_PIX
#define VOD_EC_PIX 128
#define VOD_EC_GROUPS 4
#endif
constexpr int kEcPix = VOD_EC_PIX, kEcGroups = VOD_EC_GROUPS;

#ifndef VOD_EC_MINB
#define VOD_EC_MINB 1
#endif
__global__ void __launch_bounds__(kEcPix *kEcGroups, VOD_EC_MINB)
embed_cos_kernel(const float *__restrict__ key_emb, const float *__restrict__ ref_emb, float *__restrict__ cosv, int C,
                 int HW) {
    __shared__ float red[3][kEcGroups][kEcPix];
    const int t = blockIdx.y;
    const int pl = threadIdx.x % kEcPix, g = threadIdx.x / kEcPix;
    const int p = blockIdx.x * kEcPix + pl;
    const int pc = min(p, HW - 1);
    const float *kp = key_emb + pc, *rp = ref_emb + (size_t)t * C * HW + pc;
    float dot = 0.f, nn = 0.f, kk = 0.f;
    const int c0 = g * (C / kEcGroups), c1 = (g == kEcGroups - 1) ? C : c0 + C / kEcGroups;
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
        float kv[8], rv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { kv[q] = __ldg(kp + (size_t)(c + q) * HW); rv[q] = __ldg(rp + (size_t)(c + q) * HW); }
#pragma unroll
        for (int q = 0; q < 8; ++q) { dot = fmaf(rv[q], kv[q], dot); nn = fmaf(rv[q], rv[q], nn); kk = fmaf(kv[q], kv[q], kk); }
    }
    for (; c < c1; ++c) {
        const float kv = __ldg(kp + (size_t)c * HW), rv = __ldg(rp + (size_t)c * HW);
        dot = fmaf(rv, kv, dot); nn = fmaf(rv, rv, nn); kk = fmaf(kv, kv, kk);
    }
    red[0][g][pl] = dot; red[1][g][pl] = nn; red[2][g][pl] = kk;
    __syncthreads();
    if (g == 0 && p < HW) {
        float d = 0.f, n2 = 0.f, k2 = 0.f;
#pragma unroll
        for (int q = 0; q < kEcGroups; ++q) { d += red[0][q][pl]; n2 += red[1][q][pl]; k2 += red[2][q][pl]; }
        cosv[(size_t)t * HW + p] = d / (sqrtf(n2) * sqrtf(k2));   // no epsilon, as the reference
    }
}

#ifndef VOD_EA_PIX
#define VOD_EA_PIX 128
#define VOD_EA_GROUPS 4
#endif
constexpr int kEaPix = VOD_EA_PIX, kEaGroups = VOD_EA_GROUPS, kEaChPerGroup = 4, kEaCh = kEaGroups * kEaChPerGroup;
// CTA = 128 pixels x 4 channel groups (512 threads, 16 channels).  The softmax over t is computed once per CTA (group 0)
// and shared; every thread then streams T frames of 4 channels with 8 independent loads in flight per channel pair
// (the 128-thread version kept 16 warps per SM busy at 22 % of DRAM bandwidth: too few bytes in flight).
#ifndef VOD_EA_MINB
#define VOD_EA_MINB 1
#endif
__global__ void __launch_bounds__(kEaPix *kEaGroups, VOD_EA_MINB)
embed_apply_kernel(const float *__restrict__ cosv, const float *__restrict__ ref_x, float *__restrict__ out, int T, int Cx,
                   int HW) {
    extern __shared__ float wts[];   // [T][kEaPix]
    const int pl = threadIdx.x % kEaPix, g = threadIdx.x / kEaPix;
    const int p = blockIdx.x * kEaPix + pl;
    const int pc = min(p, HW - 1);
    if (g == 0) {
        float m = -INFINITY;
        for (int t = 0; t < T; ++t) { const float v = __ldg(cosv + (size_t)t * HW + pc); wts[t * kEaPix + pl] = v; m = fmaxf(m, v); }
        float sum = 0.f;
        for (int t = 0; t < T; ++t) { const float e = expf(wts[t * kEaPix + pl] - m); wts[t * kEaPix + pl] = e; sum += e; }
        for (int t = 0; t < T; ++t) wts[t * kEaPix + pl] = wts[t * kEaPix + pl] / sum;
    }
    __syncthreads();
    if (p >= HW) return;
    const int c0 = blockIdx.y * kEaCh + g * kEaChPerGroup, c1 = min(Cx, c0 + kEaChPerGroup);
    const size_t fs = (size_t)Cx * HW;   // frame stride
    for (int c = c0; c < c1; c += 2) {
        float a0 = 0.f, a1 = 0.f;
        const bool two = c + 1 < c1;
        const float *x0 = ref_x + (size_t)c * HW + p;
        const float *x1 = two ? x0 + HW : x0;
        int t = 0;
        for (; t + 8 <= T; t += 8) {
            float u[8], v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { u[q] = __ldg(x0 + (size_t)(t + q) * fs); v[q] = __ldg(x1 + (size_t)(t + q) * fs); }
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float w = wts[(t + q) * kEaPix + pl]; a0 = fmaf(u[q], w, a0); a1 = fmaf(v[q], w, a1); }
        }
        for (; t < T; ++t) {
            const float w = wts[t * kEaPix + pl];
            a0 = fmaf(__ldg(x0 + (size_t)t * fs), w, a0);
            a1 = fmaf(__ldg(x1 + (size_t)t * fs), w, a1);
        }
        __stcs(out + (size_t)c * HW + p, a0);
        if (two) __stcs(out + (size_t)(c + 1) * HW + p, a1);
    }
}

// Fused form of the weighting (north_star (2): the flow-guided warp fused with the cosine weighting): the weighted operand is
// the bilinear warp of the RAW feature memory, recomputed on the fly -- the warped tensor is not re-read.  Same tiling as
// embed_apply_kernel (128 pixels x 16 channels, 512 threads); the (frame, pixel) taps are computed once per CTA into shared
// memory in a compact 16-byte form (base offset, +1 / +W steps, the two lerp fractions: at a clamped border the fraction is
// exactly 0, so plain bilinear weights reproduce make_taps' zeroed weights bit for bit).  The first fused kernel ran one CTA
// per 8 pixels x all channels (300 CTAs, 10 % of DRAM bandwidth, 426 us at T = 31); this one is a machine-filling grid.
struct TapC {
    int i00, step;      // step: bit 0 = step to the east neighbour (0 / 1), bits 1.. = step to the south neighbour (0 / W)
    float tx, ty;
};

__global__ void __launch_bounds__(kEaPix *kEaGroups)
embed_apply_warp_kernel(const float *__restrict__ cosv, const float *__restrict__ raw_x, const float *__restrict__ flow,
                        const float *__restrict__ key_x, int key_slot, float *__restrict__ out, int T, int Cx, int H, int W,
                        int Hf, int Wf, float s, float inv_s) {
    extern __shared__ float sm_dyn[];
    const int HW = H * W;
    float *wts = sm_dyn;                                             // [T][kEaPix]
    TapC *taps = reinterpret_cast<TapC *>(wts + (size_t)T * kEaPix);  // [T][kEaPix]
    const int pl = threadIdx.x % kEaPix, g = threadIdx.x / kEaPix;
    const int p = blockIdx.x * kEaPix + pl;
    const int pc = min(p, HW - 1);
    if (g == 0) {
        float m = -INFINITY;
        for (int t = 0; t < T; ++t) { const float v = __ldg(cosv + (size_t)t * HW + pc); wts[t * kEaPix + pl] = v; m = fmaxf(m, v); }
        float sum = 0.f;
        for (int t = 0; t < T; ++t) { const float e = expf(wts[t * kEaPix + pl] - m); wts[t * kEaPix + pl] = e; sum += e; }
        for (int t = 0; t < T; ++t) wts[t * kEaPix + pl] = wts[t * kEaPix + pl] / sum;
    }
    for (int t = g; t < T; t += kEaGroups) {
        TapC c;
        if (t == key_slot) {
            c.i00 = pc; c.step = 0; c.tx = 0.f; c.ty = 0.f;          // the key frame's own slot is taken un-warped (fgfa.py:281)
        } else {
            const TapF tp = make_tap_frac(flow + (size_t)t * 2 * Hf * Wf, pc / W, pc % W, H, W, Hf, Wf, s, inv_s);
            c.i00 = tp.i00; c.step = tp.dx | (tp.dy << 1);
            c.tx = tp.tx; c.ty = tp.ty;
        }
        taps[t * kEaPix + pl] = c;
    }
    __syncthreads();
    if (p >= HW) return;
    const int c0 = blockIdx.y * kEaCh + g * kEaChPerGroup;
    const size_t fs = (size_t)Cx * HW;   // frame stride
    // the tap of a (frame, pixel) is fetched and its four weights are formed ONCE for the thread's kEaChPerGroup channels:
    // 16 independent loads in flight per frame
    float acc[kEaChPerGroup];
    bool on[kEaChPerGroup];
#pragma unroll
    for (int q = 0; q < kEaChPerGroup; ++q) { acc[q] = 0.f; on[q] = c0 + q < Cx; }
    for (int t = 0; t < T; ++t) {
        const TapC tc = taps[t * kEaPix + pl];
        const float w = wts[t * kEaPix + pl];
        const float *b = (t == key_slot ? key_x : raw_x + (size_t)t * fs) + (size_t)c0 * HW + tc.i00;
        const int dx = tc.step & 1, dy = tc.step >> 1;
        const float w00 = (1.0f - tc.tx) * (1.0f - tc.ty), w01 = tc.tx * (1.0f - tc.ty);
        const float w10 = (1.0f - tc.tx) * tc.ty, w11 = tc.tx * tc.ty;
        float u00[kEaChPerGroup], u01[kEaChPerGroup], u10[kEaChPerGroup], u11[kEaChPerGroup];
#pragma unroll
        for (int q = 0; q < kEaChPerGroup; ++q) {
            const float *bq = on[q] ? b + (size_t)q * HW : b;
            u00[q] = __ldg(bq); u01[q] = __ldg(bq + dx); u10[q] = __ldg(bq + dy); u11[q] = __ldg(bq + dy + dx);
        }
#pragma unroll
        for (int q = 0; q < kEaChPerGroup; ++q) {
            float u = u00[q] * w00; u = fmaf(u01[q], w01, u); u = fmaf(u10[q], w10, u); u = fmaf(u11[q], w11, u);
            acc[q] = fmaf(u, w, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < kEaChPerGroup; ++q)
        if (on[q]) __stcs(out + (size_t)(c0 + q) * HW + p, acc[q]);
}

static void flow_scale(int W, int Wf, float &s, float &inv_s) {
    double sd = (double)W / (double)Wf;  // python float scale_factor, flow.py:17
    s = (float)sd;
    inv_s = (float)(1.0 / sd);           // ATen: static_cast<float>(1.0 / scale_factor)
}

}  // namespace vod

using namespace vod;

extern "C" int vod_flow_warp_shared(const float *x, const float *flow, float *out, int N, int x_frames, int C, int H, int W,
                                    int Hf, int Wf, vod_stream_t stream) {
    if (N == 0 || C == 0) return VOD_OK;
    VOD_REQUIRE(x && flow && out, "vod_flow_warp: null pointer");
    VOD_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Hf > 0 && Wf > 0, "vod_flow_warp: bad dims");
    VOD_REQUIRE(N <= 65535, "vod_flow_warp: N too large");
    VOD_REQUIRE(x_frames == 1 || x_frames == N, "vod_flow_warp: x must hold 1 (shared) or N=%d maps, got %d", N, x_frames);
    float s, inv_s;
    flow_scale(W, Wf, s, inv_s);
    dim3 grid(ceil_div(H * W, kWarpPix), ceil_div(C, kWarpCh), N);
    FlowSrc src;
    src.p = nullptr; src.Hf = Hf; src.Wf = Wf; src.hl = 0; src.wl = 0; src.up_inv = 1.f; src.m1 = 1.f; src.m2 = 1.f; src.lowres = 0;
    flow_warp_kernel<false><<<grid, kWarpPix, 0, as_stream(stream)>>>(x, flow, out, C, H, W, src, s, inv_s, x_frames); note_launch();
    return check_launch("vod_flow_warp");
}

extern "C" int vod_flow_warp_lowres(const float *x, const float *flow_lr, float *out, int N, int x_frames, int C, int H, int W,
                                    int hl, int wl, int Hf, int Wf, double up_scale, float mult1, float mult2,
                                    vod_stream_t stream) {
    if (N == 0 || C == 0) return VOD_OK;
    VOD_REQUIRE(x && flow_lr && out, "vod_flow_warp_lowres: null pointer");
    VOD_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && hl > 0 && wl > 0 && Hf > 0 && Wf > 0 && up_scale > 0, "vod_flow_warp_lowres: bad dims");
    VOD_REQUIRE(N <= 65535, "vod_flow_warp_lowres: N too large");
    VOD_REQUIRE(x_frames == 1 || x_frames == N, "vod_flow_warp_lowres: x must hold 1 (shared) or N=%d maps, got %d", N, x_frames);
    float s, inv_s;
    flow_scale(W, Wf, s, inv_s);
    dim3 grid(ceil_div(H * W, kWarpPix), ceil_div(C, kWarpCh), N);
    FlowSrc src;
    src.p = nullptr; src.Hf = Hf; src.Wf = Wf; src.hl = hl; src.wl = wl;
    src.up_inv = (float)(1.0 / up_scale);      // ATen: static_cast<float>(1.0 / scale_factor)
    src.m1 = mult1; src.m2 = mult2; src.lowres = 1;
    flow_warp_kernel<true><<<grid, kWarpPix, 0, as_stream(stream)>>>(x, flow_lr, out, C, H, W, src, s, inv_s, x_frames); note_launch();
    return check_launch("vod_flow_warp_lowres");
}

extern "C" int vod_flow_warp(const float *x, const float *flow, float *out, int N, int C, int H, int W, int Hf,
                             int Wf, vod_stream_t stream) {
    return vod_flow_warp_shared(x, flow, out, N, N, C, H, W, Hf, Wf, stream);
}

static int launch_embed(bool fused, const float *key_emb, const float *ref_emb, const float *ref_x,
                        const float *flow, const float *key_x, int key_slot, float *out, int T, int C, int Cx,
                        int H, int W, int Hf, int Wf, void *ws, size_t ws_bytes, vod_stream_t stream) {
    VOD_REQUIRE(key_emb && ref_emb && ref_x && out, "vod_embed_weighted_sum: null pointer");
    VOD_REQUIRE(T > 0 && C > 0 && Cx > 0 && H > 0 && W > 0, "vod_embed_weighted_sum: bad dims");
    size_t smem = sizeof(float) * ((size_t)T * kEwPix + 3 * kEwLanes * kEwPix) + (fused ? sizeof(Taps) * (size_t)T * kEwPix : 0);
    VOD_REQUIRE(smem <= 200 * 1024, "vod_embed_weighted_sum: T=%d too large", T);
    float s = 1.f, inv_s = 1.f;
    if (fused) flow_scale(W, Wf, s, inv_s);
    dim3 grid(ceil_div(H * W, kEwPix));
    if (!fused && ws && ws_bytes >= sizeof(float) * (size_t)T * H * W && (size_t)T * kEaPix * sizeof(float) <= 200 * 1024) {
        const int HW = H * W;
        float *cosv = reinterpret_cast<float *>(ws);   // [T, HW] per-pixel cosines
        embed_cos_kernel<<<dim3(ceil_div(HW, kEcPix), T), kEcPix * kEcGroups, 0, as_stream(stream)>>>(key_emb, ref_emb, cosv, C, HW);
        note_launch();
        const size_t sm = sizeof(float) * (size_t)T * kEaPix;
        if (sm > 40 * 1024) cudaFuncSetAttribute(embed_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        embed_apply_kernel<<<dim3(ceil_div(HW, kEaPix), ceil_div(Cx, kEaCh)), kEaPix * kEaGroups, sm, as_stream(stream)>>>(cosv, ref_x, out, T, Cx, HW);
        note_launch();
        return check_launch("vod_embed_weighted_sum(2-kernel)");
    }
    if (fused && ws && ws_bytes >= sizeof(float) * (size_t)T * H * W &&
        (size_t)T * kEaPix * (sizeof(float) + sizeof(TapC)) <= 200 * 1024) {
        // cosines per (frame, pixel), then softmax + weighted sum of the on-the-fly warp: two machine-filling kernels
        const int HW = H * W;
        float *cosv = reinterpret_cast<float *>(ws);
        embed_cos_kernel<<<dim3(ceil_div(HW, kEcPix), T), kEcPix * kEcGroups, 0, as_stream(stream)>>>(key_emb, ref_emb, cosv, C, HW);
        note_launch();
        const size_t sm = (size_t)T * kEaPix * (sizeof(float) + sizeof(TapC));
        if (sm > 40 * 1024) cudaFuncSetAttribute(embed_apply_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        embed_apply_warp_kernel<<<dim3(ceil_div(HW, kEaPix), ceil_div(Cx, kEaCh)), kEaPix * kEaGroups, sm, as_stream(stream)>>>(
            cosv, ref_x, flow, key_x, key_slot, out, T, Cx, H, W, Hf, Wf, s, inv_s);
        note_launch();
        return check_launch("vod_fgfa_warp_weighted_sum(2-kernel)");
    }
    if (fused) {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(embed_weighted_sum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        embed_weighted_sum_kernel<true><<<grid, kEwPix * kEwLanes, smem, as_stream(stream)>>>(
            key_emb, ref_emb, ref_x, flow, key_x, key_slot, out, T, C, Cx, H, W, Hf, Wf, s, inv_s); note_launch();
    } else {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(embed_weighted_sum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        embed_weighted_sum_kernel<false><<<grid, kEwPix * kEwLanes, smem, as_stream(stream)>>>(
            key_emb, ref_emb, ref_x, nullptr, nullptr, -1, out, T, C, Cx, H, W, 1, 1, s, inv_s); note_launch();
    }
    return check_launch("vod_embed_weighted_sum");
}

extern "C" int vod_embed_weighted_sum(const float *key_emb, const float *ref_emb, const float *ref_x, float *out,
                                      int T, int C, int Cx, int HW, void *ws, size_t ws_bytes, vod_stream_t stream) {
    return launch_embed(false, key_emb, ref_emb, ref_x, nullptr, nullptr, -1, out, T, C, Cx, 1, HW, 1, 1, ws, ws_bytes, stream);
}

extern "C" int vod_fgfa_warp_weighted_sum(const float *key_emb, const float *ref_emb, const float *raw_x,
                                          const float *flow, const float *key_x, int key_slot, float *out, int T,
                                          int C, int Cx, int H, int W, int Hf, int Wf, void *ws, size_t ws_bytes,
                                          vod_stream_t stream) {
    VOD_REQUIRE(flow, "vod_fgfa_warp_weighted_sum: null flow");
    VOD_REQUIRE(key_slot < 0 || key_x, "vod_fgfa_warp_weighted_sum: key_x required with key_slot");
    return launch_embed(true, key_emb, ref_emb, raw_x, flow, key_x, key_slot, out, T, C, Cx, H, W, Hf, Wf, ws, ws_bytes, stream);
}
