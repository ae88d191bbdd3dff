// deform.cu -- the two memory-bound pieces of the fork's temporal-attention fusion (Denoising2Aggergator /
// TemporalAttentionFusion, mmtracking/mmtrack/models/aggregators/denoising2_aggregator.py:117-152):
//
//   (1) vod_mdcn_im2col: the sampling half of mmcv's modulated_deform_conv2d (DCNv2; mmcv-full 1.2.x, source not vendored:
//       restated from the published algorithm; the test oracle's restatement is pinned against
//       torchvision.ops.deform_conv2d).  The reference builds, for each of the T*T (reference frame i, frame t) pairs, the
//       offsets / masks with two convolutions of cat([x_t, x_i]) and chunk / cat / sigmoid passes (:72-79,:141-143).  Both
//       convolutions are linear, so the host computes them once per FRAME (conv(cat[a, b]) = conv_a(a) + conv_b(b)) and this
//       kernel takes the pair's raw offset/mask logits as the SUM of a per-t and a per-i map, applies the chunk / cat /
//       sigmoid semantics in registers, gathers the 4 bilinear corners and writes the modulated columns channels-last, ready
//       for one library GEMM with the [Cout, K*C] weight.  HBM-bound: reads x (L2-resident, 9x reuse) and 2 x 3*G*K logits per
//       pixel, writes K*C floats per pixel.
//   (2) vod_temporal_softmax_fuse: softmax over the frames of the correlation maps and the weighted sum of the frames'
//       features (:145-147) in one pass (online softmax): reads cor [I, T, E] once and x [T, E] per i, writes [I, E].
#include "common.cuh"

namespace vod {

// CTA = an 8 x 8 tile of output pixels of one image; thread = (pixel of the tile, tap k, VEC consecutive channels of one
// deformable group; VEC = 4 when C / G allows, else 1), channel index fastest so that a pixel's K*C column row is written as one
// contiguous run.  The tile's ~(8+2)^2-pixel input neighbourhood is re-read 9 x 4 times (taps x bilinear corners) out of L1
// instead of L2 (thread = flat index over the whole launch: 756 us at stage 1, the L2 -> SM sector rate of the gathers).
// Layouts: x [B, H, W, C]; p / q: raw conv_offset outputs [.., Ho, Wo, 3*G*K] (channels [0, 2GK): offsets, group-major,
// (dy, dx) interleaved per tap; [2GK, 3GK): mask logits); col [B*Ho*Wo, K*C] with column k*C + c.
constexpr int kDcnTile = 8;

template <int VEC>
__global__ void __launch_bounds__(256)
mdcn_im2col_kernel(const float *__restrict__ x, const float *__restrict__ p, const float *__restrict__ q, float *__restrict__ col,
                   int H, int W, int C, int Ho, int Wo, int G, int kh, int kw, int stride, int pad, int dil,
                   long p_batch_stride, long q_batch_stride) {
    const int K = kh * kw, C4 = C / VEC, Cg = C / G;
    const int b = blockIdx.z, ho0 = blockIdx.y * kDcnTile, wo0 = blockIdx.x * kDcnTile;
    const int items = kDcnTile * kDcnTile * K * C4;
    for (int idx = threadIdx.x; idx < items; idx += blockDim.x) {
        const int c4 = idx % C4;
        int r1 = idx / C4;
        const int k = r1 % K;
        r1 /= K;                                           // pixel of the tile
        const int ho = ho0 + r1 / kDcnTile, wo = wo0 + r1 % kDcnTile;
        if (ho >= Ho || wo >= Wo) continue;
        const long r = ((long)b * Ho + ho) * Wo + wo;      // output pixel (b, ho, wo)
        const int c = c4 * VEC, g = c / Cg;
        const long pix = (long)ho * Wo + wo;
        const int och = 3 * G * K;
        const float *pp = p + b * p_batch_stride + pix * och;
        const float *qq = q ? q + b * q_batch_stride + pix * och : nullptr;
        const int io = g * 2 * K + 2 * k, im = 2 * G * K + g * K + k;
        float dy = __ldg(pp + io), dx = __ldg(pp + io + 1), ml = __ldg(pp + im);
        if (qq) { dy += __ldg(qq + io); dx += __ldg(qq + io + 1); ml += __ldg(qq + im); }
        const float m = 1.f / (1.f + __expf(-ml));         // torch.sigmoid(mask) (:78)
        const float h = (float)(ho * stride - pad + (k / kw) * dil) + dy;
        const float w = (float)(wo * stride - pad + (k % kw) * dil) + dx;
        float v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = 0.f;
        if (h > -1.f && w > -1.f && h < (float)H && w < (float)W) {
            const float hf = floorf(h), wf = floorf(w);
            const int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;
            const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
            const float *xb = x + ((long)b * H * W) * C + c;
            // corners outside the map contribute 0 (mmcv dmcn_im2col_bilinear)
            const float a[4] = {(h0 >= 0 && w0 >= 0) ? hh * hw * m : 0.f, (h0 >= 0 && w1 <= W - 1) ? hh * lw * m : 0.f,
                                (h1 <= H - 1 && w0 >= 0) ? lh * hw * m : 0.f, (h1 <= H - 1 && w1 <= W - 1) ? lh * lw * m : 0.f};
            const long o[4] = {((long)h0 * W + w0) * C, ((long)h0 * W + w1) * C, ((long)h1 * W + w0) * C, ((long)h1 * W + w1) * C};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (a[j] == 0.f) continue;                 // outside (or zero weight): the address may be out of range
                if (VEC == 4) {
                    const float4 t = ldg_f4(xb + o[j]);
                    v[0] += a[j] * t.x; v[1 % VEC] += a[j] * t.y; v[2 % VEC] += a[j] * t.z; v[3 % VEC] += a[j] * t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[e] += a[j] * __ldg(xb + o[j] + e);
                }
            }
        }
        float *dst = col + (r * K + k) * (long)C + c;
        if (VEC == 4) stg_cs_f4(dst, make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
        else
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = v[e];
    }
}

// The 3x3 / stride 1 / pad 1 case of the fork's pack with 64-channel slabs (C % 64 == 0, C / G a multiple of 8 that divides
// 64): CTA = (8 x 8 output pixels, 64 channels = 64 / (C/G) whole deformable groups, image).  The slab's input neighbourhood
// (tile + 3 pixels of halo, 14 x 14 x 256 B = 49 KB, four CTAs per SM) is staged in shared memory with cp.async; each thread
// then owns (pixel, tap, 8 channels): one coordinate / sigmoid / weight set-up per 32 output bytes and 4 corners x 2
// conflict-free 128-bit shared loads (a corner outside the staged halo -- offsets beyond +-2 pixels -- is read from global
// memory instead).  Against the flat version: the 36 gathers per pixel and group hit shared memory instead of costing one L1
// tag look-up per 32-byte sector, which is what bounds that version (680 us at stage 1).
constexpr int kDcnHalo = 3, kDcnTS = kDcnTile + 2 * kDcnHalo, kDcnCS = 64;

// (2 or 3 CTAs per SM with a larger register budget: 370 us at stage 1 against 334 us for 4)
#ifndef VOD_DCN_MINB
#define VOD_DCN_MINB 4
#endif
__global__ void __launch_bounds__(256, VOD_DCN_MINB)
mdcn_im2col_tile_kernel(const float *__restrict__ x, const float *__restrict__ p, const float *__restrict__ q, float *__restrict__ col,
                        int H, int W, int C, int G, long p_batch_stride, long q_batch_stride) {
    extern __shared__ __align__(16) float s_x[];            // [TS][TS][CS]
    constexpr int K = 9, CS = kDcnCS, TS = kDcnTS;
    const int Cg = C / G;
    const int slabs = C / CS, b = blockIdx.z / slabs, c0 = (blockIdx.z % slabs) * CS, g0 = c0 / Cg;
    const int ho0 = blockIdx.y * kDcnTile, wo0 = blockIdx.x * kDcnTile;
    const int och = 3 * G * K;
    // ---- stage the input neighbourhood (zero-filled outside the map)
    for (int i = threadIdx.x; i < TS * TS * (CS / 4); i += blockDim.x) {
        const int c4 = i % (CS / 4), px = i / (CS / 4);
        const int hy = ho0 - kDcnHalo + px / TS, wx = wo0 - kDcnHalo + px % TS;
        const bool in = hy >= 0 && hy < H && wx >= 0 && wx < W;
        const float *src = in ? x + (((long)b * H + hy) * W + wx) * C + c0 + c4 * 4 : x;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_x + (size_t)i * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- thread = (pixel of the tile, tap, 8 channels)
    for (int it = threadIdx.x; it < kDcnTile * kDcnTile * K * (CS / 8); it += blockDim.x) {
        const int c8 = it % (CS / 8);
        int r1 = it / (CS / 8);
        const int k = r1 % K, pl = r1 / K;
        const int ho = ho0 + pl / kDcnTile, wo = wo0 + pl % kDcnTile;
        if (ho >= H || wo >= W) continue;
        const int cl = c8 * 8, g = g0 + cl / Cg;
        const long at = ((long)ho * W + wo) * och;
        const float *pp = p + b * p_batch_stride + at;
        float2 d = __ldg(reinterpret_cast<const float2 *>(pp + g * 18 + 2 * k));      // (dy, dx): even index, 8-byte aligned
        float ml = __ldg(pp + 2 * G * K + g * 9 + k);
        if (q) {
            const float *qq = q + b * q_batch_stride + at;
            const float2 d2 = __ldg(reinterpret_cast<const float2 *>(qq + g * 18 + 2 * k));
            d.x += d2.x; d.y += d2.y;
            ml += __ldg(qq + 2 * G * K + g * 9 + k);
        }
        const float m = __frcp_rn(1.f + __expf(-ml));
        const float h = (float)(ho - 1 + k / 3) + d.x, w = (float)(wo - 1 + k % 3) + d.y;
        // 16-byte chunk pair (cl/4, cl/4 + 1): read the odd chunk first when bit 3 of the chunk index is set, so that the 8
        // threads of a quarter-warp (8 consecutive chunk pairs) cover the 8 bank groups in both loads
        const int flip = (cl >> 5) & 1;
        float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;   // sums of the chunk read first / second
        if (h > -1.f && w > -1.f && h < (float)H && w < (float)W) {
            const float hf = floorf(h), wf = floorf(w);
            const int h0 = (int)hf, w0 = (int)wf;
            const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
            const float a[4] = {hh * hw * m, hh * lw * m, lh * hw * m, lh * lw * m};
            const int sy0 = h0 - (ho0 - kDcnHalo), sx0 = w0 - (wo0 - kDcnHalo);
            if (h0 >= 0 && h0 < H - 1 && w0 >= 0 && w0 < W - 1 && sy0 >= 0 && sy0 < TS - 1 && sx0 >= 0 && sx0 < TS - 1) {
                // common case: all four corners inside the map and inside the staged neighbourhood
                const float *sp = s_x + (sy0 * TS + sx0) * CS + cl + 4 * flip;
                const int o1 = 4 - 8 * flip;                  // offset of the chunk read second
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float *sq = sp + ((j >> 1) * TS + (j & 1)) * CS;
                    const float4 t0 = *reinterpret_cast<const float4 *>(sq);
                    const float4 t1 = *reinterpret_cast<const float4 *>(sq + o1);
                    u0.x += a[j] * t0.x; u0.y += a[j] * t0.y; u0.z += a[j] * t0.z; u0.w += a[j] * t0.w;
                    u1.x += a[j] * t1.x; u1.y += a[j] * t1.y; u1.z += a[j] * t1.z; u1.w += a[j] * t1.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int hy = h0 + (j >> 1), wx = w0 + (j & 1);
                    if (hy < 0 || hy > H - 1 || wx < 0 || wx > W - 1) continue;      // corners outside the map contribute 0
                    const int sy = hy - (ho0 - kDcnHalo), sx = wx - (wo0 - kDcnHalo);
                    float4 t0, t1;
                    if (sy >= 0 && sy < TS && sx >= 0 && sx < TS) {
                        const float *sp = s_x + (sy * TS + sx) * CS + cl;
                        t0 = *reinterpret_cast<const float4 *>(sp + 4 * flip);
                        t1 = *reinterpret_cast<const float4 *>(sp + 4 - 4 * flip);
                    } else {
                        const float *gp = x + (((long)b * H + hy) * W + wx) * C + c0 + cl;
                        t0 = ldg_f4(gp + 4 * flip);
                        t1 = ldg_f4(gp + 4 - 4 * flip);
                    }
                    u0.x += a[j] * t0.x; u0.y += a[j] * t0.y; u0.z += a[j] * t0.z; u0.w += a[j] * t0.w;
                    u1.x += a[j] * t1.x; u1.y += a[j] * t1.y; u1.z += a[j] * t1.z; u1.w += a[j] * t1.w;
                }
            }
        }
        float *dst = col + ((((long)b * H + ho) * W + wo) * K + k) * (long)C + c0 + cl;
        stg_cs_f4(dst + 4 * flip, u0);
        stg_cs_f4(dst + 4 - 4 * flip, u1);
    }
}

// thread = 4 consecutive elements of one output map i; one pass over the T frames with a running maximum
__global__ void __launch_bounds__(256)
temporal_softmax_fuse_kernel(const float *__restrict__ cor, const float *__restrict__ x, float *__restrict__ out, int I, int T,
                             long E4) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long)I * E4; idx += (long)gridDim.x * blockDim.x) {
        const long e = idx % E4;
        const int i = (int)(idx / E4);
        const float4 *cp = reinterpret_cast<const float4 *>(cor) + (long)i * T * E4 + e;
        const float4 *xp = reinterpret_cast<const float4 *>(x) + e;
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, den[4] = {0.f, 0.f, 0.f, 0.f}, num[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < T; ++t) {
            const float4 c4 = __ldcs(cp + (long)t * E4);
            const float4 x4 = __ldg(xp + (long)t * E4);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float nm = fmaxf(mx[j], cv[j]);
                const float s = __expf(mx[j] - nm), w = __expf(cv[j] - nm);     // first frame: exp(-inf) = 0
                den[j] = den[j] * s + w;
                num[j] = num[j] * s + w * xv[j];
                mx[j] = nm;
            }
        }
        reinterpret_cast<float4 *>(out)[idx] = make_float4(num[0] / den[0], num[1] / den[1], num[2] / den[2], num[3] / den[3]);
    }
}

}  // namespace vod

using namespace vod;

extern "C" int vod_mdcn_im2col(const float *x, const float *p, const float *q, float *col, int B, int C, int H, int W, int G,
                               int kh, int kw, int stride, int pad, int dil, int p_shared, int q_shared, vod_stream_t stream) {
    if (B == 0) return VOD_OK;
    VOD_REQUIRE(x && p && col, "vod_mdcn_im2col: null pointer");
    VOD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && G > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0 && dil > 0,
                "vod_mdcn_im2col: bad dims");
    VOD_REQUIRE(C % G == 0, "vod_mdcn_im2col: C must be a multiple of the deformable groups (C=%d G=%d)", C, G);
    const bool vec4 = (C / G) % 4 == 0;                    // 128-bit accesses when a group's channels allow it
    VOD_REQUIRE(!vec4 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(col)) & 15) == 0,
                "vod_mdcn_im2col: x and col must be 16-byte aligned");
    VOD_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 7) == 0, "vod_mdcn_im2col: p and q must be 8-byte aligned");
    const int Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    VOD_REQUIRE(Ho > 0 && Wo > 0, "vod_mdcn_im2col: empty output");
    VOD_REQUIRE(B <= 65535 && ceil_div(Ho, kDcnTile) <= 65535, "vod_mdcn_im2col: grid too large");
    const long och = 3L * G * kh * kw, map = (long)Ho * Wo * och;
    const dim3 grid(ceil_div(Wo, kDcnTile), ceil_div(Ho, kDcnTile), B);
    const int Cg = C / G;
    if (kh == 3 && kw == 3 && stride == 1 && pad == 1 && dil == 1 && C % kDcnCS == 0 && Cg % 8 == 0 && kDcnCS % Cg == 0 &&
        (long)B * (C / kDcnCS) <= 65535) {
        const size_t smem = (size_t)kDcnTS * kDcnTS * kDcnCS * sizeof(float);
        cudaFuncSetAttribute(mdcn_im2col_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const dim3 tgrid(grid.x, grid.y, B * (C / kDcnCS));
        mdcn_im2col_tile_kernel<<<tgrid, 256, smem, as_stream(stream)>>>(x, p, q, col, H, W, C, G, p_shared ? 0 : map, q_shared ? 0 : map);
    } else if (vec4)
        mdcn_im2col_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, p, q, col, H, W, C, Ho, Wo, G, kh, kw, stride, pad, dil,
                                                                  p_shared ? 0 : map, q_shared ? 0 : map);
    else
        mdcn_im2col_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, p, q, col, H, W, C, Ho, Wo, G, kh, kw, stride, pad, dil,
                                                                  p_shared ? 0 : map, q_shared ? 0 : map);
    note_launch();
    return check_launch("vod_mdcn_im2col");
}

extern "C" int vod_temporal_softmax_fuse(const float *cor, const float *x, float *out, int I, int T, long E, vod_stream_t stream) {
    if (I == 0 || E == 0) return VOD_OK;
    VOD_REQUIRE(cor && x && out, "vod_temporal_softmax_fuse: null pointer");
    VOD_REQUIRE(I > 0 && T > 0 && E > 0 && E % 4 == 0, "vod_temporal_softmax_fuse: bad dims (E must be a multiple of 4)");
    VOD_REQUIRE(((reinterpret_cast<uintptr_t>(cor) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "vod_temporal_softmax_fuse: pointers must be 16-byte aligned");
    const long total = (long)I * (E / 4);
    const int grid = (int)min((total + 255) / 256, (long)num_sms() * 32);
    temporal_softmax_fuse_kernel<<<grid, 256, 0, as_stream(stream)>>>(cor, x, out, I, T, E / 4);
    note_launch();
    return check_launch("vod_temporal_softmax_fuse");
}
