/*
 * oracle/vod_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C CPU restatement of the exact-arithmetic pieces of the reference's
 * multi-frame feature-aggregation hot path.  The reference is 100% Python; the
 * arithmetic for RoIAlign and NMS lives in the un-vendored third party
 * mmcv-full (pinned >=1.2.4,<=1.4.0 at mmdetection/mmdet/__init__.py:18-26).
 * Their published algorithms are restated here and pinned against
 * torchvision.ops.{roi_align,nms} (the executable truth in this image) by
 * tests/test_oracle.py, and against golden vectors generated from the
 * unmodified reference Python files (tests/golden/make_golden.py).
 *
 * Compiled with -O2 -ffp-contract=off so no multiply-add is fused: the IoU and
 * bilinear arithmetic is the un-contracted fp32 sequence a CPU build of
 * mmcv/torchvision executes.
 *
 * Call sites in the reference that define the parameters:
 *   RoIAlign   mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55
 *              mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:72-75
 *   NMS        mmdetection/mmdet/core/post_processing/bbox_nms.py:84
 *              mmdetection/mmdet/models/dense_heads/rpn_head.py:233-235
 *   flow warp  mmtracking/mmtrack/core/motion/flow.py:4-41
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RoIAlign
 * mmcv roi_align forward, pool_mode='avg' (SURVEY Appendix A.1).
 * feat: [B,C,H,W] fp32 contiguous; rois: [K,5] (batch, x1,y1,x2,y2) image px;
 * out : [K,C,ph,pw].
 */
static float bilinear_tap(const float *plane, int H, int W, float y, float x)
{
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.0f;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
    float ly = y - (float)yl, lx = x - (float)xl;
    float hy = 1.0f - ly, hx = 1.0f - lx;
    float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    return w1 * plane[yl * W + xl] + w2 * plane[yl * W + xh] +
           w3 * plane[yh * W + xl] + w4 * plane[yh * W + xh];
}

void oracle_roi_align(const float *feat, const float *rois, float *out,
                      int B, int C, int H, int W, int K, int ph, int pw,
                      float spatial_scale, int sampling_ratio, int aligned)
{
    (void)B;
#pragma omp parallel for schedule(dynamic, 4)
    for (int k = 0; k < K; ++k) {
        const float *r = rois + 5 * k;
        int b = (int)r[0];
        float off = aligned ? 0.5f : 0.0f;
        float rsw = r[1] * spatial_scale - off;
        float rsh = r[2] * spatial_scale - off;
        float rew = r[3] * spatial_scale - off;
        float reh = r[4] * spatial_scale - off;
        float rw = rew - rsw, rh = reh - rsh;
        if (!aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
        float bh = rh / (float)ph, bw = rw / (float)pw;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
        float count = (float)(gh * gw > 1 ? gh * gw : 1);
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((size_t)b * C + c) * H * W;
            float *o = out + ((size_t)k * C + c) * ph * pw;
            for (int i = 0; i < ph; ++i)
                for (int j = 0; j < pw; ++j) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = rsh + (float)i * bh + ((float)iy + 0.5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = rsw + (float)j * bw + ((float)ix + 0.5f) * bw / (float)gw;
                            acc += bilinear_tap(plane, H, W, y, x);
                        }
                    }
                    o[i * pw + j] = acc / count;
                }
        }
    }
}

/* ----------------------------------------------------------------------- NMS
 * mmcv nms (offset=0) == torchvision nms: greedy over descending score, a kept
 * box suppresses later boxes with IoU > thr (strict).
 * order: indices sorted by descending score (n). boxes [n,4] already carry the
 * per-class coordinate offset when the caller is batched_nms.
 * keep_out receives the kept ORIGINAL indices in descending-score order.
 * returns number kept.
 */
int oracle_nms_sorted(const float *boxes, const int64_t *order, int n,
                      float thr, int64_t *keep_out)
{
    uint8_t *sup = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        int64_t i = order[a];
        if (sup[i]) continue;
        keep_out[nk++] = i;
        const float *bi = boxes + 4 * i;
        float ia = area[i];
        for (int c = a + 1; c < n; ++c) {
            int64_t j = order[c];
            if (sup[j]) continue;
            const float *bj = boxes + 4 * j;
            float xx1 = fmaxf(bi[0], bj[0]), yy1 = fmaxf(bi[1], bj[1]);
            float xx2 = fminf(bi[2], bj[2]), yy2 = fminf(bi[3], bj[3]);
            float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if (ovr > thr) sup[j] = 1;
        }
    }
    free(sup);
    free(area);
    return nk;
}

/* ----------------------------------------------------------------- flow warp
 * Closed form of mmtracking/mmtrack/core/motion/flow.py:4-41:
 *   1. flow' = s * bilinear_resize(flow, scale_factor=s, align_corners=False),
 *      s = Wx / Wf (width ratio used for both axes, flow.py:17). ATen's
 *      upsample_bilinear2d with a user scale factor maps dst->src with
 *      src = (dst + 0.5) / s - 0.5 clamped at 0, taps clamped to the edge.
 *   2. grid = (w + fx, h + fy) normalised by W, H (flow.py:33-34) and sampled
 *      with align_corners=True, padding_mode='border' (flow.py:39-40), i.e.
 *      px = clamp((w+fx)/W*2-1 -> ((g+1)/2)*(W-1), 0, W-1).
 * x: [N,C,H,W]; flow: [N,2,Hf,Wf]; out: [N,C,H,W].
 */
static void resize_src(int dst, float inv_scale, int in_size, int *i0, int *i1, float *l1)
{
    float src = ((float)dst + 0.5f) * inv_scale - 0.5f;
    if (src < 0.0f) src = 0.0f;
    int a = (int)src;
    if (a > in_size - 1) a = in_size - 1;
    int b = a + (a < in_size - 1 ? 1 : 0);
    *i0 = a; *i1 = b; *l1 = src - (float)a;
}

void oracle_flow_warp(const float *x, const float *flow, float *out,
                      int N, int C, int H, int W, int Hf, int Wf)
{
    double sd = (double)W / (double)Wf; /* scale_factor (python float), flow.py:17 */
    float s = (float)sd;                /* flow * scale_factor, flow.py:20 */
    float inv = (float)(1.0 / sd);      /* ATen: static_cast<float>(1.0 / scale_factor) */
#pragma omp parallel for schedule(static)
    for (int nh = 0; nh < N * H; ++nh) {
        int n = nh / H, h = nh % H;
        int y0, y1; float ly;
        resize_src(h, inv, Hf, &y0, &y1, &ly);
        for (int w = 0; w < W; ++w) {
            int x0, x1; float lx;
            resize_src(w, inv, Wf, &x0, &x1, &lx);
            float f[2];
            for (int ch = 0; ch < 2; ++ch) {
                const float *p = flow + ((size_t)n * 2 + ch) * Hf * Wf;
                float v = (1.0f - ly) * ((1.0f - lx) * p[y0 * Wf + x0] + lx * p[y0 * Wf + x1]) +
                          ly * ((1.0f - lx) * p[y1 * Wf + x0] + lx * p[y1 * Wf + x1]);
                f[ch] = v * s;
            }
            float gx = ((float)w + f[0]) / (float)W * 2.0f - 1.0f;
            float gy = ((float)h + f[1]) / (float)H * 2.0f - 1.0f;
            float px = (gx + 1.0f) / 2.0f * (float)(W - 1);
            float py = (gy + 1.0f) / 2.0f * (float)(H - 1);
            px = fminf(fmaxf(px, 0.0f), (float)(W - 1));
            py = fminf(fmaxf(py, 0.0f), (float)(H - 1));
            int ix0 = (int)floorf(px), iy0 = (int)floorf(py);
            float tx = px - (float)ix0, ty = py - (float)iy0;
            int ix1 = ix0 + 1, iy1 = iy0 + 1;
            /* out-of-range taps carry zero weight after the border clamp */
            int vx1 = ix1 <= W - 1, vy1 = iy1 <= H - 1;
            if (!vx1) ix1 = W - 1;
            if (!vy1) iy1 = H - 1;
            float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
            float w10 = (1.0f - tx) * ty, w11 = tx * ty;
            for (int c = 0; c < C; ++c) {
                const float *p = x + ((size_t)n * C + c) * H * W;
                float v = p[iy0 * W + ix0] * w00;
                if (vx1) v += p[iy0 * W + ix1] * w01;
                if (vy1) v += p[iy1 * W + ix0] * w10;
                if (vx1 && vy1) v += p[iy1 * W + ix1] * w11;
                out[(((size_t)n * C + c) * H + h) * W + w] = v;
            }
        }
    }
}

int oracle_version(void) { return 1; }
