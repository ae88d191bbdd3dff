// decode.cu -- the elementwise front half of BBoxHead.get_bboxes + multiclass_nms as ONE kernel, producing the
// fixed-shape candidate arrays the device NMS consumes (no nonzero / index / host sync):
//   scores = softmax(cls_score)                                   mmdet/models/roi_heads/bbox_heads/bbox_head.py:319-320
//   bboxes = delta2bbox(rois[:, 1:], bbox_pred, means, stds, max_shape)   mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:134-237
//   bboxes /= scale_factor (rescale)                              bbox_head.py:345-353
//   candidates (proposal p, class c): box, score, label; valid iff score > score_thr   mmdet/core/post_processing/bbox_nms.py:34-73
// Invalid candidates get score = -inf (they sort last in the NMS rank sort) and n_valid counts the valid ones.
// Every arithmetic step is a separately rounded fp32 operation in the reference (one torch elementwise kernel
// each), so the kernel uses __fmul_rn/__fadd_rn (no FMA contraction) and the same expf / IEEE division.
#include "common.cuh"

namespace vod {

struct DecodeParams {
    const float *rois;       // [N,5]
    const float *cls_score;  // [N, ncls+1]
    const float *bbox_pred;  // [N, 4*ncls] (or [N,4] when class-agnostic)
    float *cand_boxes;       // [N*ncls, 4]
    float *cand_scores;      // [N*ncls]
    int64_t *cand_labels;    // [N*ncls]
    int *n_valid;            // [1]
    int N, ncls, agnostic;
    float means[4], stds[4];
    float max_ratio;         // |log(wh_ratio_clip)|
    float img_h, img_w;      // clip bounds; < 0: no clipping
    float inv_scale[4];      // unused (division is done with the scale factors below)
    float scale[4];          // rescale divisors; scale[0] <= 0: no rescale
    float score_thr;
};

constexpr int kDecWarps = 4;

__global__ void __launch_bounds__(kDecWarps * 32)
bbox_decode_kernel(const DecodeParams p) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kDecWarps + (threadIdx.x >> 5);
    if (row >= p.N) return;
    const int nc1 = p.ncls + 1;
    const float *cs = p.cls_score + (size_t)row * nc1;
    // softmax over ncls+1 logits: max, sum of exp(x - max), exp(x - max) / sum
    float mx = -INFINITY;
    for (int c = lane; c < nc1; c += 32) mx = fmaxf(mx, cs[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < nc1; c += 32) sum += expf(cs[c] - mx);
    sum = warp_sum(sum);

    const float *r = p.rois + (size_t)row * 5;
    const float x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
    const float px = __fmul_rn(__fadd_rn(x1, x2), 0.5f), py = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
    const float pw = __fsub_rn(x2, x1), ph = __fsub_rn(y2, y1);
    int valid = 0;
    for (int c = lane; c < p.ncls; c += 32) {
        const float score = __fdiv_rn(expf(cs[c] - mx), sum);
        const float *d = p.bbox_pred + (size_t)row * (p.agnostic ? 4 : 4 * p.ncls) + (p.agnostic ? 0 : 4 * c);
        const float dx = __fadd_rn(__fmul_rn(d[0], p.stds[0]), p.means[0]);
        const float dy = __fadd_rn(__fmul_rn(d[1], p.stds[1]), p.means[1]);
        float dw = __fadd_rn(__fmul_rn(d[2], p.stds[2]), p.means[2]);
        float dh = __fadd_rn(__fmul_rn(d[3], p.stds[3]), p.means[3]);
        dw = fminf(fmaxf(dw, -p.max_ratio), p.max_ratio);
        dh = fminf(fmaxf(dh, -p.max_ratio), p.max_ratio);
        const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
        const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
        float bx1 = __fsub_rn(gx, __fmul_rn(gw, 0.5f)), by1 = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
        float bx2 = __fadd_rn(gx, __fmul_rn(gw, 0.5f)), by2 = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
        if (p.img_w >= 0.f) {
            bx1 = fminf(fmaxf(bx1, 0.f), p.img_w); bx2 = fminf(fmaxf(bx2, 0.f), p.img_w);
            by1 = fminf(fmaxf(by1, 0.f), p.img_h); by2 = fminf(fmaxf(by2, 0.f), p.img_h);
        }
        if (p.scale[0] > 0.f) {
            bx1 = __fdiv_rn(bx1, p.scale[0]); by1 = __fdiv_rn(by1, p.scale[1]);
            bx2 = __fdiv_rn(bx2, p.scale[2]); by2 = __fdiv_rn(by2, p.scale[3]);
        }
        const size_t id = (size_t)row * p.ncls + c;
        reinterpret_cast<float4 *>(p.cand_boxes)[id] = make_float4(bx1, by1, bx2, by2);
        const bool ok = score > p.score_thr;
        p.cand_scores[id] = ok ? score : -INFINITY;
        p.cand_labels[id] = c;
        valid += ok;
    }
    valid = (int)warp_sum((float)valid);
    if (lane == 0 && valid) atomicAdd(p.n_valid, valid);
}

// RPN proposal decode (RPNHead._get_bboxes, mmdet/models/dense_heads/rpn_head.py:163-188 for one level): the anchors and
// deltas of the nms_pre best-scoring positions (indices from the score sort) -> clipped proposal boxes, one launch for
// all images.  Same separately-rounded arithmetic as delta2bbox with means 0 / stds 1.
__global__ void __launch_bounds__(256)
rpn_decode_kernel(const int64_t *__restrict__ topk_idx /*[B,K]*/, const float *__restrict__ deltas /*[B,A,4]*/,
                  const float *__restrict__ anchors /*[A,4]*/, float *__restrict__ boxes /*[B*K,4]*/, int B, int K, int A,
                  float max_ratio, float img_h, float img_w) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long)B * K) return;
    const int b = (int)(i / K);
    const long a = topk_idx[i];
    const float4 an = reinterpret_cast<const float4 *>(anchors)[a];
    const float4 d = reinterpret_cast<const float4 *>(deltas)[(size_t)b * A + a];
    const float px = __fmul_rn(__fadd_rn(an.x, an.z), 0.5f), py = __fmul_rn(__fadd_rn(an.y, an.w), 0.5f);
    const float pw = __fsub_rn(an.z, an.x), ph = __fsub_rn(an.w, an.y);
    // deltas * stds + means with stds = 1, means = 0: x * 1 + 0 is exact, nothing to round
    const float dw = fminf(fmaxf(d.z, -max_ratio), max_ratio), dh = fminf(fmaxf(d.w, -max_ratio), max_ratio);
    const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
    const float gx = __fadd_rn(px, __fmul_rn(pw, d.x)), gy = __fadd_rn(py, __fmul_rn(ph, d.y));
    float x1 = __fsub_rn(gx, __fmul_rn(gw, 0.5f)), y1 = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
    float x2 = __fadd_rn(gx, __fmul_rn(gw, 0.5f)), y2 = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
    if (img_w >= 0.f) {
        x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
        y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
    }
    reinterpret_cast<float4 *>(boxes)[i] = make_float4(x1, y1, x2, y2);
}

}  // namespace vod

using namespace vod;

extern "C" int vod_bbox_decode_candidates(const float *rois, const float *cls_score, const float *bbox_pred, int N,
                                          int ncls, int reg_class_agnostic, const float *means_host,
                                          const float *stds_host, float max_ratio, float img_h, float img_w,
                                          const float *scale_factor_host, float score_thr, float *cand_boxes,
                                          float *cand_scores, int64_t *cand_labels, int *n_valid_dev,
                                          vod_stream_t stream) {
    VOD_REQUIRE(n_valid_dev, "vod_bbox_decode_candidates: null n_valid");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(n_valid_dev, 0, sizeof(int), st);
    if (N == 0) return check_launch("vod_bbox_decode_candidates(memset)");
    VOD_REQUIRE(rois && cls_score && bbox_pred && cand_boxes && cand_scores && cand_labels && means_host && stds_host,
                "vod_bbox_decode_candidates: null pointer");
    VOD_REQUIRE(N > 0 && ncls > 0, "vod_bbox_decode_candidates: bad dims");
    VOD_REQUIRE((reinterpret_cast<uintptr_t>(cand_boxes) & 15) == 0, "vod_bbox_decode_candidates: cand_boxes must be 16-byte aligned");
    DecodeParams p;
    p.rois = rois; p.cls_score = cls_score; p.bbox_pred = bbox_pred;
    p.cand_boxes = cand_boxes; p.cand_scores = cand_scores; p.cand_labels = cand_labels; p.n_valid = n_valid_dev;
    p.N = N; p.ncls = ncls; p.agnostic = reg_class_agnostic;
    for (int i = 0; i < 4; ++i) {
        p.means[i] = means_host[i]; p.stds[i] = stds_host[i];
        p.scale[i] = scale_factor_host ? scale_factor_host[i] : 0.f;
        p.inv_scale[i] = 0.f;
    }
    p.max_ratio = max_ratio; p.img_h = img_h; p.img_w = img_w; p.score_thr = score_thr;
    bbox_decode_kernel<<<ceil_div(N, kDecWarps), kDecWarps * 32, 0, st>>>(p); note_launch();
    return check_launch("vod_bbox_decode_candidates");
}

extern "C" int vod_rpn_decode_topk(const int64_t *topk_idx, const float *deltas, const float *anchors, float *boxes, int B,
                                   int K, int A, float max_ratio, float img_h, float img_w, vod_stream_t stream) {
    if (B == 0 || K == 0) return VOD_OK;
    VOD_REQUIRE(topk_idx && deltas && anchors && boxes, "vod_rpn_decode_topk: null pointer");
    VOD_REQUIRE(B > 0 && K > 0 && A >= K, "vod_rpn_decode_topk: bad dims (B=%d K=%d A=%d)", B, K, A);
    VOD_REQUIRE(((reinterpret_cast<uintptr_t>(deltas) | reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(boxes)) & 15) == 0,
                "vod_rpn_decode_topk: deltas / anchors / boxes must be 16-byte aligned");
    rpn_decode_kernel<<<(unsigned)ceil_div((long)B * K, 256L), 256, 0, as_stream(stream)>>>(topk_idx, deltas, anchors, boxes, B, K, A,
                                                                                           max_ratio, img_h, img_w);
    note_launch();
    return check_launch("vod_rpn_decode_topk");
}
