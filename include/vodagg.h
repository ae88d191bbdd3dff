/*
 * vodagg.h -- C ABI of libvodagg.so: hand-written sm_100a CUDA kernels for the
 * multi-frame feature-aggregation hot path of the MMTracking / MMDetection
 * video-object-detection stack (SELSA, TemporalRoIAlign, FGFA/DFF warp,
 * RoIAlign, batched NMS).
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in _host.  No torch types, no allocation inside the
 *     library, no retained pointers, no host synchronisation: every entry
 *     point only enqueues work on `stream` and is CUDA-graph capturable.
 *   - workspace is caller-provided; ask vod_*_workspace_bytes() first.
 *   - return value: 0 = ok; <0 = error (VOD_E_*).  vod_last_error() gives a
 *     thread-local message.  Nothing throws or aborts.
 *   - re-entrant: no global mutable state except lazily created, immutable
 *     function attributes / driver entry points.
 *
 * "replaces" lines cite the reference interface each entry point sits behind
 * (paths relative to the reference root).
 */
#ifndef VODAGG_H_
#define VODAGG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *vod_stream_t; /* cudaStream_t */

#define VOD_OK 0
#define VOD_E_BADARG (-1)      /* null pointer / bad size / unsupported shape */
#define VOD_E_LAUNCH (-2)      /* cudaGetLastError() after a launch */
#define VOD_E_WORKSPACE (-3)   /* workspace too small */
#define VOD_E_UNSUPPORTED (-4) /* shape not supported by this kernel variant */

#define VOD_DTYPE_F32 0
#define VOD_DTYPE_BF16 1

int vod_version(void);
const char *vod_last_error(void);
/* 1 when the visible device is compute capability 10.x (tcgen05 / TMEM / TMA paths usable). */
int vod_device_is_sm100(void);
/* Total number of CUDA kernels this library has launched in the process (monotonic; for benchmarks). */
long long vod_kernel_launch_count(void);

/* ------------------------------------------------------------------ layout
 * NCHW fp32 -> NHWC fp32 (+ optional per-pixel ||x||_2 over C and optional
 * L2-normalised bf16 copy [B*H*W, C]).  Feeds (1) and (4).
 * replaces: the .permute(...).contiguous() / x / x.norm(p=2, dim=1) passes at
 *   mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:127-140,159-161
 */
int vod_nchw_to_nhwc(const float *in_nchw, float *out_nhwc, float *norm_out /*nullable*/,
                     void *out_unit_bf16 /*nullable*/, int B, int C, int H, int W,
                     vod_stream_t stream);
/* rows [R, C] fp32 -> ||row||_2, unit-norm bf16 rows (either output nullable). */
int vod_rows_l2norm(const float *rows, float *norm_out, void *out_unit_bf16, int R, int C,
                    vod_stream_t stream);

/* -------------------------------------------------------------- (1) RoIAlign
 * Aligned/legacy average RoIAlign over NHWC features; one launch covers the
 * key frame plus all reference frames (rois[:,0] is the frame index).
 *   feat_nhwc [B,H,W,C] fp32; rois [K,5] = (b, x1, y1, x2, y2) image px;
 *   out_layout 0: out [K,C,ph,pw] (the reference's layout); 1: out [K,ph,pw,C].
 * replaces: mmcv.ops.RoIAlign.forward as constructed at
 *   mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55 and called at
 *   mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:72-75
 */
int vod_roi_align_fwd(const float *feat_nhwc, const float *rois, float *out, int B, int C, int H,
                      int W, int K, int ph, int pw, float spatial_scale, int sampling_ratio,
                      int aligned, int out_layout, vod_stream_t stream);

/* ------------------------------------------------- (2) flow warp / FGFA weights
 * x [N,C,H,W], flow [N,2,Hf,Wf] (channel 0 = dx, 1 = dy, flow-image px) -> out [N,C,H,W].
 * replaces: mmtrack.core.flow_warp_feats, mmtracking/mmtrack/core/motion/flow.py:4-41
 */
int vod_flow_warp(const float *x, const float *flow, float *out, int N, int C, int H, int W, int Hf,
                  int Wf, vod_stream_t stream);
/* Same with x [x_frames,C,H,W], x_frames = N or 1: with 1 every flow warps the SAME map in one launch -- the non-key frames
 * of a DFF key interval all warp the key frame's features (mmtracking/mmtrack/models/vid/dff.py:210-216 runs them one frame
 * at a time). */
int vod_flow_warp_shared(const float *x, const float *flow, float *out, int N, int x_frames, int C, int H,
                         int W, int Hf, int Wf, vod_stream_t stream);
/* Warp driven by the flow network's LOW-RESOLUTION prediction flow_lr [N,2,hl,wl]: the full-resolution flow
 *   mult2 * (mult1 * interpolate(flow_lr, scale_factor=up_scale, bilinear, align_corners=False))   [N,2,Hf,Wf]
 * that FlowNetSimple materialises (mmtracking/mmtrack/models/motion/flownet_simple.py:229-236: up_scale = mult1 =
 * 4 / img_scale_factor, mult2 = flow_scale_factor) is evaluated on the fly at the 4 of every 256 values flow_warp_feats reads,
 * instead of being written and re-read.  Same result as vod_flow_warp_shared on the materialised flow (same arithmetic).
 * replaces: the tail of FlowNetSimple.forward + mmtrack.core.flow_warp_feats (core/motion/flow.py:4-41). */
int vod_flow_warp_lowres(const float *x, const float *flow_lr, float *out, int N, int x_frames, int C, int H,
                         int W, int hl, int wl, int Hf, int Wf, double up_scale, float mult1, float mult2,
                         vod_stream_t stream);
/* key_emb [1,C,H,W], ref_emb [T,C,H,W], ref_x [T,Cx,H,W] -> out [1,Cx,H,W]:
 * cosine(key_emb, ref_emb[t]) over C, softmax over t, weighted sum of ref_x.
 * replaces: the weighting half of EmbedAggregator.forward,
 *   mmtracking/mmtrack/models/aggregators/embed_aggregator.py:71-81
 * ws (nullable): >= T*HW*4 bytes of scratch for the per-pixel cosines; with it the op runs as two
 *   machine-filling kernels (cosine per (frame, pixel), then softmax + weighted sum), without it as one
 *   smaller-grid kernel.
 */
int vod_embed_weighted_sum(const float *key_emb, const float *ref_emb, const float *ref_x,
                           float *out, int T, int C, int Cx, int HW, void *ws, size_t ws_bytes,
                           vod_stream_t stream);
/* Same weighting, but the weighted operand is re-warped on the fly from the raw
 * feature memory + flows (the warped tensor is never re-read); slot `key_slot`
 * (or -1) uses key_x un-warped, as FGFA does at mmtracking/mmtrack/models/vid/fgfa.py:277-282.
 * ws (nullable): >= T*H*W*4 bytes of scratch, as for vod_embed_weighted_sum: with it the op runs as two machine-filling
 *   kernels (cosines, then softmax + weighted sum of the on-the-fly warp).
 */
int vod_fgfa_warp_weighted_sum(const float *key_emb, const float *ref_emb, const float *raw_x,
                               const float *flow, const float *key_x, int key_slot, float *out,
                               int T, int C, int Cx, int H, int W, int Hf, int Wf, void *ws,
                               size_t ws_bytes, vod_stream_t stream);

/* ------------------------------------------------------ (3) SELSA aggregation
 * Per-head softmax(Q K^T * scale) V.  q [N, heads*d], k [M, heads*d] row-major
 * (head h = columns h*d..h*d+d-1) -> out [N, heads*d] fp32.
 *   v_layout 0: v [M, heads*d] row-major (ldv ignored)
 *   v_layout 1: v is V^T [heads*d, ldv] row-major, ldv >= M (what a projection GEMM can emit
 *               directly; the tensor-core path consumes it without a transposition pass)
 * dtype: VOD_DTYPE_F32 (tf32 tensor-core math when d == 64 on sm_100, otherwise fp32 SIMT)
 *        or VOD_DTYPE_BF16 inputs (bf16 tensor-core math, fp32 accumulate/softmax).
 * impl: 0 auto, 1 fp32 SIMT, 2 tcgen05 (error if unsupported shape/device).
 * replaces: the bmm/softmax/bmm core of SelsaAggregator.forward,
 *   mmtracking/mmtrack/models/aggregators/selsa_aggregator.py:51-70
 */
size_t vod_selsa_attn_workspace_bytes(int N, int M, int heads, int d);
int vod_selsa_attn(const void *q, const void *k, const void *v, float *out, int N, int M, int heads,
                   int d, float scale, int dtype, int v_layout, int ldv, int impl, void *ws,
                   size_t ws_bytes, vod_stream_t stream);

/* The element-wise tail of one SelsaBBoxHead layer, in place and in one launch:
 *   x[r, c] = relu(x[r, c] + y[r, c] + bias[c])  for r < rows  (x [rows, cols] fp32, y the aggregator's
 *             output WITHOUT its bias, bias [cols] = the aggregator's fused output bias)
 *   ref[i]  = relu(ref[i])                        for i < ref_elems (nullable when ref_elems == 0)
 * cols and ref_elems must be multiples of 4, pointers 16-byte aligned.
 * replaces: `x = x + self.aggregator[i](x, ref_x); ref_x = self.relu(ref_x); x = self.relu(x)`,
 *   mmtracking/mmtrack/models/roi_heads/bbox_heads/selsa_bbox_head.py:56-58 (and the `+ bias` of
 *   the aggregator's last linear, mmtracking/mmtrack/models/aggregators/selsa_aggregator.py:72)
 */
int vod_selsa_residual_relu(float *x, const float *y, const float *bias, int rows, int cols, float *ref,
                            long ref_elems, vod_stream_t stream);

/* ------------------------------- (4) TemporalRoIAlign: most-similar sampling
 * roi_feats [N*P, C] fp32 (P = ph*pw bins, NHWC-style rows), ref_nhwc [T, HW, C] fp32.
 * For every (row, frame): cosine similarity against all HW locations, top-k
 * (k <= 4), softmax over the k values, weighted sum of the raw ref features.
 *   out     [T, N*P, C] fp32  (row-major; the host views it as [T,N,ph,pw,C])
 *   idx_out [N*P, T, k] int32 flat h*W+w locations, descending similarity (nullable)
 *   val_out [N*P, T, k] fp32 similarities (nullable)
 * impl: 0 auto, 1 exact fp32 SIMT scan, 2 bf16 tcgen05 candidate GEMM + exact fp32 re-score; a (row, frame) whose
 *       candidate lists could have dropped a member of the exact top-k (more than 4 near-tied locations congruent mod 4) is
 *       detected and re-scanned in exact fp32, so the selected locations are the exact top-k for every input.
 * replaces: TemporalRoIAlign.most_similar_roi_align,
 *   mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:99-181
 */
size_t vod_msra_workspace_bytes(int NP, int C, int T, int HW, int k);
/* ref_norm [T*HW] fp32 and ref_unit_bf16 [T*HW, C] may be passed when the caller already has them
 * from vod_nchw_to_nhwc (nullable: recomputed into the workspace).  When HW % 4 != 0 the ref_unit_bf16 buffer
 * must be readable for 3 rows (3*C*2 bytes) past row T*HW (contents ignored): the tensor-core pass fetches
 * locations in groups of 4. */
int vod_msra_topk_sample(const float *roi_feats, const float *ref_nhwc, const float *ref_norm,
                         const void *ref_unit_bf16, float *out, int *idx_out, float *val_out, int NP,
                         int C, int T, int HW, int k, int impl, void *ws, size_t ws_bytes,
                         vod_stream_t stream);

/* Byte offset, inside the workspace of vod_msra_topk_sample, of an int32 counter block: [0] = number of (row, frame) pairs of
 * the last tensor-core call whose candidate lists may have lost a member of the exact top-k and were therefore re-scanned in
 * exact fp32 (the result is exact either way; the counter is a diagnostic for tests and benchmarks). */
size_t vod_msra_overflow_counter_offset(int NP, int C, int T, int HW);

/* The tensor-core half of (4) alone (what vod_msra_topk_sample runs before its fp32 re-score): unit-norm bf16
 * rows roi_unit [NP, C] x ref_unit [T*HW, C] -> cand_out [NP, T, 16] packed keys
 * round((1.5 + similarity) * 2^11) << 12 | location: the 4 best of each of the four 32-column
 * groups of the 128-location tiles; 0 = empty slot.  Exposed so the GEMM can be profiled / roofline-timed on
 * its own and its candidate recall tested.
  * ref_unit: same 3-row readable padding as above when HW % 4 != 0.
 */
int vod_msra_gemm_candidates(const void *roi_unit_bf16, const void *ref_unit_bf16, uint32_t *cand_out,
                             int NP, int C, int T, int HW, vod_stream_t stream);

/* ------------------------- (4') TemporalRoIAlign: temporal attention weighting
 * x_all, emb_all [T1, N, P, C] fp32 (frame 0 = the key's own RoI features);
 * emb_bias [C] (nullable) is added to every embedding vector on load, so the caller's embed conv
 * can run without its bias and no separate bias-add pass touches emb_all;
 * per (n, bin, head) dot(emb[t], emb[0]) over C/heads channels / sqrt(C/heads),
 * softmax over t, out = sum_t w * x_all[t].
 *   out_layout 0: out [N, C, P] (== [N,C,ph,pw]); 1: out [N, P, C].
 *   heads <= 0: plain mean over T1 (temporal_roi_align.py:203-206); emb_all may be null.
 * replaces: the weighting half of temporal_attentional_feature_aggregation,
 *   mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:77-97
 */
int vod_tafa_weighted_sum(const float *x_all, const float *emb_all, const float *emb_bias, float *out,
                          int T1, int N, int P, int C, int heads, int out_layout, vod_stream_t stream);

/* Key-projected attention logits: the same weighting WITHOUT embedding the T reference slots.
 * The embed conv is linear, so <conv(x_t), ek>_head = sum_{tap,c} x_t[p+tap,c] * G[h,(n,p),tap,c] with
 * G = ek_head . W_head (one batched library GEMM of the KEY embedding ek = conv(x_0)+b against the conv
 * weight); the bias term is constant over t and cancels in the softmax.  Only the N key patches go
 * through the conv instead of (T+1)*N.
 *   vod_tafa_keyproj_chunk: channel-chunk width CC the kernel wants G laid out with, 0 = unsupported
 *     shape (heads != 4, C % 32 != 0, or the [T1,P,CC] tile exceeds shared memory) -> use
 *     vod_tafa_weighted_sum with full embeddings instead.
 *   x_all [T1, N, P, C] fp32;  G [heads, N*P, C/CC, 9, CC] (tap = ky*3+kx of the 3x3, pad-1 conv), g_dtype VOD_DTYPE_F32 or
 *   VOD_DTYPE_BF16 (G is 70 % of this kernel's bytes; a bf16 G -- e.g. the output of a bf16 library GEMM -- is unpacked to fp32
 *   on the fly, everything else stays fp32);
 *   parts [C/CC, N, P, heads, T1] fp32 out: per-chunk partial logits (unscaled).
 *   vod_tafa_weighted_sum_logits: sums the chunks, scales by 1/sqrt(C/heads), softmax over t, weighted sum.
 * replaces: temporal_roi_align.py:72-97 (embed_network over img_n*roi_n patches + multi-head weighting)
 */
int vod_tafa_keyproj_chunk(int T1, int P, int C, int heads);
int vod_tafa_keyproj_logits(const float *x_all, const void *G, int g_dtype, float *parts, int T1, int N, int ph,
                            int pw, int C, int heads, int cc, vod_stream_t stream);
int vod_tafa_weighted_sum_logits(const float *x_all, const float *logit_parts, int nparts, float *out,
                                 int T1, int N, int P, int C, int heads, int out_layout, vod_stream_t stream);

/* -------------------------------------------------------- (5) batched NMS
 * Bitmask NMS with the sort, mask and the greedy sweep all on the device (no
 * host round trip).  Boxes of `n_images` independent images are concatenated;
 * seg_offsets_host [n_images+1] gives each image's [begin, end) range.
 *   boxes [n,4] fp32 (x1,y1,x2,y2), scores [n] fp32, labels [n] int64 (nullable)
 *   mode 0: class-agnostic (labels ignored)
 *   mode 1: mmcv coordinate-offset trick: boxes + label * (max(boxes of the image) + 1)
 *   mode 2: class-aware on raw coordinates (mmcv's per-class split path, n >= split_thr)
 *   keep_out  [n] int64: per image, kept indices relative to the image's first box, in
 *             descending-score order, written at keep_out[seg_begin ...]
 *   num_keep_out [n_images] int32
 *   max_keep <= 0: unlimited; else the sweep stops after max_keep survivors per image.
 * replaces: mmcv.ops.nms.batched_nms / mmcv.ops.nms (mmcv-full 1.2.x, un-vendored) as called at
 *   mmdetection/mmdet/core/post_processing/bbox_nms.py:84 and
 *   mmdetection/mmdet/models/dense_heads/rpn_head.py:233-235
 */
size_t vod_nms_workspace_bytes(int n_total, int max_seg);
int vod_batched_nms(const float *boxes, const float *scores, const int64_t *labels, int n_total,
                    const int *seg_offsets_host, int n_images, float iou_thr, int mode, int max_keep,
                    int64_t *keep_out, int *num_keep_out, void *ws, size_t ws_bytes,
                    vod_stream_t stream);

/* Superset of vod_batched_nms for the fixed-shape, sync-free multiclass path:
 *   n_valid_dev [n_images] (device, nullable): boxes whose score is -inf are invalid, the image's effective
 *               candidate count is n_valid_dev[img] (they sort last and are never visited)
 *   mode 3:     decided on the device per image: coordinate-offset trick when n_valid < split_thr, per-class
 *               raw-coordinate NMS otherwise (what mmcv's batched_nms does on the host)
 *   dets_out [n_images][max_keep][5], labels_out [n_images][max_keep] (nullable): survivors' (box, score) and
 *               label gathered by the sweep itself; keep_out may then be null.
 */
int vod_batched_nms_ex(const float *boxes, const float *scores, const int64_t *labels, int n_total,
                       const int *seg_offsets_host, int n_images, float iou_thr, int mode, int max_keep,
                       const int *n_valid_dev, int split_thr, int64_t *keep_out, int *num_keep_out,
                       float *dets_out, int64_t *labels_out, void *ws, size_t ws_bytes,
                       vod_stream_t stream);

/* ---------------------------------------------- get_bboxes front half (caller of (5))
 * softmax(cls_score) + delta2bbox (+ clip to img_h/img_w when >= 0, + division by scale_factor_host when
 * non-null) + the multiclass candidate expansion, in one launch: candidate id = proposal * ncls + class,
 * score = -inf when score <= score_thr, *n_valid_dev = number of valid candidates.
 * replaces: the elementwise half of BBoxHead.get_bboxes (mmdet/models/roi_heads/bbox_heads/bbox_head.py:319-353),
 *   delta2bbox (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:134-237) and the candidate filtering of
 *   multiclass_nms (mmdet/core/post_processing/bbox_nms.py:34-73)
 */
int vod_bbox_decode_candidates(const float *rois, const float *cls_score, const float *bbox_pred, int N,
                               int ncls, int reg_class_agnostic, const float *means_host,
                               const float *stds_host, float max_ratio, float img_h, float img_w,
                               const float *scale_factor_host, float score_thr, float *cand_boxes,
                               float *cand_scores, int64_t *cand_labels, int *n_valid_dev,
                               vod_stream_t stream);

/* RPN proposal decode, RPNHead._get_bboxes (mmdetection/mmdet/models/dense_heads/rpn_head.py:163-188, one feature
 * level): topk_idx [B,K] int64 = positions of the K best-scoring anchors of each image (from the score sort),
 * deltas [B,A,4], anchors [A,4] -> boxes [B*K,4] = delta2bbox(anchors[idx], deltas[idx], means 0, stds 1) clipped to
 * the image (img_w < 0: no clipping).  Feeds vod_batched_nms with B segments of K boxes. */
int vod_rpn_decode_topk(const int64_t *topk_idx, const float *deltas, const float *anchors, float *boxes, int B, int K,
                        int A, float max_ratio, float img_h, float img_w, vod_stream_t stream);

/* ------------------------------------------------------------ temporal attention fusion (the fork's Denoising2Aggergator)
 * Sampling half of mmcv's modulated_deform_conv2d (DCNv2) as used by ModulatedDCNPack.forward
 *   (mmtracking/mmtrack/models/aggregators/denoising2_aggregator.py:72-82): x [B,H,W,C] channels-last; p (and optional q,
 *   added to it) = raw conv_offset outputs, channels-last [B or 1, Ho, Wo, 3*G*kh*kw] -- channels [0, 2GK) are the offsets
 *   (deformable-group major, (dy, dx) interleaved per tap: chunk + cat of :76-77 leave them in place), [2GK, 3GK) the mask
 *   logits (sigmoid applied here, :78).  p_shared / q_shared != 0: that map is one [Ho,Wo,3GK] map broadcast over B.
 *   col [B*Ho*Wo, kh*kw*C] (column k*C + c) = mask * bilinear(x, tap position + offset), zero outside the map; the caller
 *   multiplies by the [Cout, kh*kw*C] weight (groups == 1).  128-bit accesses when C / G is a multiple of 4, scalar otherwise. */
int vod_mdcn_im2col(const float *x, const float *p, const float *q, float *col, int B, int C, int H, int W, int G,
                    int kh, int kw, int stride, int pad, int dil, int p_shared, int q_shared, vod_stream_t stream);
/* cor [I,T,E], x [T,E] -> out [I,E]: out[i] = sum_t softmax_t(cor[i,:,e])[t] * x[t,e]  (E % 4 == 0; any common element order).
 * replaces: torch.softmax(x_cor, dim=0) and torch.sum(x_cor * x, dim=0) of TemporalAttentionFusion.forward (:145-146). */
int vod_temporal_softmax_fuse(const float *cor, const float *x, float *out, int I, int T, long E, vod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VODAGG_H_ */
