"""Host-side mirror of the mmcv symbols the hot path binds, plus thin functional
wrappers over the C ABI (include/vodagg.h).  PyTorch is used only for device
memory and streams; every op below launches hand-written sm_100a kernels from
libvodagg.so and raises if the library or a CUDA device is missing.

Mirrors (reference call sites, paths relative to the reference root):
  RoIAlign      mmcv.ops.RoIAlign, built at
                mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55
  nms           mmcv.ops.nms
  batched_nms   mmcv.ops.nms.batched_nms, called at
                mmdetection/mmdet/core/post_processing/bbox_nms.py:84 and
                mmdetection/mmdet/models/dense_heads/rpn_head.py:233-235
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib

_ws = _lib.Workspace()


def _pair(x):
    return (int(x), int(x)) if isinstance(x, int) else (int(x[0]), int(x[1]))


def _f32c(t):
    """fp32 + contiguous view of a tensor (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------- layout
# Two-entry memo: SelsaRoIHead._bbox_forward hands the SAME reference feature tensor to the extractor twice (key call with
# ref_feats=, then the reference-RoI call) with the key map's own layout pass in between; the second NHWC transposition of
# the reference maps is skipped.  An entry keeps the source tensor alive and is validated by identity + in-place version
# counter (writes through raw pointers are invisible to it: CUDA-graph capture clears the memo, see heads.capture_callable).
_nhwc_memo = {}


def to_nhwc(x, want_norm=False, want_unit_bf16=False):
    for slot in ('a', 'b'):
        m = _nhwc_memo.get(slot)
        if m is not None and m[0] is x and m[1] == x._version and (m[3] is not None or not want_norm) \
                and (m[4] is not None or not want_unit_bf16):
            return m[2], m[3], m[4]
    res = _to_nhwc(x, want_norm, want_unit_bf16)
    _nhwc_memo['b'] = _nhwc_memo.get('a')
    _nhwc_memo['a'] = (x, x._version, res[0], res[1], res[2])
    return res


# ----------------------------------------------------------------------------- fork / join on a side stream
_side_streams = {}


class fork:
    """``with fork(device) as f: ...`` runs the body's launches on a per-device side stream that first waits for everything
    already queued on the current stream; ``f.join()`` makes the current stream wait for the side stream.  Independent
    branches of a step (the key-slot embed conv + G product next to the most-similar search, the reference-RoI branch next to
    the key branch) then overlap on the device -- also inside a captured CUDA graph, where the fork becomes two graph branches.

    Discipline that keeps the caching allocator safe without record_stream: side-stream work only ever happens inside a
    fork, and every fork starts with side.wait_stream(current), so a block freed by the consumer on the current stream is
    never reused on the side stream before that consumer has been queued."""

    def __init__(self, device):
        self.device = torch.device(device)
        key = (self.device.type, self.device.index)
        if key not in _side_streams:
            _side_streams[key] = torch.cuda.Stream(self.device)
        self.side = _side_streams[key]
        self.main = None
        self._ctx = None

    def __enter__(self):
        self.main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(self.main)
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        if exc[0] is not None:                  # do not leave a dangling branch behind an exception (graph capture needs the join)
            self.main.wait_stream(self.side)
        return False

    def join(self):
        self.main.wait_stream(self.side)


def _to_nhwc(x, want_norm=False, want_unit_bf16=False, out=None):
    """[B,C,H,W] fp32 -> ([B,H,W,C] fp32 contiguous, norm [B*H*W] | None, unit bf16 [B*H*W, C] | None).

    A channels_last input is consumed in place (zero copy) unless norms are requested.
    ``out`` = (nhwc, norm, unit) preallocated contiguous destinations (e.g. the slots of a reference-frame cache); the unit
    buffer must stay readable for 3 rows past its end when H*W % 4 != 0 (include/vodagg.h)."""
    _lib.require_cuda(x)
    assert x.dim() == 4
    x = x.float() if x.dtype != torch.float32 else x
    B, C, H, W = x.shape
    if out is not None:
        nhwc, norm, unit = out
        assert nhwc.is_contiguous() and nhwc.numel() == x.numel() and nhwc.dtype == torch.float32
        assert norm is None or (norm.is_contiguous() and norm.numel() == B * H * W)
        assert unit is None or (unit.is_contiguous() and unit.numel() == B * H * W * C and unit.dtype == torch.bfloat16)
        if B * H * W:
            if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
                nhwc.view(B, H, W, C).copy_(x.permute(0, 2, 3, 1))
                _lib.call('vod_rows_l2norm', _lib.ptr(nhwc), _lib.ptr(norm), _lib.ptr(unit), B * H * W, C,
                          _lib.stream_ptr(x.device))
            else:
                _lib.call('vod_nchw_to_nhwc', _lib.ptr(x.contiguous()), _lib.ptr(nhwc), _lib.ptr(norm), _lib.ptr(unit), B, C, H, W,
                          _lib.stream_ptr(x.device))
        return nhwc.view(B, H, W, C), norm, unit
    if not (want_norm or want_unit_bf16) and x.is_contiguous(memory_format=torch.channels_last) and B * C * H * W > 0:
        return x.permute(0, 2, 3, 1), None, None
    nhwc = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
    norm = torch.empty((B * H * W,), dtype=torch.float32, device=x.device) if (want_norm or want_unit_bf16) else None
    # 4 spare rows behind the unit copy: the msra GEMM's TMA box fetches locations in groups of 4 (vodagg.h)
    unit = torch.empty((B * H * W + 4, C), dtype=torch.bfloat16, device=x.device)[:B * H * W] if want_unit_bf16 else None
    if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        # already NHWC in memory: only the norms / unit copy are missing
        nhwc = x.permute(0, 2, 3, 1)
        if B * H * W:
            _lib.call('vod_rows_l2norm', _lib.ptr(nhwc), _lib.ptr(norm), _lib.ptr(unit), B * H * W, C,
                      _lib.stream_ptr(x.device))
        return nhwc, norm, unit
    x = x.contiguous()
    if B * H * W:
        _lib.call('vod_nchw_to_nhwc', _lib.ptr(x), _lib.ptr(nhwc), _lib.ptr(norm), _lib.ptr(unit), B, C, H, W,
                  _lib.stream_ptr(x.device))
    return nhwc, norm, unit


# ----------------------------------------------------------------------------- (1) RoIAlign
def roi_align_nhwc(feat_nhwc, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True,
                   out_nhwc=False, out=None):
    """feat_nhwc [B,H,W,C] contiguous fp32, rois [K,5] -> [K,C,ph,pw] (or [K,ph,pw,C] when out_nhwc).
    ``out``: optional preallocated contiguous fp32 destination of that many elements."""
    _lib.require_cuda(feat_nhwc, rois)
    ph, pw = _pair(output_size)
    B, H, W, C = feat_nhwc.shape
    rois = _f32c(rois)
    K = rois.shape[0]
    shape = (K, ph, pw, C) if out_nhwc else (K, C, ph, pw)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=feat_nhwc.device)
    else:
        assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == K * ph * pw * C
    if K:
        assert feat_nhwc.is_contiguous() and feat_nhwc.dtype == torch.float32
        _lib.call('vod_roi_align_fwd', _lib.ptr(feat_nhwc), _lib.ptr(rois), _lib.ptr(out), B, C, H, W, K, ph, pw,
                  float(spatial_scale), int(sampling_ratio), int(bool(aligned)), int(bool(out_nhwc)),
                  _lib.stream_ptr(feat_nhwc.device))
    return out


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg', aligned=True,
              channels_last_out=False):
    """Functional mmcv.ops.roi_align (NCHW in, [K,C,ph,pw] out).

    ``channels_last_out``: the result has the same logical shape [K,C,ph,pw] and values but channels_last
    strides (memory [K,ph,pw,C]); the kernel then writes straight from registers, without the shared-memory
    transposition tile.  Consumers that flatten in (C,ph,pw) order still get correct values (torch copies)."""
    if pool_mode != 'avg':
        raise NotImplementedError("vodagg RoIAlign implements pool_mode='avg' only (the reference's configs)")
    nhwc, _, _ = to_nhwc(input)
    if channels_last_out:
        return roi_align_nhwc(nhwc.contiguous(), rois, output_size, spatial_scale, sampling_ratio, aligned,
                              out_nhwc=True).permute(0, 3, 1, 2)
    return roi_align_nhwc(nhwc.contiguous(), rois, output_size, spatial_scale, sampling_ratio, aligned)


class RoIAlign(nn.Module):
    """Drop-in for mmcv.ops.RoIAlign (mmcv-full 1.2.x signature)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg', aligned=True,
                 use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision  # accepted for signature parity; the CUDA kernel is always used
        self.channels_last_out = False          # set by SelsaRoIHead: emit channels_last strides (same values)

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.pool_mode,
                         self.aligned, channels_last_out=self.channels_last_out)

    def __repr__(self):
        return ('%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s, pool_mode=%s, aligned=%s)' %
                (self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio, self.pool_mode,
                 self.aligned))


# ----------------------------------------------------------------------------- (5) NMS
NMS_MODE_AGNOSTIC, NMS_MODE_OFFSET, NMS_MODE_CLASS = 0, 1, 2


def nms_device(boxes, scores, labels, iou_threshold, mode, seg_offsets=None, max_keep=-1):
    """Raw device NMS.  Returns (keep [n] int64 padded, num_keep [n_images] int32), both on the device, no sync.

    keep holds, per image, indices relative to the image's first box in descending-score order."""
    _lib.require_cuda(boxes, scores, labels)
    boxes = _f32c(boxes)
    scores = _f32c(scores)
    n = boxes.shape[0]
    if seg_offsets is None:
        seg_offsets = [0, n]
    n_images = len(seg_offsets) - 1
    max_seg = max([seg_offsets[i + 1] - seg_offsets[i] for i in range(n_images)] + [0])
    dev = boxes.device
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    num_keep = torch.empty((n_images,), dtype=torch.int32, device=dev)
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    lib = _lib.load()
    ws_bytes = lib.vod_nms_workspace_bytes(n, max_seg)
    ws = _ws.get(ws_bytes, dev)
    offs = (ctypes.c_int * (n_images + 1))(*[int(o) for o in seg_offsets])
    _lib.call('vod_batched_nms', _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(labels), n, offs, n_images,
              float(iou_threshold), int(mode), int(max_keep), _lib.ptr(keep), _lib.ptr(num_keep), _lib.ptr(ws),
              ws.numel(), _lib.stream_ptr(dev))
    return keep, num_keep


def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    """Drop-in for mmcv.ops.nms: returns (dets [k,5], keep [k] int64), keep in descending-score order."""
    assert boxes.size(1) == 4 and boxes.size(0) == scores.size(0)
    if offset != 0:
        raise NotImplementedError('vodagg nms implements offset=0 (mmdet default)')
    if score_threshold > 0:
        valid = (scores > score_threshold).nonzero(as_tuple=False).squeeze(1)
        boxes_v, scores_v = boxes[valid], scores[valid]
    else:
        valid, boxes_v, scores_v = None, boxes, scores
    if boxes_v.shape[0] == 0:
        keep = torch.zeros((0,), dtype=torch.int64, device=boxes.device)
    else:
        keep_pad, num = nms_device(boxes_v, scores_v, None, iou_threshold, NMS_MODE_AGNOSTIC, max_keep=max_num)
        keep = keep_pad[:int(num.item())]      # the one host sync a variable-length result requires
    if valid is not None:
        keep = valid[keep]
    dets = torch.cat((boxes[keep].float(), scores[keep].float().reshape(-1, 1)), dim=1)
    return dets, keep


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """Drop-in for mmcv.ops.nms.batched_nms (mmcv-full 1.2.x semantics).

    Below ``split_thr`` boxes mmcv adds ``idx * (max_coordinate + 1)`` to every box and runs one NMS;
    at or above it, it runs NMS per class on the raw coordinates.  Both are reproduced on the device
    (mode 1 / mode 2 of vod_batched_nms).  ``nms_cfg['max_num']`` (mmcv >= 1.3) bounds the survivors."""
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    nms_type = nms_cfg_.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError("vodagg batched_nms implements type='nms' (got %r)" % nms_type)
    split_thr = nms_cfg_.pop('split_thr', 10000)
    thr = nms_cfg_.pop('iou_threshold', nms_cfg_.pop('iou_thr', None))
    max_num = nms_cfg_.pop('max_num', -1)
    n = boxes.shape[0]
    if n == 0:
        return boxes.new_zeros((0, 5)), torch.zeros((0,), dtype=torch.int64, device=boxes.device)
    if class_agnostic:
        mode = NMS_MODE_AGNOSTIC
    else:
        mode = NMS_MODE_OFFSET if n < split_thr else NMS_MODE_CLASS
    keep_pad, num = nms_device(boxes, scores, None if class_agnostic else idxs, thr, mode, max_keep=max_num)
    keep = keep_pad[:int(num.item())]
    dets = torch.cat([boxes[keep].float(), scores[keep].float()[:, None]], -1)
    return dets, keep


def bbox_decode_candidates(rois, cls_score, bbox_pred, num_classes, means, stds, img_shape=None, scale_factor=None,
                           score_thr=0.0, wh_ratio_clip=16 / 1000, reg_class_agnostic=False):
    """softmax + delta2bbox + clip (+ rescale) + multiclass candidate expansion in one launch (fixed shapes).
    Returns (cand_boxes [N*ncls,4], cand_scores [N*ncls] (-inf where score <= thr), cand_labels [N*ncls] int64,
    n_valid [1] int32 on the device)."""
    _lib.require_cuda(rois, cls_score, bbox_pred)
    rois, cls_score, bbox_pred = _f32c(rois), _f32c(cls_score), _f32c(bbox_pred)
    N = rois.shape[0]
    dev = rois.device
    n = N * num_classes
    boxes = torch.empty((n, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((n,), dtype=torch.float32, device=dev)
    labels = torch.empty((n,), dtype=torch.int64, device=dev)
    n_valid = torch.empty((1,), dtype=torch.int32, device=dev)
    f4 = ctypes.c_float * 4
    sf = f4(*[float(v) for v in scale_factor]) if scale_factor is not None else None
    h, w = (float(img_shape[0]), float(img_shape[1])) if img_shape is not None else (-1.0, -1.0)
    _lib.call('vod_bbox_decode_candidates', _lib.ptr(rois), _lib.ptr(cls_score), _lib.ptr(bbox_pred), N,
              int(num_classes), int(bool(reg_class_agnostic)), f4(*[float(v) for v in means]),
              f4(*[float(v) for v in stds]), float(abs(math.log(wh_ratio_clip))), h, w, sf, float(score_thr),
              _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(labels), _lib.ptr(n_valid), _lib.stream_ptr(dev))
    return boxes, scores, labels, n_valid


def multiclass_nms_device(cand_boxes, cand_scores, cand_labels, n_valid, iou_threshold, max_num, split_thr=10000,
                          class_agnostic=False):
    """Fixed-shape multiclass NMS without any host synchronisation (CUDA-graph capturable):
    returns (dets [max_num,5] zero-padded, labels [max_num] int64, count [1] int32), all on the device."""
    _lib.require_cuda(cand_boxes, cand_scores, cand_labels, n_valid)
    n = cand_boxes.shape[0]
    dev = cand_boxes.device
    dets = torch.zeros((max_num, 5), dtype=torch.float32, device=dev)
    labels = torch.zeros((max_num,), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n == 0:
        return dets, labels, count
    lib = _lib.load()
    ws = _ws.get(lib.vod_nms_workspace_bytes(n, n), dev)
    offs = (ctypes.c_int * 2)(0, n)
    _lib.call('vod_batched_nms_ex', _lib.ptr(cand_boxes), _lib.ptr(cand_scores), _lib.ptr(cand_labels), n, offs, 1,
              float(iou_threshold), NMS_MODE_AGNOSTIC if class_agnostic else 3, int(max_num), _lib.ptr(n_valid),
              int(split_thr), None, _lib.ptr(count), _lib.ptr(dets), _lib.ptr(labels), _lib.ptr(ws), ws.numel(),
              _lib.stream_ptr(dev))
    return dets, labels, count


def rpn_decode_topk(topk_idx, deltas, anchors, img_shape=None, wh_ratio_clip=16 / 1000):
    """topk_idx [B,K] int64, deltas [B,A,4], anchors [A,4] -> proposal boxes [B*K,4] (delta2bbox with means 0 / stds 1,
    clipped to img_shape = (h, w[, c]) when given)."""
    _lib.require_cuda(topk_idx, deltas, anchors)
    topk_idx = topk_idx.to(torch.int64).contiguous()
    deltas, anchors = _f32c(deltas), _f32c(anchors)
    B, K = topk_idx.shape
    A = anchors.shape[0]
    assert deltas.shape == (B, A, 4)
    boxes = torch.empty((B * K, 4), dtype=torch.float32, device=deltas.device)
    if B * K:
        import math
        h, w = (float(img_shape[0]), float(img_shape[1])) if img_shape is not None else (-1.0, -1.0)
        _lib.call('vod_rpn_decode_topk', _lib.ptr(topk_idx), _lib.ptr(deltas), _lib.ptr(anchors), _lib.ptr(boxes), B, K, A,
                  float(abs(math.log(wh_ratio_clip))), h, w, _lib.stream_ptr(deltas.device))
    return boxes


# ----------------------------------------------------------------------------- (2) warp / FGFA weighting
def flow_warp(x, flow):
    """x [N,C,H,W] (or [1,C,H,W]: one map shared by all N flows), flow [N,2,Hf,Wf] -> [N,C,H,W]."""
    _lib.require_cuda(x, flow)
    x, flow = _f32c(x), _f32c(flow)
    Nx, C, H, W = x.shape
    N = flow.shape[0]
    assert Nx in (1, N)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    if out.numel():
        _lib.call('vod_flow_warp_shared', _lib.ptr(x), _lib.ptr(flow), _lib.ptr(out), N, Nx, C, H, W, flow.shape[2],
                  flow.shape[3], _lib.stream_ptr(x.device))
    return out


def flow_warp_lowres(x, flow_lr, full_size, up_scale, mult1, mult2):
    """x [N,C,H,W] (or [1,C,H,W] shared), flow_lr [N,2,hl,wl]: warp by mult2 * (mult1 * upsample(flow_lr, up_scale)) of
    size ``full_size`` = (Hf, Wf), the upsample evaluated on the fly (include/vodagg.h)."""
    _lib.require_cuda(x, flow_lr)
    x, flow_lr = _f32c(x), _f32c(flow_lr)
    Nx, C, H, W = x.shape
    N = flow_lr.shape[0]
    assert Nx in (1, N)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    if out.numel():
        _lib.call('vod_flow_warp_lowres', _lib.ptr(x), _lib.ptr(flow_lr), _lib.ptr(out), N, Nx, C, H, W, flow_lr.shape[2],
                  flow_lr.shape[3], int(full_size[0]), int(full_size[1]), float(up_scale), float(mult1), float(mult2),
                  _lib.stream_ptr(x.device))
    return out


def embed_weighted_sum(key_emb, ref_emb, ref_x):
    """key_emb [1,C,H,W], ref_emb [T,C,H,W], ref_x [T,Cx,H,W] -> [1,Cx,H,W]."""
    _lib.require_cuda(key_emb, ref_emb, ref_x)
    key_emb, ref_emb, ref_x = _f32c(key_emb), _f32c(ref_emb), _f32c(ref_x)
    T, C, H, W = ref_emb.shape
    Cx = ref_x.shape[1]
    out = torch.empty((1, Cx, H, W), dtype=torch.float32, device=ref_x.device)
    ws = _ws.get(T * H * W * 4, ref_x.device)
    _lib.call('vod_embed_weighted_sum', _lib.ptr(key_emb), _lib.ptr(ref_emb), _lib.ptr(ref_x), _lib.ptr(out), T, C,
              Cx, H * W, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(ref_x.device))
    return out


def fgfa_warp_weighted_sum(key_emb, ref_emb, raw_x, flow, key_x=None, key_slot=-1):
    """Weighting with the warp recomputed on the fly from raw features + flows (fgfa.py:275-283)."""
    _lib.require_cuda(key_emb, ref_emb, raw_x, flow)
    key_emb, ref_emb, raw_x, flow = _f32c(key_emb), _f32c(ref_emb), _f32c(raw_x), _f32c(flow)
    key_x = _f32c(key_x) if key_x is not None else None
    T, C, H, W = ref_emb.shape
    Cx = raw_x.shape[1]
    out = torch.empty((1, Cx, H, W), dtype=torch.float32, device=raw_x.device)
    ws = _ws.get(T * H * W * 4, raw_x.device)
    _lib.call('vod_fgfa_warp_weighted_sum', _lib.ptr(key_emb), _lib.ptr(ref_emb), _lib.ptr(raw_x), _lib.ptr(flow),
              _lib.ptr(key_x), int(key_slot), _lib.ptr(out), T, C, Cx, H, W, flow.shape[2], flow.shape[3],
              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(raw_x.device))
    return out


# ----------------------------------------------------------------------------- (3) SELSA attention
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


def selsa_attention(q, k, v, num_heads, v_transposed=False, impl=IMPL_AUTO):
    """softmax(Q_h K_h^T / sqrt(d)) V_h per head.  q [N,D], k [M,D], v [M,D] (or V^T [D,ld] when
    v_transposed; ld >= M).  fp32 or bf16 inputs, fp32 output [N,D]."""
    _lib.require_cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype and q.dtype in (torch.float32, torch.bfloat16)
    q, k = q.contiguous(), k.contiguous()
    N, D = q.shape
    M = k.shape[0]
    d = D // num_heads
    assert d * num_heads == D
    if v_transposed:
        assert v.shape[0] == D and v.stride(1) == 1 and v.shape[1] >= M
        ldv = v.stride(0)
    else:
        v = v.contiguous()
        ldv = D
    out = torch.empty((N, D), dtype=torch.float32, device=q.device)
    if N == 0:
        return out
    lib = _lib.load()
    ws = _ws.get(lib.vod_selsa_attn_workspace_bytes(N, M, num_heads, d), q.device)
    # reference: weights = bmm(...) / (x_embed.shape[-1] ** 0.5)   (selsa_aggregator.py:61)
    scale = 1.0 / math.sqrt(d)
    _lib.call('vod_selsa_attn', _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(out), N, M, num_heads, d, scale,
              _lib.VOD_DTYPE_F32 if q.dtype == torch.float32 else _lib.VOD_DTYPE_BF16, int(bool(v_transposed)),
              int(ldv), int(impl), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(q.device))
    return out


def selsa_residual_relu_(x, y, bias, ref_x=None):
    """In place, one launch: x = relu(x + y + bias) and (optionally) ref_x = relu(ref_x) -- the element-wise tail of a
    SelsaBBoxHead layer (selsa_bbox_head.py:56-58).  x, y [rows, cols] fp32 contiguous, bias [cols], ref_x any contiguous fp32."""
    _lib.require_cuda(x, y, bias)
    assert x.dtype == y.dtype == bias.dtype == torch.float32 and x.shape == y.shape and x.dim() == 2
    assert x.is_contiguous() and y.is_contiguous() and bias.is_contiguous() and bias.numel() == x.shape[1]
    ref_ptr, ref_n = None, 0
    if ref_x is not None and ref_x.numel():
        assert ref_x.is_cuda and ref_x.dtype == torch.float32 and ref_x.is_contiguous()
        ref_ptr, ref_n = _lib.ptr(ref_x), ref_x.numel()
    _lib.call('vod_selsa_residual_relu', _lib.ptr(x), _lib.ptr(y), _lib.ptr(bias), x.shape[0], x.shape[1], ref_ptr, ref_n,
              _lib.stream_ptr(x.device))
    return x


# ----------------------------------------------------------------------------- (4) TemporalRoIAlign pieces
def _padded_unit(ref_unit, rows):
    """The tensor-core pass may read up to 3 rows past the last reference row (vodagg.h): copy into a padded buffer
    unless the tensor's storage already extends that far (the copies made by ``to_nhwc`` do)."""
    if ref_unit is None:
        return None
    ref_unit = ref_unit.contiguous()
    C = ref_unit.shape[-1]
    have = ref_unit.untyped_storage().nbytes() - ref_unit.storage_offset() * ref_unit.element_size()
    if have >= (rows + 3) * C * ref_unit.element_size():
        return ref_unit
    buf = torch.empty((rows + 4, C), dtype=ref_unit.dtype, device=ref_unit.device)
    buf[:rows].copy_(ref_unit.reshape(rows, C))
    return buf[:rows]


def msra_topk_sample(roi_rows, ref_nhwc, k=2, ref_norm=None, ref_unit=None, impl=IMPL_AUTO, return_indices=False,
                     out=None):
    """roi_rows [NP, C] fp32, ref_nhwc [T, H, W, C] fp32 -> out [T, NP, C] (+ idx [NP,T,k] int32, val fp32)."""
    _lib.require_cuda(roi_rows, ref_nhwc)
    roi_rows, ref_nhwc = _f32c(roi_rows), _f32c(ref_nhwc)
    NP, C = roi_rows.shape
    T, H, W, _ = ref_nhwc.shape
    dev = roi_rows.device
    if out is None:
        out = torch.empty((T, NP, C), dtype=torch.float32, device=dev)
    else:
        assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == T * NP * C
    idx = torch.empty((NP, T, k), dtype=torch.int32, device=dev) if return_indices else None
    val = torch.empty((NP, T, k), dtype=torch.float32, device=dev) if return_indices else None
    if NP and T:
        lib = _lib.load()
        ref_unit = _padded_unit(ref_unit, T * H * W)
        ws = _ws.get(lib.vod_msra_workspace_bytes(NP, C, T, H * W, k), dev)
        _lib.call('vod_msra_topk_sample', _lib.ptr(roi_rows), _lib.ptr(ref_nhwc), _lib.ptr(ref_norm),
                  _lib.ptr(ref_unit), _lib.ptr(out), _lib.ptr(idx), _lib.ptr(val), NP, C, T, H * W, int(k), int(impl),
                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    if return_indices:
        return out, idx, val
    return out


def msra_overflow_count(NP, C, T, HW, device):
    """Diagnostic: number of (row, frame) pairs of the LAST tensor-core ``msra_topk_sample`` call of these dimensions (on the
    current stream) whose candidate lists might have lost a member of the exact top-k and were re-scanned in exact fp32.
    Reads the counter the library keeps in the caller's workspace (one host sync)."""
    lib = _lib.load()
    ws = _ws.get(lib.vod_msra_workspace_bytes(NP, C, T, HW, 2), torch.device(device))
    off = int(lib.vod_msra_overflow_counter_offset(NP, C, T, HW))
    return int(ws[off:off + 4].view(torch.int32).item())


def msra_gemm_candidates(roi_unit, ref_unit, T):
    """bf16 unit rows roi_unit [NP, C], ref_unit [T*HW, C] -> packed candidate keys [NP, T, 16]
    (uint32 bit patterns in an int32 tensor; location = key & 0xFFF, 0 = empty slot)."""
    _lib.require_cuda(roi_unit, ref_unit)
    assert roi_unit.dtype == torch.bfloat16 and ref_unit.dtype == torch.bfloat16
    roi_unit = roi_unit.contiguous()
    NP, C = roi_unit.shape
    HW = ref_unit.shape[0] // T
    ref_unit = _padded_unit(ref_unit, T * HW)
    cand = torch.empty((NP, T, 16), dtype=torch.int32, device=roi_unit.device)
    _lib.call('vod_msra_gemm_candidates', _lib.ptr(roi_unit), _lib.ptr(ref_unit), _lib.ptr(cand), NP, C, T, HW,
              _lib.stream_ptr(roi_unit.device))
    return cand


def _tafa_out(out, N, P, C, out_nhwc, device):
    if out is None:
        return torch.empty((N, P, C) if out_nhwc else (N, C, P), dtype=torch.float32, device=device)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == N * P * C
    return out


def tafa_weighted_sum(x_all, emb_all, num_heads, out_nhwc=False, emb_bias=None, out=None):
    """x_all, emb_all [T1, N, P, C] fp32 -> [N, C, P] (or [N, P, C]).  ``emb_bias`` [C] is added to the
    embeddings on load (lets the embed conv run bias-free)."""
    _lib.require_cuda(x_all, emb_all, emb_bias)
    if emb_bias is not None:
        emb_bias = _f32c(emb_bias)
    assert x_all.is_contiguous() and x_all.dtype == torch.float32
    T1, N, P, C = x_all.shape
    if emb_all is not None:
        assert emb_all.is_contiguous() and emb_all.shape == x_all.shape and emb_all.dtype == torch.float32
    out = _tafa_out(out, N, P, C, out_nhwc, x_all.device)
    if N:
        _lib.call('vod_tafa_weighted_sum', _lib.ptr(x_all), _lib.ptr(emb_all), _lib.ptr(emb_bias), _lib.ptr(out), T1, N, P, C,
                  int(num_heads), int(bool(out_nhwc)), _lib.stream_ptr(x_all.device))
    return out


def tafa_keyproj_chunk(T1, P, C, num_heads):
    """Channel-chunk width the key-projected logits kernel wants G laid out with; 0 = shape unsupported."""
    return int(_lib.load().vod_tafa_keyproj_chunk(int(T1), int(P), int(C), int(num_heads)))


def tafa_keyproj_logits(x_all, G, output_size, num_heads, cc):
    """x_all [T1, N, P, C], G [heads, N*P, C/cc, 9, cc] fp32 or bf16 (key embedding x conv weight, see include/vodagg.h)
    -> per-chunk partial logits [C/cc, N, P, heads, T1] (unscaled)."""
    _lib.require_cuda(x_all, G)
    ph, pw = _pair(output_size)
    T1, N, P, C = x_all.shape
    assert P == ph * pw and x_all.is_contiguous() and x_all.dtype == torch.float32
    assert G.is_contiguous() and G.dtype in (torch.float32, torch.bfloat16) and G.numel() == num_heads * N * P * 9 * C
    parts = torch.empty((C // cc, N, P, num_heads, T1), dtype=torch.float32, device=x_all.device)
    if N:
        _lib.call('vod_tafa_keyproj_logits', _lib.ptr(x_all), _lib.ptr(G),
                  _lib.VOD_DTYPE_F32 if G.dtype == torch.float32 else _lib.VOD_DTYPE_BF16, _lib.ptr(parts), T1, N, ph, pw, C,
                  int(num_heads), int(cc), _lib.stream_ptr(x_all.device))
    return parts


def tafa_weighted_sum_logits(x_all, parts, num_heads, out_nhwc=False, out=None):
    """x_all [T1, N, P, C], partial logits [nparts, N, P, heads, T1] -> [N, C, P] (or [N, P, C])."""
    _lib.require_cuda(x_all, parts)
    assert x_all.is_contiguous() and x_all.dtype == torch.float32
    T1, N, P, C = x_all.shape
    assert parts.is_contiguous() and parts.dtype == torch.float32 and parts.shape[1:] == (N, P, num_heads, T1)
    out = _tafa_out(out, N, P, C, out_nhwc, x_all.device)
    if N:
        _lib.call('vod_tafa_weighted_sum_logits', _lib.ptr(x_all), _lib.ptr(parts), int(parts.shape[0]), _lib.ptr(out),
                  T1, N, P, C, int(num_heads), int(bool(out_nhwc)), _lib.stream_ptr(x_all.device))
    return out


def mdcn_im2col(x_nhwc, p, q=None, deform_groups=1, kernel_size=3, stride=1, padding=1, dilation=1, out=None):
    """Sampling half of mmcv's ``modulated_deform_conv2d`` (DCNv2) for ModulatedDCNPack
    (denoising2_aggregator.py:72-82): x_nhwc [B,H,W,C]; ``p`` (+ optional ``q``) raw ``conv_offset`` outputs, channels-last
    [B or 1, Ho, Wo, 3*G*K] (offsets then mask logits; a leading dim of 1 is broadcast over B).  Returns the modulated columns
    [B*Ho*Wo, K*C]; ``columns @ weight.permute(0, 2, 3, 1).reshape(Cout, K*C).t() + bias`` is the op's channels-last output."""
    _lib.require_cuda(x_nhwc, p, q)
    kh, kw = _pair(kernel_size)
    x_nhwc, p = _f32c(x_nhwc), _f32c(p)
    B, H, W, C = x_nhwc.shape
    Ho = (H + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    och = 3 * deform_groups * kh * kw
    assert p.shape[1:] == (Ho, Wo, och) and p.shape[0] in (1, B), (tuple(p.shape), (B, Ho, Wo, och))
    if q is not None:
        q = _f32c(q)
        assert q.shape[1:] == (Ho, Wo, och) and q.shape[0] in (1, B), tuple(q.shape)
    if out is None:
        out = torch.empty((B * Ho * Wo, kh * kw * C), dtype=torch.float32, device=x_nhwc.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == B * Ho * Wo * kh * kw * C
    if B:
        _lib.call('vod_mdcn_im2col', _lib.ptr(x_nhwc), _lib.ptr(p), _lib.ptr(q), _lib.ptr(out), B, C, H, W, int(deform_groups),
                  kh, kw, int(stride), int(padding), int(dilation), int(p.shape[0] == 1 and B > 1), int(q is not None and q.shape[0] == 1 and B > 1),
                  _lib.stream_ptr(x_nhwc.device))
    return out


def temporal_softmax_fuse(cor, x, out=None):
    """cor [I, T, ...], x [T, ...] (same trailing element order) -> [I, ...]: softmax over the T frames of ``cor`` and the
    weighted sum of ``x`` in one pass (TemporalAttentionFusion.forward, denoising2_aggregator.py:145-146)."""
    _lib.require_cuda(cor, x)
    cor, x = _f32c(cor), _f32c(x)
    I, T = cor.shape[:2]
    assert x.shape[0] == T and cor.shape[2:] == x.shape[1:], (tuple(cor.shape), tuple(x.shape))
    E = x[0].numel()
    assert E % 4 == 0, 'temporal_softmax_fuse: per-frame element count must be a multiple of 4'
    if out is None:
        out = torch.empty((I,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    assert out.is_contiguous() and out.numel() == I * E
    if I and E:
        _lib.call('vod_temporal_softmax_fuse', _lib.ptr(cor), _lib.ptr(x), _lib.ptr(out), I, T, E, _lib.stream_ptr(x.device))
    return out


def test_gemm_nt(a, b, a_in_tmem=False):
    """D = A @ B^T on the tcgen05 path (unit-test hook for the descriptor/pipeline building blocks; lives in
    libvodagg_selftest.so, not in the product library).
    ``a_in_tmem``: bf16 only; the A tile is written to TMEM with tcgen05.st and consumed by the TS-form MMA."""
    _lib.require_cuda(a, b)
    assert a.dtype == b.dtype and a.dtype in (torch.float32, torch.bfloat16)
    a, b = a.contiguous(), b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    d = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _lib.call_selftest('vod_test_gemm_nt', _lib.ptr(a), _lib.ptr(b), _lib.ptr(d), M, N, K,
              2 if a_in_tmem else (_lib.VOD_DTYPE_F32 if a.dtype == torch.float32 else _lib.VOD_DTYPE_BF16),
              _lib.stream_ptr(a.device))
    return d
