"""Callers of kernel (5) restated from mmdet (they stay host Python, as in the reference):

  multiclass_nms   mmdetection/mmdet/core/post_processing/bbox_nms.py:7-93
  delta2bbox       mmdetection/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:134-237
  bbox2roi         mmdetection/mmdet/core/bbox/transforms.py:58-77
  rpn_batched_nms  the per-image batched_nms of RPNHead._get_bboxes
                   (mmdetection/mmdet/models/dense_heads/rpn_head.py:233-236), batched over the
                   key + reference images of one step in a single launch set
"""
import numpy as np
import torch

from . import ops


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None,
                   return_inds=False):
    """NMS for multi-class bboxes (same contract as mmdet's; labels are created on the boxes' device).

    ``keep`` indexes the score-filtered candidate list, exactly as in the reference.  When ``max_num > 0``
    it is also handed to the device sweep (``nms_cfg['max_num']``, an mmcv>=1.3 key) so the sweep stops
    after ``max_num`` survivors; the result equals truncating the full result."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, 4)
    scores = multi_scores[:, :-1]
    labels = torch.arange(num_classes, dtype=torch.long, device=scores.device)
    labels = labels.view(1, -1).expand_as(scores)
    bboxes = bboxes.reshape(-1, 4)
    scores = scores.reshape(-1)
    labels = labels.reshape(-1)
    valid_mask = scores > score_thr
    if score_factors is not None:
        score_factors = score_factors.view(-1, 1).expand(multi_scores.size(0), num_classes)
        scores = scores * score_factors.reshape(-1)
    inds = valid_mask.nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if bboxes.numel() == 0:
        if return_inds:
            return bboxes, labels, inds
        return bboxes, labels
    cfg = dict(nms_cfg)
    if max_num > 0 and 'max_num' not in cfg:
        cfg['max_num'] = max_num
    dets, keep = ops.batched_nms(bboxes, scores, labels, cfg)
    if max_num > 0:
        dets = dets[:max_num]
        keep = keep[:max_num]
    if return_inds:
        return dets, labels[keep], keep
    return dets, labels[keep]


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None,
               wh_ratio_clip=16 / 1000, clip_border=True):
    """delta_xywh_bbox_coder.py:134-237 (elementwise glue; stays torch)."""
    means = deltas.new_tensor(means).view(1, -1).repeat(1, deltas.size(-1) // 4)
    stds = deltas.new_tensor(stds).view(1, -1).repeat(1, deltas.size(-1) // 4)
    denorm_deltas = deltas * stds + means
    dx = denorm_deltas[..., 0::4]
    dy = denorm_deltas[..., 1::4]
    dw = denorm_deltas[..., 2::4]
    dh = denorm_deltas[..., 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    x1, y1 = rois[..., 0], rois[..., 1]
    x2, y2 = rois[..., 2], rois[..., 3]
    px = ((x1 + x2) * 0.5).unsqueeze(-1).expand_as(dx)
    py = ((y1 + y2) * 0.5).unsqueeze(-1).expand_as(dy)
    pw = (x2 - x1).unsqueeze(-1).expand_as(dw)
    ph = (y2 - y1).unsqueeze(-1).expand_as(dh)
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = px + pw * dx
    gy = py + ph * dy
    x1 = gx - gw * 0.5
    y1 = gy - gh * 0.5
    x2 = gx + gw * 0.5
    y2 = gy + gh * 0.5
    bboxes = torch.stack([x1, y1, x2, y2], dim=-1).view(deltas.size())
    if clip_border and max_shape is not None:
        x = bboxes.view(-1, 4)
        x[:, 0::2].clamp_(min=0, max=max_shape[1])
        x[:, 1::2].clamp_(min=0, max=max_shape[0])
    return bboxes


def bbox2roi(bbox_list):
    """transforms.py:58-77: list of [n,4(+)] per image -> [sum n, 5] with the image index in column 0."""
    if len(bbox_list) > 1 and all(b.shape == bbox_list[0].shape and b.shape[0] > 0 for b in bbox_list):
        # equal-sized proposal sets (the test-time case: max_num per frame): one stack instead of 2 kernels per image
        n = bbox_list[0].shape[0]
        boxes = torch.stack([b[:, :4] for b in bbox_list], 0)
        ids = torch.arange(len(bbox_list), device=boxes.device, dtype=boxes.dtype).view(-1, 1, 1).expand(-1, n, 1)
        return torch.cat([ids, boxes], dim=-1).view(-1, 5)
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def rpn_batched_nms(proposals_list, scores_list, iou_threshold=0.7, max_num=300):
    """NMS of the RPN proposals of several images in one launch set (rpn_head.py:233-236 per image:
    single level => all-zero level ids => class-agnostic arithmetic after a zero offset).

    proposals_list[i] [n_i, 4], scores_list[i] [n_i] -> list of dets [<=max_num, 5] per image."""
    offs = [0]
    for p in proposals_list:
        offs.append(offs[-1] + p.shape[0])
    boxes = torch.cat(proposals_list, 0)
    scores = torch.cat(scores_list, 0)
    if boxes.shape[0] == 0:
        return [boxes.new_zeros((0, 5)) for _ in proposals_list]
    # mmcv adds idxs * (max + 1) with idxs == 0: boxes + 0.0 is exact, so mode 0 is bit-identical
    keep, num = ops.nms_device(boxes, scores, None, iou_threshold, ops.NMS_MODE_AGNOSTIC, seg_offsets=offs,
                               max_keep=max_num)
    nums = num.tolist()
    out = []
    for i in range(len(proposals_list)):
        k = keep[offs[i]:offs[i] + nums[i]]
        out.append(torch.cat([proposals_list[i][k].float(), scores_list[i][k].float()[:, None]], dim=1))
    return out


def rpn_get_bboxes_device(cls_scores, bbox_preds, anchors, img_shape, nms_pre=6000, nms_thr=0.7, max_per_img=300):
    """The proposal stage of RPNHead._get_bboxes (rpn_head.py:126-236) for ONE feature level (the R-50-DC5 configs),
    all images of the batch at once, fixed shapes, no host synchronisation (CUDA-graph capturable):

    cls_scores [B, A_per, H, W] logits (sigmoid scores), bbox_preds [B, A_per*4, H, W], anchors [H*W*A_per, 4]
    -> (proposals [B, max_per_img, 5] zero-padded (x1, y1, x2, y2, score), counts [B] int32).

    sigmoid -> best nms_pre positions by score (sorted) -> delta2bbox + clip (one kernel for all images) -> segmented NMS
    with the bounded-survivor sweep.  ``min_bbox_size`` is 0 in every config of the reference (faster_rcnn_r50_dc5.py:104-110)."""
    B = cls_scores.shape[0]
    scores = cls_scores.permute(0, 2, 3, 1).reshape(B, -1).sigmoid()              # :131-135
    deltas = bbox_preds.permute(0, 2, 3, 1).reshape(B, -1, 4)                      # :143-144
    A = scores.shape[1]
    K = min(int(nms_pre), A) if nms_pre > 0 else A
    ranked, order = scores.sort(dim=1, descending=True)                           # :163-170 (a full sort beats topk here, as the reference notes)
    top_scores, top_idx = ranked[:, :K].contiguous(), order[:, :K].contiguous()
    boxes = ops.rpn_decode_topk(top_idx, deltas, anchors, img_shape)              # :187-188
    offs = [K * i for i in range(B + 1)]
    keep, num = ops.nms_device(boxes, top_scores.reshape(-1), None, nms_thr, ops.NMS_MODE_AGNOSTIC, seg_offsets=offs,
                               max_keep=max_per_img)                              # :233-236 (level ids all 0)
    # keep[b*K : b*K + num[b]] are the kept positions of image b in descending-score order
    k_idx = keep.view(B, K)[:, :max_per_img]
    if k_idx.shape[1] < max_per_img:
        k_idx = torch.nn.functional.pad(k_idx, (0, max_per_img - k_idx.shape[1]))
    valid = torch.arange(max_per_img, device=keep.device)[None, :] < num[:, None]
    k_idx = torch.where(valid, k_idx, torch.zeros_like(k_idx))
    gb = boxes.view(B, K, 4).gather(1, k_idx[:, :, None].expand(-1, -1, 4))
    gs = top_scores.gather(1, k_idx)
    props = torch.cat([gb, gs[:, :, None]], dim=2) * valid[:, :, None]
    return props, num


def rpn_get_bboxes(cls_scores, bbox_preds, anchors, img_shape, nms_pre=6000, nms_thr=0.7, max_per_img=300):
    """List-of-tensors form of ``rpn_get_bboxes_device`` (the reference's result_list, rpn_head.py:219-236): one host
    read of the per-image counts trims the padded proposals."""
    props, num = rpn_get_bboxes_device(cls_scores, bbox_preds, anchors, img_shape, nms_pre, nms_thr, max_per_img)
    return [props[i, :n] for i, n in enumerate(num.tolist())]
