"""oracle/vod_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's multi-frame feature-aggregation hot path
(SURVEY.md section 8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product package never does (tests/test_boundary.py greps for it).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  GEMM-shaped arithmetic uses torch CPU ops because that is
where the reference's own arithmetic executes (ATen); the exact-arithmetic
pieces whose source is the un-vendored mmcv-full (RoIAlign, NMS) and the
closed-form flow warp are restated in plain C (oracle/vod_oracle.c).

Pinning: the reference's tests hold NO numeric golden vectors for
this path ("parity unpinned" upstream).  The oracle is therefore pinned against
(i) outputs of the reference's own unmodified Python files run in the build
container under the shims of oracle/ref_shim.py, committed as fixtures under
tests/golden/ by tests/golden/make_golden.py, and (ii) torchvision's
roi_align/nms, the executable stand-in for mmcv's ops (mmcv's own
``use_torchvision`` switch asserts that equivalence).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle():
    """Compile oracle/vod_oracle.c with the committed Makefile (idempotent)."""
    subprocess.run(['make', '-s', '-C', _HERE], check=True)
    return os.path.join(_HERE, '_build', 'libvod_oracle.so')


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libvod_oracle.so')
        if not os.path.exists(path):
            build_c_oracle()
        lib = ctypes.CDLL(path)
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_roi_align.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 7 + \
            [ctypes.c_float, ctypes.c_int, ctypes.c_int]
        lib.oracle_roi_align.restype = None
        lib.oracle_nms_sorted.argtypes = [f32p, i64p, ctypes.c_int, ctypes.c_float, i64p]
        lib.oracle_nms_sorted.restype = ctypes.c_int
        lib.oracle_flow_warp.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 6
        lib.oracle_flow_warp.restype = None
        _LIB = lib
    return _LIB


def _f32(t):
    t = t.detach().to('cpu', torch.float32).contiguous()
    return t, ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------- a1
def roi_align(feat, rois, output_size=7, spatial_scale=1.0 / 16, sampling_ratio=2,
              aligned=True):
    """mmcv.ops.RoIAlign(pool_mode='avg') forward.

    Call sites: mmdetection/mmdet/models/roi_heads/roi_extractors/
    base_roi_extractor.py:49-55 (construction) and
    single_level_roi_extractor.py:72-75 (single-level fast path).
    feat [B,C,H,W], rois [K,5] -> [K,C,ph,pw].
    """
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    feat, fp = _f32(feat)
    rois, rp = _f32(rois)
    B, C, H, W = feat.shape
    K = rois.shape[0]
    out = torch.zeros(K, C, ph, pw, dtype=torch.float32)
    if K:
        _lib().oracle_roi_align(fp, rp, ctypes.cast(out.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                B, C, H, W, K, ph, pw, float(spatial_scale),
                                int(sampling_ratio), int(bool(aligned)))
    return out


# --------------------------------------------------------------------- a8/a9
def nms(boxes, scores, iou_threshold):
    """mmcv.ops.nms (offset=0): returns (dets[k,5], keep[k] int64), keep in
    descending-score order.  Sort = stable descending (score desc, index asc)."""
    boxes, bp = _f32(boxes)
    scores = scores.detach().to('cpu', torch.float32).contiguous()
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, 5), torch.zeros(0, dtype=torch.int64)
    order = torch.sort(scores, descending=True, stable=True).indices.contiguous()
    keep = torch.empty(n, dtype=torch.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    nk = _lib().oracle_nms_sorted(bp, ctypes.cast(order.data_ptr(), i64p), n,
                                  float(iou_threshold), ctypes.cast(keep.data_ptr(), i64p))
    keep = keep[:nk].clone()
    return torch.cat([boxes[keep], scores[keep, None]], dim=1), keep


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.nms.batched_nms (mmcv 1.2.x).

    Call sites: mmdetection/mmdet/core/post_processing/bbox_nms.py:84 and
    mmdetection/mmdet/models/dense_heads/rpn_head.py:233-235.
    """
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    boxes = boxes.detach().to('cpu', torch.float32)
    scores = scores.detach().to('cpu', torch.float32)
    idxs = idxs.detach().to('cpu')
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + 1)
        boxes_for_nms = boxes + offsets[:, None]
    cfg.pop('type', 'nms')
    split_thr = cfg.pop('split_thr', 10000)
    thr = cfg.pop('iou_threshold', cfg.pop('iou_thr', None))
    if boxes_for_nms.shape[0] < split_thr:
        dets, keep = nms(boxes_for_nms, scores, thr)
        boxes = boxes[keep]
        scores = dets[:, -1]
    else:
        total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
        for id in torch.unique(idxs):
            mask = (idxs == id).nonzero(as_tuple=False).view(-1)
            dets, keep = nms(boxes_for_nms[mask], scores[mask], thr)
            total_mask[mask[keep]] = True
        keep = total_mask.nonzero(as_tuple=False).view(-1)
        keep = keep[scores[keep].argsort(descending=True, stable=True)]
        boxes = boxes[keep]
        scores = scores[keep]
    return torch.cat([boxes, scores[:, None]], -1), keep


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1,
                   return_inds=False):
    """mmdetection/mmdet/core/post_processing/bbox_nms.py:7-93 (score_factors=None)."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, 4)
    scores = multi_scores[:, :-1]
    labels = torch.arange(num_classes, dtype=torch.long)
    labels = labels.view(1, -1).expand_as(scores)
    bboxes = bboxes.reshape(-1, 4)
    scores = scores.reshape(-1)
    labels = labels.reshape(-1)
    valid_mask = scores > score_thr
    inds = valid_mask.nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if bboxes.numel() == 0:
        if return_inds:
            return bboxes, labels, inds
        return bboxes, labels
    dets, keep = batched_nms(bboxes, scores, labels, nms_cfg)
    if max_num > 0:
        dets = dets[:max_num]
        keep = keep[:max_num]
    if return_inds:
        return dets, labels[keep], keep
    return dets, labels[keep]


# --------------------------------------------------------------------- a5
def selsa_attention(q, k, v, num_heads):
    """Core of SelsaAggregator.forward, mmtracking/mmtrack/models/aggregators/
    selsa_aggregator.py:51-70: per-head softmax(Q K^T / sqrt(d)) V."""
    n, m = q.shape[0], k.shape[0]
    qh = q.view(n, num_heads, -1).permute(1, 0, 2)
    kh = k.view(m, num_heads, -1).permute(1, 2, 0)
    w = torch.bmm(qh, kh) / (qh.shape[-1] ** 0.5)
    w = w.softmax(dim=2)
    vh = v.view(m, num_heads, -1).permute(1, 0, 2)
    return torch.bmm(w, vh).permute(1, 0, 2).contiguous().view(n, -1)


def selsa_aggregate(x, ref_x, p, num_heads=16):
    """SelsaAggregator.forward, selsa_aggregator.py:29-73.  ``p`` holds
    fc_embed/ref_fc_embed/fc/ref_fc .weight/.bias (state_dict names)."""
    lin = torch.nn.functional.linear
    q = lin(x, p['fc_embed.weight'], p['fc_embed.bias'])
    k = lin(ref_x, p['ref_fc_embed.weight'], p['ref_fc_embed.bias'])
    v = lin(ref_x, p['ref_fc.weight'], p['ref_fc.bias'])
    o = selsa_attention(q, k, v, num_heads)
    return lin(o, p['fc.weight'], p['fc.bias'])


# --------------------------------------------------------------------- a6
def flow_warp_feats(x, flow):
    """mmtracking/mmtrack/core/motion/flow.py:4-41 (closed form, C)."""
    assert len(x.shape) == 4
    assert len(flow.shape) == 4 and flow.shape[1] == 2
    x, xp = _f32(x)
    flow, fp = _f32(flow)
    N, C, H, W = x.shape
    out = torch.empty_like(x)
    _lib().oracle_flow_warp(xp, fp, ctypes.cast(out.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                            N, C, H, W, flow.shape[2], flow.shape[3])
    return out


# --------------------------------------------------------------------- a7
def embed_weighted_sum(x_embed, ref_x_embed, ref_x):
    """Weighting half of EmbedAggregator.forward, mmtracking/mmtrack/models/
    aggregators/embed_aggregator.py:71-81 (after the embed convs)."""
    x_embed = x_embed / x_embed.norm(p=2, dim=1, keepdim=True)
    ref_x_embed = ref_x_embed / ref_x_embed.norm(p=2, dim=1, keepdim=True)
    ada = torch.sum(ref_x_embed * x_embed, dim=1, keepdim=True).softmax(dim=0)
    return torch.sum(ref_x * ada, dim=0, keepdim=True)


def embed_aggregate(x, ref_x, convs):
    """EmbedAggregator.forward, embed_aggregator.py:50-81.  ``convs`` is a list
    of (weight, bias, relu: bool) for the 3x3 embed convs."""
    assert len(x.shape) == 4 and len(x) == 1
    def run(t):
        for w, b, relu in convs:
            t = torch.nn.functional.conv2d(t, w, b, padding=(w.shape[-1] - 1) // 2)
            if relu:
                t = torch.relu(t)
        return t
    return embed_weighted_sum(run(x), run(ref_x), ref_x)


# --------------------------------------------------------------------- a2
def most_similar_roi_align(roi_feats, ref_feats, k=2, return_indices=False):
    """TemporalRoIAlign.most_similar_roi_align, mmtracking/mmtrack/models/
    roi_heads/roi_extractors/temporal_roi_align.py:99-181."""
    roi_e = roi_feats / roi_feats.norm(p=2, dim=1, keepdim=True)          # :127
    ref_e = ref_feats / ref_feats.norm(p=2, dim=1, keepdim=True)          # :129
    roi_n, c, rh, rw = roi_e.shape
    img_n, _, ih, iw = ref_e.shape
    a = roi_e.permute(0, 2, 3, 1).contiguous().view(-1, c)                 # :134-136
    b = ref_e.permute(1, 0, 2, 3).contiguous().view(c, -1)                 # :138-140
    sim = a.mm(b).view(-1, img_n, ih * iw)                                 # :142-145
    values, indices = sim.topk(k=k, dim=2, largest=True)                   # :149-153
    weights = values.softmax(dim=2)                                        # :155
    ref_r = ref_feats.permute(2, 3, 0, 1).contiguous().view(-1, img_n, c)  # :159-161
    outs = []
    for i in range(img_n):                                                 # :165-176
        feats = ref_r[indices[:, i], i, :]
        outs.append((feats * weights[:, i].unsqueeze(-1)).sum(dim=1))
    out = torch.stack(outs, 0).view(img_n, roi_n, rh, rw, c).permute(0, 1, 4, 2, 3)
    if return_indices:
        return out, indices, sim
    return out


# --------------------------------------------------------------------- a3
def tafa_weighted_sum(x_all, x_embed, num_blocks):
    """Weighting half of temporal_attentional_feature_aggregation,
    temporal_roi_align.py:77-97.  x_all, x_embed: [T+1, N, C, h, w]."""
    img_n, roi_n, c, h, w = x_embed.shape
    e = x_embed.view(img_n, roi_n, num_blocks, -1, h, w)
    tgt = e[[0]]
    ada = torch.sum(e * tgt, dim=3, keepdim=True) / (float(c / num_blocks) ** 0.5)
    ada = ada.expand(-1, -1, -1, int(c / num_blocks), -1, -1).contiguous()
    ada = ada.view(img_n, roi_n, c, h, w).softmax(dim=0)
    return (x_all * ada).sum(dim=0)


def tafa(x, ref_x, conv_w, conv_b, num_blocks):
    """temporal_attentional_feature_aggregation, temporal_roi_align.py:44-97."""
    x_all = torch.cat((x, ref_x), dim=0)
    img_n, roi_n, c, h, w = x_all.shape
    emb = torch.nn.functional.conv2d(x_all.reshape(img_n * roi_n, c, h, w), conv_w, conv_b, padding=1)
    return tafa_weighted_sum(x_all, emb.view(img_n, roi_n, -1, h, w), num_blocks)


def temporal_roi_align(feat, rois, ref_feat, conv_w, conv_b, k=2, num_blocks=4,
                       output_size=7, spatial_scale=1.0 / 16, sampling_ratio=2):
    """TemporalRoIAlign.forward with ref_feats, temporal_roi_align.py:183-207."""
    roi_feats = roi_align(feat, rois, output_size, spatial_scale, sampling_ratio, True)
    ref_roi_feats = most_similar_roi_align(roi_feats, ref_feat, k)
    roi_feats = roi_feats.unsqueeze(0)
    if num_blocks > 0:
        return tafa(roi_feats, ref_roi_feats, conv_w, conv_b, num_blocks)
    return torch.cat((roi_feats, ref_roi_feats), dim=0).mean(dim=0)


# --------------------------------------------------------------------- a10 (callers)
def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               max_shape=None, wh_ratio_clip=16 / 1000):
    """mmdetection/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:134-237."""
    means = deltas.new_tensor(means).view(1, -1).repeat(1, deltas.size(1) // 4)
    stds = deltas.new_tensor(stds).view(1, -1).repeat(1, deltas.size(1) // 4)
    d = deltas * stds + means
    dx, dy, dw, dh = d[:, 0::4], d[:, 1::4], d[:, 2::4], d[:, 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dh)
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + pw * dx, py + ph * dy
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1]); y1 = y1.clamp(min=0, max=max_shape[0])
        x2 = x2.clamp(min=0, max=max_shape[1]); y2 = y2.clamp(min=0, max=max_shape[0])
    return torch.stack([x1, y1, x2, y2], dim=-1).view(deltas.size())


def selsa_bbox_head(x, ref_x, p, num_shared_fcs, num_heads=16):
    """SelsaBBoxHead.forward, mmtracking/mmtrack/models/roi_heads/bbox_heads/
    selsa_bbox_head.py:25-84 (no shared convs / avg pool, as in the configs).
    ``p``: shared_fcs.{i}.weight/bias, aggregator.{i}.<...>, fc_cls.*, fc_reg.*"""
    lin = torch.nn.functional.linear
    x = x.flatten(1)
    ref_x = ref_x.flatten(1)
    for i in range(num_shared_fcs):
        w, b = p['shared_fcs.%d.weight' % i], p['shared_fcs.%d.bias' % i]
        x = lin(x, w, b)
        ref_x = lin(ref_x, w, b)
        agg = {kk[len('aggregator.%d.' % i):]: vv for kk, vv in p.items()
               if kk.startswith('aggregator.%d.' % i)}
        x = x + selsa_aggregate(x, ref_x, agg, num_heads)
        ref_x = torch.relu(ref_x)
        x = torch.relu(x)
    return lin(x, p['fc_cls.weight'], p['fc_cls.bias']), lin(x, p['fc_reg.weight'], p['fc_reg.bias'])


def get_bboxes(rois, cls_score, bbox_pred, img_shape, scale_factor, rescale, score_thr,
               nms_cfg, max_per_img, return_inds=False, target_means=(0., 0., 0., 0.),
               target_stds=(0.2, 0.2, 0.2, 0.2)):
    """BBoxHead.get_bboxes, mmdetection/mmdet/models/roi_heads/bbox_heads/
    bbox_head.py:269-373 (softmax scores, class-specific regression)."""
    scores = torch.softmax(cls_score, dim=1)
    bboxes = delta2bbox(rois[:, 1:], bbox_pred, target_means, target_stds, max_shape=img_shape)
    if rescale and bboxes.size(0) > 0:
        sf = bboxes.new_tensor(scale_factor)
        bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size()[0], -1)
    return multiclass_nms(bboxes, scores, score_thr, nms_cfg, max_per_img, return_inds=return_inds)


def rpn_get_bboxes(cls_score, bbox_pred, anchors, img_shape, nms_pre=6000, nms_thr=0.7, max_per_img=300,
                   min_bbox_size=0):
    """RPNHead._get_bboxes for one image and one feature level (sigmoid scores),
    mmdetection/mmdet/models/dense_heads/rpn_head.py:126-236:
    cls_score [A_per, H, W] logits, bbox_pred [A_per*4, H, W], anchors [H*W*A_per, 4] -> dets [<=max_per_img, 5]."""
    scores = cls_score.permute(1, 2, 0).reshape(-1).sigmoid()                     # :131-135
    deltas = bbox_pred.permute(1, 2, 0).reshape(-1, 4)                            # :143-144
    if nms_pre > 0 and scores.shape[0] > nms_pre:                                 # :163-170
        ranked, rank_inds = scores.sort(descending=True)
        topk = rank_inds[:nms_pre]
        scores, deltas, anchors = ranked[:nms_pre], deltas[topk], anchors[topk]
    proposals = delta2bbox(anchors, deltas, (0., 0., 0., 0.), (1., 1., 1., 1.), max_shape=img_shape)   # :187-188
    if min_bbox_size > 0:                                                         # :222-231
        w, h = proposals[:, 2] - proposals[:, 0], proposals[:, 3] - proposals[:, 1]
        ok = (w >= min_bbox_size) & (h >= min_bbox_size)
        proposals, scores = proposals[ok], scores[ok]
    ids = torch.zeros(len(scores), dtype=torch.long)                              # single level -> level id 0
    dets, _ = batched_nms(proposals, scores, ids, dict(type='nms', iou_threshold=nms_thr))   # :233-235
    return dets[:max_per_img]


# ----------------------------------------------------------------------------- temporal attention fusion (Denoising2Aggergator)
def _cpu32(t):
    return t.detach().to('cpu', torch.float32).contiguous()


def modulated_deform_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deform_groups=1):
    """mmcv.ops.modulated_deform_conv2d (DCNv2) -- mmcv-full 1.2.x, NOT vendored in /root/reference; call site
    mmtracking/mmtrack/models/aggregators/denoising2_aggregator.py:79-82.  Published algorithm (modulated_deformable_im2col):
    for output pixel (ho, wo), tap k = (i, j) and deformable group g, the sample point is
    (ho*stride - pad + i*dil + offset[g*2K + 2k], wo*stride - pad + j*dil + offset[g*2K + 2k + 1]); its value is the bilinear
    interpolation of the input (corners outside the map contribute 0; 0 when the point is not inside (-1, H) x (-1, W)) times
    mask[g*K + k]; the columns are then multiplied by the weight.  Pinned against torchvision.ops.deform_conv2d (the same
    DCN-derived layout and border rule) in tests/test_oracle.py."""
    assert groups == 1
    x, offset, mask, weight = _cpu32(x), _cpu32(offset), _cpu32(mask), _cpu32(weight)
    B, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    K, G = kh * kw, deform_groups
    Ho = (H + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    assert offset.shape == (B, 2 * G * K, Ho, Wo) and mask.shape == (B, G * K, Ho, Wo)
    Cg = C // G
    ho = torch.arange(Ho, dtype=torch.float32).view(1, 1, Ho, 1) * stride - padding
    wo = torch.arange(Wo, dtype=torch.float32).view(1, 1, 1, Wo) * stride - padding
    ki = (torch.arange(K) // kw).float().view(1, K, 1, 1) * dilation
    kj = (torch.arange(K) % kw).float().view(1, K, 1, 1) * dilation
    cols = x.new_zeros(B, C, K, Ho, Wo)
    xf = x.reshape(B, C, H * W)
    for g in range(G):
        off = offset[:, g * 2 * K:(g + 1) * 2 * K].view(B, K, 2, Ho, Wo)
        h = ho + ki + off[:, :, 0]
        w = wo + kj + off[:, :, 1]                                        # [B, K, Ho, Wo]
        inside = (h > -1) & (w > -1) & (h < H) & (w < W)
        h0, w0 = torch.floor(h), torch.floor(w)
        lh, lw = h - h0, w - w0
        val = x.new_zeros(B, Cg, K, Ho, Wo)
        for dh, dw, wt in ((0, 0, (1 - lh) * (1 - lw)), (0, 1, (1 - lh) * lw), (1, 0, lh * (1 - lw)), (1, 1, lh * lw)):
            hi, wi = (h0 + dh).long(), (w0 + dw).long()
            ok = inside & (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
            lin = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).view(B, 1, -1).expand(B, Cg, -1)
            v = torch.gather(xf[:, g * Cg:(g + 1) * Cg], 2, lin).view(B, Cg, K, Ho, Wo)
            val = val + v * (wt * ok)[:, None]
        cols[:, g * Cg:(g + 1) * Cg] = val * mask[:, g * K:(g + 1) * K][:, None]
    out = torch.einsum('ock,bckp->bop', weight.reshape(Cout, C, K), cols.reshape(B, C, K, Ho * Wo)).view(B, Cout, Ho, Wo)
    if bias is not None:
        out = out + _cpu32(bias).view(1, -1, 1, 1)
    return out


def temporal_attention_fusion(x, p):
    """TemporalAttentionFusion.forward (denoising2_aggregator.py:135-152) with the parameters in ``p`` (state_dict names).
    Written as the reference runs it: per reference frame i, T pair-wise offset convs, DCN, embed convs, softmax over frames."""
    F_ = torch.nn.functional
    x = torch.relu(F_.conv2d(_cpu32(x), p['conv1.weight'], p['conv1.bias'], padding=1))
    T = x.shape[0]
    n_emb = len([k for k in p if k.startswith('emb_conv.') and k.endswith('.weight')])
    G = p['dcn_pack.conv_offset.weight'].shape[0] // 27
    feat = []
    for i in range(T):
        x_ref = x[i:i + 1].repeat(T, 1, 1, 1)
        x_set = F_.conv2d(torch.cat([x, x_ref], 1), p['offset_conv.weight'], p['offset_conv.bias'], padding=1)
        o = F_.conv2d(x_set, p['dcn_pack.conv_offset.weight'], p['dcn_pack.conv_offset.bias'], padding=1)
        o1, o2, m = torch.chunk(o, 3, dim=1)
        x_dcn = modulated_deform_conv2d(x, torch.cat((o1, o2), 1), torch.sigmoid(m), p['dcn_pack.weight'], p['dcn_pack.bias'],
                                        1, 1, 1, 1, G)
        x_cor = x_dcn * x_ref
        for e in range(n_emb):
            x_cor = F_.conv2d(x_cor, p['emb_conv.%d.weight' % e], p['emb_conv.%d.bias' % e], padding=1)
        feat.append((torch.softmax(x_cor, 0) * x).sum(0, keepdim=True))
    out = torch.cat(feat, 0)
    return torch.relu(F_.conv2d(out, p['conv2.weight'], p['conv2.bias'], padding=1))
