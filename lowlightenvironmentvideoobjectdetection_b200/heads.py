"""Callers of the hot path, restated as thin host Python (they stay Python/torch in the reference too;
the FC layers remain nn.Linear / cuBLAS).  Inference only.

  SelsaBBoxHead  mmtracking/mmtrack/models/roi_heads/bbox_heads/selsa_bbox_head.py:8-84 on top of
                 mmdet ConvFCBBoxHead (mmdetection/mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:85-124)
                 and BBoxHead.get_bboxes (mmdetection/mmdet/models/roi_heads/bbox_heads/bbox_head.py:269-373)
  SelsaRoIHead   mmtracking/mmtrack/models/roi_heads/selsa_roi_head.py:80-97,147-187

Parameter names follow the reference state_dict: shared_fcs.{i}, aggregator.{i}.{fc_embed,ref_fc_embed,fc,ref_fc},
fc_cls, fc_reg, bbox_roi_extractor.embed_network.conv.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .post_processing import bbox2roi, delta2bbox, multiclass_nms
from .registry import HEADS, build_aggregator, build_roi_extractor


@HEADS.register_module()
class SelsaBBoxHead(nn.Module):
    """Shared-FC bbox head with a SELSA aggregator after every shared FC (no shared convs / avg pool,
    as in every config of the reference)."""

    def __init__(self, aggregator, num_shared_fcs=2, in_channels=512, fc_out_channels=1024, roi_feat_size=7,
                 num_classes=30, reg_class_agnostic=False, target_means=(0., 0., 0., 0.),
                 target_stds=(0.2, 0.2, 0.2, 0.2), **kwargs):
        super().__init__()
        self.num_shared_fcs = num_shared_fcs
        self._in_channels = in_channels
        self.num_classes = num_classes
        self.reg_class_agnostic = reg_class_agnostic
        self.target_means, self.target_stds = tuple(target_means), tuple(target_stds)
        self.shared_fcs = nn.ModuleList()
        last = in_channels * roi_feat_size * roi_feat_size
        for _ in range(num_shared_fcs):
            self.shared_fcs.append(nn.Linear(last, fc_out_channels))
            last = fc_out_channels
        self.fc_cls = nn.Linear(last, num_classes + 1)
        self.fc_reg = nn.Linear(last, 4 if reg_class_agnostic else 4 * num_classes)
        self.aggregator = nn.ModuleList([build_aggregator(aggregator) for _ in range(num_shared_fcs)])
        self._fc0_cl = None   # (weight version, weight columns re-ordered for channels_last RoI features)
        self.init_weights()

    def init_weights(self):
        """mmdet defaults: xavier for shared fcs, N(0, 0.01) for fc_cls, N(0, 0.001) for fc_reg."""
        for fc in self.shared_fcs:
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)
        nn.init.normal_(self.fc_cls.weight, 0, 0.01)
        nn.init.constant_(self.fc_cls.bias, 0)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001)
        nn.init.constant_(self.fc_reg.bias, 0)

    @torch.no_grad()
    def forward(self, x, ref_x):
        """x [N, C, 7, 7] key RoI features, ref_x [M, C, 7, 7] reference RoI features
        -> (cls_score [N, classes+1], bbox_pred [N, 4*classes])   (selsa_bbox_head.py:25-84)."""
        x, x_cl = self._flatten(x)
        ref_x, ref_cl = self._flatten(ref_x)
        for i, fc in enumerate(self.shared_fcs):
            if i == 0:
                x = F.linear(x, self._fc0_weight(x_cl), fc.bias)
                ref_x = F.linear(ref_x, self._fc0_weight(ref_cl), fc.bias)
            else:
                x = fc(x)
                ref_x = fc(ref_x)
            x = x + self.aggregator[i](x, ref_x)   # aggregator sees the PRE-ReLU features (:56)
            ref_x = F.relu(ref_x)
            x = F.relu(x)
        return self.fc_cls(x), self.fc_reg(x)

    @staticmethod
    def _flatten(t):
        """x.flatten(1) (selsa_bbox_head.py:50-51).  RoI features that arrive with channels_last strides
        (memory [K,7,7,C], what the RoIAlign / TAFA kernels write without a transposition tile) are flattened
        in (7,7,C) order for free; fc_0 then uses its column-permuted weight, so the product is unchanged."""
        if t.dim() == 4 and not t.is_contiguous() and t.permute(0, 2, 3, 1).is_contiguous():
            return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), True
        return t.flatten(1), False

    def _fc0_weight(self, channels_last):
        w = self.shared_fcs[0].weight
        if not channels_last:
            return w
        key = (w._version, w.data_ptr())
        if self._fc0_cl is None or self._fc0_cl[0] != key:
            out_f, in_f = w.shape
            p = in_f // self._in_channels
            w_cl = w.detach().view(out_f, self._in_channels, p).permute(0, 2, 1).reshape(out_f, in_f).contiguous()
            self._fc0_cl = (key, w_cl)
        return self._fc0_cl[1]

    @torch.no_grad()
    def get_bboxes(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False, cfg=None):
        """bbox_head.py:269-373 (non-batch mode)."""
        if (cfg is not None and cls_score.is_cuda and cfg['nms'].get('type', 'nms') == 'nms' and cls_score.size(0) > 0
                and cfg.get('max_per_img', 0) > 0):
            # same result through the fused decode kernel + device NMS; one host read (the detection count) trims the
            # fixed-size buffers to the reference's variable-length return
            dets, labels, count = self.get_bboxes_device(rois, cls_score, bbox_pred, img_shape, scale_factor, rescale, cfg)
            n = int(count)
            return dets[:n], labels[:n]
        scores = F.softmax(cls_score, dim=-1)
        bboxes = delta2bbox(rois[:, 1:], bbox_pred, self.target_means, self.target_stds, max_shape=img_shape)
        if rescale and bboxes.size(0) > 0:
            sf = bboxes.new_tensor(scale_factor)
            bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size(0), -1)
        if cfg is None:
            return bboxes, scores
        return multiclass_nms(bboxes, scores, cfg['score_thr'], cfg['nms'], cfg['max_per_img'])

    @torch.no_grad()
    def get_bboxes_device(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False, cfg=None):
        """Same result as ``get_bboxes`` but sync-free and fixed-shape: softmax + delta2bbox + clip + rescale +
        candidate expansion run as one kernel, the NMS (sort, mask, sweep, gather) stays on the device.
        Returns (dets [max_per_img,5] zero-padded, labels [max_per_img], count [1] int32 device tensor)."""
        from . import ops
        nms_cfg = dict(cfg['nms'])
        assert nms_cfg.pop('type', 'nms') == 'nms'
        thr = nms_cfg.pop('iou_threshold', nms_cfg.pop('iou_thr', None))
        cand = ops.bbox_decode_candidates(rois, cls_score, bbox_pred, self.num_classes, self.target_means,
                                          self.target_stds, img_shape, scale_factor if rescale else None,
                                          cfg['score_thr'], reg_class_agnostic=self.reg_class_agnostic)
        return ops.multiclass_nms_device(*cand, thr, cfg['max_per_img'], split_thr=nms_cfg.pop('split_thr', 10000),
                                         class_agnostic=nms_cfg.pop('class_agnostic', False))


@HEADS.register_module()
class SelsaRoIHead(nn.Module):
    """selsa roi head (bbox branch, test path)."""

    def __init__(self, bbox_roi_extractor, bbox_head, test_cfg=None, **kwargs):
        super().__init__()
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        head_cfg = dict(bbox_head)
        head_cfg.setdefault('type', 'SelsaBBoxHead')
        if 'bbox_coder' in head_cfg:
            coder = head_cfg.pop('bbox_coder')
            head_cfg.setdefault('target_means', coder.get('target_means', (0., 0., 0., 0.)))
            head_cfg.setdefault('target_stds', coder.get('target_stds', (0.2, 0.2, 0.2, 0.2)))
        self.bbox_head = HEADS.build(head_cfg)
        # our bbox head consumes channels_last RoI features directly (fc_0 weight columns permuted once), so the
        # extractor kernels can skip the [C][49] shared-memory transposition
        for layer in self.bbox_roi_extractor.roi_layers:
            layer.channels_last_out = True
        self.test_cfg = test_cfg or dict(score_thr=0.0001, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100)

    @torch.no_grad()
    def _bbox_forward(self, x, ref_x, rois, ref_rois):
        """selsa_roi_head.py:80-97."""
        n_in = self.bbox_roi_extractor.num_inputs
        bbox_feats = self.bbox_roi_extractor(x[:n_in], rois, ref_feats=ref_x[:n_in])
        ref_bbox_feats = self.bbox_roi_extractor(ref_x[:n_in], ref_rois)
        cls_score, bbox_pred = self.bbox_head(bbox_feats, ref_bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)

    @torch.no_grad()
    def simple_test_bboxes(self, x, ref_x, proposals, ref_proposals, img_metas, rcnn_test_cfg, rescale=False):
        """selsa_roi_head.py:147-187."""
        rois = bbox2roi(proposals)
        ref_rois = bbox2roi(ref_proposals)
        bbox_results = self._bbox_forward(x, ref_x, rois, ref_rois)
        img_shapes = tuple(meta['img_shape'] for meta in img_metas)
        scale_factors = tuple(meta['scale_factor'] for meta in img_metas)
        cls_score, bbox_pred = bbox_results['cls_score'], bbox_results['bbox_pred']
        num_per_img = tuple(len(p) for p in proposals)
        rois = rois.split(num_per_img, 0)
        cls_score = cls_score.split(num_per_img, 0)
        bbox_pred = bbox_pred.split(num_per_img, 0)
        det_bboxes, det_labels = [], []
        for i in range(len(proposals)):
            det_bbox, det_label = self.bbox_head.get_bboxes(rois[i], cls_score[i], bbox_pred[i], img_shapes[i],
                                                            scale_factors[i], rescale=rescale, cfg=rcnn_test_cfg)
            det_bboxes.append(det_bbox)
            det_labels.append(det_label)
        return det_bboxes, det_labels

    @torch.no_grad()
    def simple_test_device(self, x, ref_x, rois, ref_rois, img_shape, scale_factor, rescale=False):
        """One key frame, fixed shapes, no host synchronisation (the form CUDA graphs capture):
        x / ref_x feature tuples, rois [N,5] of the key frame, ref_rois [M,5] -> (dets [max,5], labels [max], count [1])."""
        res = self._bbox_forward(x, ref_x, rois, ref_rois)
        return self.bbox_head.get_bboxes_device(rois, res['cls_score'], res['bbox_pred'], img_shape, scale_factor,
                                                rescale=rescale, cfg=self.test_cfg)

    def capture_graph(self, x, ref_x, rois, ref_rois, img_shape, scale_factor, rescale=False, warmup=2):
        """Captures ``simple_test_device`` for these (static) input tensors into a CUDA graph.  The caller refills
        the same input tensors in place and calls ``graph.replay()``; outputs are the returned static tensors."""
        from . import ops
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):          # allocates workspaces, loads modules, sets kernel attributes
                self.simple_test_device(x, ref_x, rois, ref_rois, img_shape, scale_factor, rescale)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        ops._nhwc_memo.clear()               # every layout pass must be recorded inside the graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = self.simple_test_device(x, ref_x, rois, ref_rois, img_shape, scale_factor, rescale)
        ops._nhwc_memo.clear()
        return graph, outs

    @torch.no_grad()
    def simple_test(self, x, ref_x, proposals_list, ref_proposals_list, img_metas, proposals=None, rescale=False):
        """selsa_roi_head.py:115-145; returns (det_bboxes, det_labels) lists (device tensors; the
        per-class numpy split of bbox2result is left to the caller)."""
        return self.simple_test_bboxes(x, ref_x, proposals_list, ref_proposals_list, img_metas, self.test_cfg,
                                       rescale=rescale)
