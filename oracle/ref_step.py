"""oracle/ref_step.py -- TEST INFRASTRUCTURE ONLY.

One key-frame step of the SELSA + TemporalRoIAlign RoI head built from the reference's OWN modules, loaded unmodified by
``oracle/ref_shim.py`` (from /root/reference, or from the staged copy oracle/_ref/ on the GPU box):

  TemporalRoIAlign / SingleRoIExtractor   mmtracking/mmtrack/models/roi_heads/roi_extractors/*.py   (reference classes)
  SelsaAggregator                          mmtracking/mmtrack/models/aggregators/selsa_aggregator.py (reference class)
  multiclass_nms                           mmdetection/mmdet/core/post_processing/bbox_nms.py         (reference function)
  delta2bbox                               mmdetection/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py (reference function)

Only the ~25 lines of caller glue that would need all of mmdet to import are restated, each citing what it follows:
SelsaRoIHead._bbox_forward (mmtrack/models/roi_heads/selsa_roi_head.py:80-97), SelsaBBoxHead.forward's shared-FC loop
(roi_heads/bbox_heads/selsa_bbox_head.py:50-58,82-83) and BBoxHead.get_bboxes (mmdet/models/roi_heads/bbox_heads/bbox_head.py:
319-373).  mmcv's RoIAlign / batched_nms are torchvision's ops, on the CPU or -- for the
torch-eager GPU comparator -- on CUDA.

Used by ``bench.py --impl reference`` (host cores, ``cpu_baseline.kind = "reference"``) and by the ``eager_cuda_reference`` leg
of the GPU arm.  Never imported by the product package.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_shim


def available():
    return ref_shim.available()


class ReferenceSelsaRoIHead(nn.Module):
    """The reference's RoI head for one key frame (test path), parameters named as in its state_dict."""

    def __init__(self, in_channels=512, fc_out_channels=1024, num_shared_fcs=3, num_classes=30, temporal_roi_align=True,
                 num_attention_blocks=16, test_cfg=None):
        super().__init__()
        ns = ref_shim.load()
        rpn = ref_shim.load_rpn()
        self._multiclass_nms = ns.multiclass_nms
        self._delta2bbox = rpn.delta2bbox
        roi_layer = dict(type='RoIAlign', output_size=7, sampling_ratio=2)
        if temporal_roi_align:
            self.bbox_roi_extractor = ns.TemporalRoIAlign(num_most_similar_points=2, num_temporal_attention_blocks=4,
                                                          roi_layer=roi_layer, out_channels=in_channels, featmap_strides=[16])
        else:
            self.bbox_roi_extractor = ns.SingleRoIExtractor(roi_layer=roi_layer, out_channels=in_channels, featmap_strides=[16])
        head = nn.Module()
        head.shared_fcs = nn.ModuleList()
        last = in_channels * 49
        for _ in range(num_shared_fcs):
            head.shared_fcs.append(nn.Linear(last, fc_out_channels))
            last = fc_out_channels
        head.aggregator = nn.ModuleList([ns.SelsaAggregator(in_channels=fc_out_channels, num_attention_blocks=num_attention_blocks)
                                         for _ in range(num_shared_fcs)])
        head.fc_cls = nn.Linear(last, num_classes + 1)
        head.fc_reg = nn.Linear(last, 4 * num_classes)
        self.bbox_head = head
        self.test_cfg = test_cfg or dict(score_thr=0.0001, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100)

    @torch.no_grad()
    def step(self, x, ref_x, rois, ref_rois, img_shape, scale_factor=(1., 1., 1., 1.), rescale=False, return_all=False):
        """x [1,C,H,W], ref_x [T,C,H,W], rois [N,5], ref_rois [T*N,5] -> (det_bboxes [k,5], det_labels [k])."""
        ext, head = self.bbox_roi_extractor, self.bbox_head
        # SelsaRoIHead._bbox_forward, selsa_roi_head.py:83-93
        bbox_feats = ext((x,), rois, ref_feats=(ref_x,))
        ref_bbox_feats = ext((ref_x,), ref_rois)
        # SelsaBBoxHead.forward, selsa_bbox_head.py:50-58
        a, r = bbox_feats.flatten(1), ref_bbox_feats.flatten(1)
        for i, fc in enumerate(head.shared_fcs):
            a = fc(a)
            r = fc(r)
            a = a + head.aggregator[i](a, r)
            r = F.relu(r)
            a = F.relu(a)
        cls_score, bbox_pred = head.fc_cls(a), head.fc_reg(a)                      # :82-83
        # BBoxHead.get_bboxes (non-batch mode), bbox_head.py:319-373
        scores = F.softmax(cls_score, dim=-1)
        bboxes = self._delta2bbox(rois[:, 1:], bbox_pred, (0., 0., 0., 0.), (0.2, 0.2, 0.2, 0.2), img_shape)
        if rescale and bboxes.size(0) > 0:
            sf = bboxes.new_tensor(scale_factor)
            bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size()[0], -1)
        cfg = self.test_cfg
        dets, labels = self._multiclass_nms(bboxes, scores, cfg['score_thr'], dict(cfg['nms']), cfg['max_per_img'])
        if return_all:
            return dict(bbox_feats=bbox_feats, cls_score=cls_score, bbox_pred=bbox_pred, dets=dets, labels=labels)
        return dets, labels


class ReferenceFeatureLevelDetector(nn.Module):
    """The feature-level detectors' test path after the backbone, from the reference's own functions:

      FGFA  fgfa.py:275-283  flow_warp_feats(memo, flows); agg[num_left] = x; EmbedAggregator(x, agg)   (reference classes)
      DFF   dff.py:210-216   flow_warp_feats(key map, flow)
      then  RPNHead._get_bboxes (rpn_head.py:82-236, reference method), RoIAlign (torchvision == mmcv, Appendix A.1),
            Shared2FCBBoxHead.forward (convfc_bbox_head.py:172-182, restated: fc+relu x2, fc_cls, fc_reg), BBoxHead.get_bboxes
            with the reference's delta2bbox and multiclass_nms.
    The RPN / FlowNet convs are outside the hot path on both sides: their outputs (objectness / delta maps, flows) are inputs."""

    def __init__(self, in_channels=512, fc_out_channels=1024, num_classes=30, with_aggregator=True, test_cfg=None, rpn_cfg=None):
        super().__init__()
        ns = ref_shim.load()
        self._R = ref_shim.load_rpn()
        self._ns = ns
        if with_aggregator:
            self.aggregator = ns.EmbedAggregator(num_convs=1, channels=in_channels, kernel_size=3)
        self.roi_layer = ns.RoIAlign(output_size=7, spatial_scale=1 / 16, sampling_ratio=2)
        head = nn.Module()
        head.shared_fcs = nn.ModuleList([nn.Linear(in_channels * 49, fc_out_channels), nn.Linear(fc_out_channels, fc_out_channels)])
        head.fc_cls = nn.Linear(fc_out_channels, num_classes + 1)
        head.fc_reg = nn.Linear(fc_out_channels, 4 * num_classes)
        self.bbox_head = head
        self.test_cfg = test_cfg or dict(score_thr=0.0001, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100)
        self.rpn_cfg = rpn_cfg or dict(nms_pre=6000, nms_thr=0.7, max_per_img=300)
        rpn = object.__new__(self._R.RPNHead)
        nn.Module.__init__(rpn)
        rpn.use_sigmoid_cls = True
        rpn.bbox_coder = self._R.DeltaXYWHBBoxCoder(target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.))
        object.__setattr__(self, '_rpn', rpn)

    def _proposals(self, rpn_cls, rpn_reg, anchors, img_shape):
        import warnings
        C = self._R.ConfigDict
        cfg = C(nms_pre=self.rpn_cfg['nms_pre'], nms=C(type='nms', iou_threshold=self.rpn_cfg['nms_thr']),
                max_per_img=self.rpn_cfg['max_per_img'], min_bbox_size=0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return self._rpn._get_bboxes([rpn_cls], [rpn_reg], [anchors], [img_shape] * rpn_cls.shape[0], None, cfg)

    def _detect(self, feat, rpn_cls, rpn_reg, anchors, img_shape):
        props = self._proposals(rpn_cls, rpn_reg, anchors, img_shape)[0]
        rois = torch.cat([props.new_zeros(props.shape[0], 1), props[:, :4]], 1)
        a = self.roi_layer(feat.float(), rois).flatten(1)
        for fc in self.bbox_head.shared_fcs:
            a = F.relu(fc(a))
        cls_score, bbox_pred = self.bbox_head.fc_cls(a), self.bbox_head.fc_reg(a)
        scores = F.softmax(cls_score, dim=-1)
        bboxes = self._R.delta2bbox(rois[:, 1:], bbox_pred, (0., 0., 0., 0.), (0.2, 0.2, 0.2, 0.2), img_shape)
        cfg = self.test_cfg
        return self._ns.multiclass_nms(bboxes, scores, cfg['score_thr'], dict(cfg['nms']), cfg['max_per_img'])

    @torch.no_grad()
    def fgfa_step(self, x, memo, flows, num_left, rpn_cls, rpn_reg, anchors, img_shape):
        agg = self._ns.flow_warp_feats(memo, flows)
        agg[num_left] = x[0]
        feat = self.aggregator(x, agg)
        return self._detect(feat, rpn_cls, rpn_reg, anchors, img_shape)

    @torch.no_grad()
    def dff_step(self, key_feat, flow, rpn_cls, rpn_reg, anchors, img_shape):
        feat = key_feat if flow is None else self._ns.flow_warp_feats(key_feat, flow)
        return self._detect(feat, rpn_cls, rpn_reg, anchors, img_shape)
