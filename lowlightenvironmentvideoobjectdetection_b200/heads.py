"""Callers of the hot path, restated as thin host Python (they stay Python/torch in the reference too;
the FC layers remain nn.Linear / cuBLAS).  Inference only.

  SelsaBBoxHead  mmtracking/mmtrack/models/roi_heads/bbox_heads/selsa_bbox_head.py:8-84 on top of
                 mmdet ConvFCBBoxHead (mmdetection/mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:85-124)
                 and BBoxHead.get_bboxes (mmdetection/mmdet/models/roi_heads/bbox_heads/bbox_head.py:269-373)
  SelsaRoIHead   mmtracking/mmtrack/models/roi_heads/selsa_roi_head.py:80-97,147-187

Parameter names follow the reference state_dict: shared_fcs.{i}, aggregator.{i}.{fc_embed,ref_fc_embed,fc,ref_fc},
fc_cls, fc_reg, bbox_roi_extractor.embed_network.conv.

Reference-frame cache.  SELSA.extract_feats keeps a per-video memory of reference feature maps that does not
change between key frames (adaptive stride: the 14 memory frames are fixed for the whole video, only the key frame's own slot
changes; fixed stride: a FIFO that advances every ``frame_stride`` frames -- mmtracking/mmtrack/models/vid/selsa.py:207-249), yet
the reference re-extracts the 4500 reference RoIs (selsa_roi_head.py:88-89) and re-runs every shared FC and every K/V
projection on them (selsa_bbox_head.py:53-58, selsa_aggregator.py:55,64) for every key frame.  ``RefFrameCache`` keeps, per
reference frame, exactly what those calls produce -- the NHWC map with its norms and unit-norm bf16 copy (TemporalRoIAlign's
search space) and each layer's K rows and V^T columns -- so a key-frame step only computes them for the frames that are new.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
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
        if aggregator is not None:
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
    def ref_fc0(self, ref_x):
        """The first shared FC on the reference RoI features alone (selsa_bbox_head.py:51,55 for i = 0): independent of the
        key branch, so SelsaRoIHead runs it on a side stream next to TemporalRoIAlign's key path."""
        ref_x, ref_cl = self._flatten(ref_x)
        fc = self.shared_fcs[0]
        return F.linear(ref_x, self._fc0_weight(ref_cl), fc.bias)

    @torch.no_grad()
    def forward(self, x, ref_x, ref_fc0=None):
        """x [N, C, 7, 7] key RoI features, ref_x [M, C, 7, 7] reference RoI features
        -> (cls_score [N, classes+1], bbox_pred [N, 4*classes])   (selsa_bbox_head.py:25-84).
        ``ref_fc0``: ``self.ref_fc0(ref_x)`` when the caller already computed it (ref_x is then not touched)."""
        x, x_cl = self._flatten(x)
        if ref_fc0 is None:
            ref_x, ref_cl = self._flatten(ref_x)
        for i, fc in enumerate(self.shared_fcs):
            if i == 0:
                x = F.linear(x, self._fc0_weight(x_cl), fc.bias)
                ref_x = ref_fc0 if ref_fc0 is not None else F.linear(ref_x, self._fc0_weight(ref_cl), fc.bias)
            else:
                x = fc(x)
                ref_x = fc(ref_x)
            agg = self.aggregator[i]
            if self._fusable(x, ref_x):
                # :56-58 with the three element-wise steps (and the aggregator's output bias) as ONE launch; x and ref_x are
                # this layer's own FC outputs, so they are updated in place
                k, v, use_vt = agg.project_ref(ref_x)      # aggregator sees the PRE-ReLU features (:56)
                y = agg.attend(x, k, v, ref_x.shape[0], use_vt, with_bias=False)
                ops.selsa_residual_relu_(x, y, agg.out_bias(use_vt), ref_x)
            else:
                x = x + agg(x, ref_x)
                ref_x = F.relu(ref_x)
                x = F.relu(x)
        return self.fc_cls(x), self.fc_reg(x)

    @staticmethod
    def _fusable(x, ref_x):
        return (x.is_cuda and x.dtype == torch.float32 and ref_x.dtype == torch.float32 and x.is_contiguous()
                and ref_x.is_contiguous() and ref_x.shape[0] > 0 and x.shape[0] > 0 and x.shape[1] % 4 == 0)

    @torch.no_grad()
    def forward_cached(self, rows, n_key, cache, slots, channels_last=True):
        """SelsaBBoxHead.forward (selsa_bbox_head.py:25-84) with the reference side taken from ``cache``.

        rows [n_key + F*N, in_channels*49]: the flattened RoI features of the key-frame proposals (first ``n_key`` rows; may be
        0 when only the cache is being filled) followed by those of the F reference frames that go into ``slots`` (N each,
        in slot order).  Each layer: one FC over all rows, the new frames' K / V^T written into their cache slots, then the key
        rows attend over the cache's T*N reference proposals.  ``rows`` is consumed (activations are updated in place).
        Returns (cls_score, bbox_pred) of the key rows, or None when n_key == 0."""
        N, T = cache.N, cache.T
        F_new = len(slots)
        assert rows.shape[0] == n_key + F_new * N
        runs = _slot_runs(slots)
        y = rows
        for i, fc in enumerate(self.shared_fcs):
            w = self._fc0_weight(channels_last) if i == 0 else fc.weight
            y = self._linear_few_rows(y, w, fc.bias)                      # :53-55, key and new reference rows in one GEMM
            agg = self.aggregator[i]

            # (running these small GEMMs on a side stream next to the key rows' Q projection was measured: 547.8 vs 547.4 frames/s)
            for j0, s0, n_run in runs:                                    # :55,64 of selsa_aggregator.py, new frames only
                r = y[n_key + j0 * N:n_key + (j0 + n_run) * N]
                lo, hi = s0 * N, (s0 + n_run) * N
                if cache.v_transposed:
                    agg.project_ref(r, k_out=cache.K[i][lo:hi], vt_out=cache.V[i][:, lo:hi])
                else:
                    k, v, _ = agg.project_ref(r)
                    cache.K[i][lo:hi].copy_(k)
                    cache.V[i][lo:hi].copy_(v)
            if n_key and y.is_contiguous() and y.shape[1] % 4 == 0:
                x = y[:n_key]
                a = agg.attend(x, cache.K[i], cache.V[i], T * N, cache.v_transposed, with_bias=False)   # :56, pre-ReLU features on both sides
                ops.selsa_residual_relu_(x, a, agg.out_bias(cache.v_transposed), y[n_key:])              # :56-58 in one launch
            else:
                if n_key:
                    x = y[:n_key]
                    x += agg.attend(x, cache.K[i], cache.V[i], T * N, cache.v_transposed)
                if i + 1 < len(self.shared_fcs) or n_key:
                    torch.relu_(y)                                         # :57-58
        if not n_key:
            return None
        x = y[:n_key]
        return self.fc_cls(x), self.fc_reg(x)

    @staticmethod
    def _linear_few_rows(x, w, b, split=16):
        """F.linear for a handful of rows against a long reduction (fc_0: 600 x 25088 -> 1024).  The library tiles the
        [rows, out] plane only -- 40 CTAs on 148 SMs for 600 rows, each streaming 6 MB through its shared memory -- so the
        reduction axis is split into ``split`` slabs run as one batched GEMM (640 CTAs) and summed."""
        rows, k = x.shape
        while split > 1 and k % split:
            split //= 2
        if rows > 1024 or k < 8192 or split == 1 or not x.is_contiguous() or not w.is_contiguous():
            return F.linear(x, w, b)
        xs = x.view(rows, split, k // split).transpose(0, 1)              # [split, rows, k/split], no copy
        ws = w.view(w.shape[0], split, k // split).permute(1, 2, 0)       # [split, k/split, out], no copy
        y = torch.bmm(xs, ws).sum(dim=0)
        return y.add_(b) if b is not None else y

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
class Shared2FCBBoxHead(SelsaBBoxHead):
    """mmdet's Shared2FCBBoxHead (mmdetection/mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:85-124,172-182): the bbox
    head of the FGFA / DFF detectors -- ``num_shared_fcs`` FC + ReLU layers, then fc_cls / fc_reg.  A caller
    of the hot path like SelsaBBoxHead: FCs stay nn.Linear; ``get_bboxes`` / ``get_bboxes_device`` (fused decode + device NMS)
    are inherited.  State-dict keys: shared_fcs.{i}, fc_cls, fc_reg."""

    def __init__(self, num_shared_fcs=2, in_channels=512, fc_out_channels=1024, roi_feat_size=7, num_classes=30, **kwargs):
        kwargs.pop('aggregator', None)
        super().__init__(aggregator=None, num_shared_fcs=num_shared_fcs, in_channels=in_channels,
                         fc_out_channels=fc_out_channels, roi_feat_size=roi_feat_size, num_classes=num_classes, **kwargs)

    @torch.no_grad()
    def forward(self, x):
        x, cl = self._flatten(x)
        for i, fc in enumerate(self.shared_fcs):
            x = F.relu(F.linear(x, self._fc0_weight(cl) if i == 0 else fc.weight, fc.bias))
        return self.fc_cls(x), self.fc_reg(x)


@HEADS.register_module()
class StandardRoIHead(nn.Module):
    """mmdet StandardRoIHead, bbox branch, test path (mmdetection/mmdet/models/roi_heads/standard_roi_head.py:
    simple_test -> test_mixins.py:simple_test_bboxes): RoIAlign of the proposals on the (aggregated / warped) feature map,
    the bbox head, get_bboxes + multiclass NMS.  What FGFA and DFF run after their feature-level aggregation."""

    def __init__(self, bbox_roi_extractor, bbox_head, test_cfg=None, **kwargs):
        super().__init__()
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        head_cfg = dict(bbox_head)
        head_cfg.setdefault('type', 'Shared2FCBBoxHead')
        if 'bbox_coder' in head_cfg:
            coder = head_cfg.pop('bbox_coder')
            head_cfg.setdefault('target_means', coder.get('target_means', (0., 0., 0., 0.)))
            head_cfg.setdefault('target_stds', coder.get('target_stds', (0.2, 0.2, 0.2, 0.2)))
        self.bbox_head = HEADS.build(head_cfg)
        for layer in self.bbox_roi_extractor.roi_layers:
            layer.channels_last_out = True
        self.test_cfg = test_cfg or dict(score_thr=0.0001, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100)

    @torch.no_grad()
    def simple_test_device(self, x, rois, img_shape, scale_factor, rescale=False):
        """One image, fixed shapes, sync-free: (dets [max,5], labels [max], count [1])."""
        n_in = self.bbox_roi_extractor.num_inputs
        feats = self.bbox_roi_extractor(x[:n_in], rois)
        cls_score, bbox_pred = self.bbox_head(feats)
        return self.bbox_head.get_bboxes_device(rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=rescale,
                                                cfg=self.test_cfg)

    @torch.no_grad()
    def simple_test(self, x, proposal_list, img_metas, proposals=None, rescale=False):
        dets, labels = [], []
        for i, props in enumerate(proposal_list):
            rois = bbox2roi([props])
            if rois.shape[0] == 0:
                dets.append(rois.new_zeros((0, 5))); labels.append(rois.new_zeros((0,), dtype=torch.long))
                continue
            feat = tuple(f[i:i + 1] for f in x)
            d, l, c = self.simple_test_device(feat, rois, img_metas[i]['img_shape'], img_metas[i]['scale_factor'], rescale)
            n = int(c)
            dets.append(d[:n]); labels.append(l[:n])
        return dets, labels


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
        self.use_ref_cache = True     # simple_test(..., ref_img_metas=...) goes through the reference-frame cache
        self.overlap = True           # independent branches of a step run on two streams (see _bbox_forward)
        self._clip_cache = None
        # opt-in: the cached drop-in call replays one CUDA graph per key frame instead of launching ~50 kernels from Python
        # (simple_test(..., ref_img_metas=...) then runs at the speed of the captured step; see _cached_step)
        self.use_cuda_graphs = False
        self._step_graphs = {}

    @torch.no_grad()
    def _bbox_forward(self, x, ref_x, rois, ref_rois):
        """selsa_roi_head.py:80-97."""
        n_in = self.bbox_roi_extractor.num_inputs
        if self.overlap and ref_x[0].is_cuda and ref_rois.shape[0] > 0 and n_in == 1:
            # the reference branch (RoIAlign of the M reference RoIs + the first shared FC on them) is independent of the key
            # branch (TemporalRoIAlign over the key RoIs): it runs on a side stream next to it and joins before layer 0's
            # aggregator.  The layout pass of the reference maps is queued first so that both branches consume it.
            from . import ops
            from .roi_extractors import TemporalRoIAlign
            if isinstance(self.bbox_roi_extractor, TemporalRoIAlign):
                C = ref_x[0].shape[1]
                ops.to_nhwc(ref_x[0], want_norm=True, want_unit_bf16=(C % 64 == 0 and C <= 512 and
                                                                      self.bbox_roi_extractor.impl != ops.IMPL_SIMT))
            else:
                ops.to_nhwc(ref_x[0])
            with ops.fork(ref_x[0].device) as branch:
                ref_bbox_feats = self.bbox_roi_extractor(ref_x[:n_in], ref_rois)
                ref_pre = self.bbox_head.ref_fc0(ref_bbox_feats)
            bbox_feats = self.bbox_roi_extractor(x[:n_in], rois, ref_feats=ref_x[:n_in])
            branch.join()
            cls_score, bbox_pred = self.bbox_head(bbox_feats, None, ref_fc0=ref_pre)
        else:
            bbox_feats = self.bbox_roi_extractor(x[:n_in], rois, ref_feats=ref_x[:n_in])
            ref_bbox_feats = self.bbox_roi_extractor(ref_x[:n_in], ref_rois)
            cls_score, bbox_pred = self.bbox_head(bbox_feats, ref_bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)

    @torch.no_grad()
    def simple_test_bboxes(self, x, ref_x, proposals, ref_proposals, img_metas, rcnn_test_cfg, rescale=False):
        """selsa_roi_head.py:147-187."""
        rois = bbox2roi(proposals)
        ref_rois = bbox2roi(ref_proposals)
        if (self.use_cuda_graphs and len(proposals) == 1 and len(x) == 1 and len(ref_x) == 1 and x[0].is_cuda and rois.shape[0] > 0
                and ref_rois.shape[0] > 0 and rcnn_test_cfg is self.test_cfg and rcnn_test_cfg.get('max_per_img', 0) > 0
                and rcnn_test_cfg['nms'].get('type', 'nms') == 'nms' and not torch.cuda.is_current_stream_capturing()):
            dets, labels, count = self._graph_step(x[0], ref_x[0], rois, ref_rois, img_metas[0]['img_shape'],
                                                   img_metas[0]['scale_factor'], rescale)
            n = int(count)
            return [dets[:n]], [labels[:n]]
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

    @staticmethod
    def capture_callable(fn, warmup=2):
        """Captures ``fn()`` (any of the sync-free, fixed-shape entry points of this head over static tensors) into a CUDA graph:
        warm-up calls on a side stream first (workspaces, module loading, kernel attributes).  Returns (graph, fn's outputs)."""
        from . import ops
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        ops._nhwc_memo.clear()               # every layout pass must be recorded inside the graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = fn()
        ops._nhwc_memo.clear()
        return graph, outs

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

    # ------------------------------------------------------------------ reference-frame cache
    def new_ref_cache(self, num_frames, num_proposals, feat_shape, device=None):
        """A cache for clips whose reference set holds ``num_frames`` maps of ``feat_shape`` = (C, H, W) with
        ``num_proposals`` proposals each."""
        device = device if device is not None else next(self.parameters()).device
        return RefFrameCache(self, num_frames, num_proposals, feat_shape, device)

    def _ref_rows_into(self, cache, slots, feats, ref_rois, rows_out):
        """Layout pass + RoIAlign of the reference frames ``feats`` [F,C,H,W] that go into ``slots``; their RoI features are
        written into rows_out [F*N, 49*C] (channels_last rows).  ref_rois[:, 0] indexes ``feats``."""
        ext = self.bbox_roi_extractor
        layer = ext.roi_layers[0]
        from . import ops
        if cache.maps is not None:
            runs = _slot_runs(slots)
            for j0, s0, n_run in runs:
                hw = cache.H * cache.W
                ops._to_nhwc(feats[j0:j0 + n_run], out=(cache.maps[s0:s0 + n_run], cache.norm[s0 * hw:(s0 + n_run) * hw],
                                                        cache.unit[s0 * hw:(s0 + n_run) * hw] if cache.unit is not None else None))
            rois2 = ref_rois.clone()
            if len(runs) == 1:
                rois2[:, 0] += float(slots[0])            # one run of consecutive slots (sync-free: graph capturable)
            else:
                slot_of = torch.as_tensor(slots, dtype=ref_rois.dtype, device=ref_rois.device)
                rois2[:, 0] = slot_of[ref_rois[:, 0].long()]
            ops.roi_align_nhwc(cache.maps, rois2, layer.output_size, layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                               out_nhwc=True, out=rows_out)
        else:
            nhwc, _, _ = ops.to_nhwc(feats)
            ops.roi_align_nhwc(nhwc.contiguous(), ref_rois, layer.output_size, layer.spatial_scale, layer.sampling_ratio,
                               layer.aligned, out_nhwc=True, out=rows_out)

    @torch.no_grad()
    def update_ref_cache(self, cache, slots, feats, ref_rois, keys=None):
        """Computes everything the head needs from the reference frames ``feats`` [F,C,H,W] (their ``ref_rois`` [F*N,5],
        column 0 = index into ``feats``) and stores it in ``slots`` of the cache: what selsa_roi_head.py:88-89 and the
        reference side of selsa_bbox_head.py:53-58 recompute on every key frame.  Fixed shapes, no host synchronisation."""
        if not len(slots):
            return
        N = cache.N
        assert feats.shape[0] == len(slots) and ref_rois.shape[0] == len(slots) * N
        rows = torch.empty((len(slots) * N, cache.row_len), dtype=torch.float32, device=feats.device)
        self._ref_rows_into(cache, slots, feats, ref_rois, rows)
        self.bbox_head.forward_cached(rows, 0, cache, slots)
        cache.mark(slots, keys)

    @torch.no_grad()
    def simple_test_cached_device(self, x, rois, key_ref_rois, cache, key_slot, img_shape, scale_factor, rescale=False,
                                  return_feats=False):
        """One key frame against a filled cache; fixed shapes, no host synchronisation (CUDA-graph capturable).

        x: the key frame's feature tuple ([1,C,H,W]); rois [N,5] its proposals; the key frame is itself a member of the reference
        set (selsa.py:220-223): it occupies ``key_slot`` with the reference proposals ``key_ref_rois`` [N,5] (column 0 ignored;
        in the reference's pipeline these equal ``rois``).  Same outputs as ``simple_test_device`` on the full reference set."""
        from . import ops
        ext = self.bbox_roi_extractor
        layer = ext.roi_layers[0]
        N = cache.N
        assert rois.shape[0] > 0 and key_ref_rois.shape[0] == N
        n_key = rois.shape[0]
        feat = x[0]
        rows = torch.empty((n_key + N, cache.row_len), dtype=torch.float32, device=feat.device)
        zero_rois = key_ref_rois.clone()
        zero_rois[:, 0] = 0
        if cache.maps is not None:
            hw = cache.H * cache.W
            s = key_slot
            ops._to_nhwc(feat, out=(cache.maps[s:s + 1], cache.norm[s * hw:(s + 1) * hw],
                                    cache.unit[s * hw:(s + 1) * hw] if cache.unit is not None else None))
            key_nhwc = cache.maps[s:s + 1]
            ext.forward_from_layout(key_nhwc, rois, cache.maps, cache.norm, cache.unit, out=rows[:n_key])
        else:
            key_nhwc = ops.to_nhwc(feat)[0].contiguous()
            ops.roi_align_nhwc(key_nhwc, rois, layer.output_size, layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                               out_nhwc=True, out=rows[:n_key])
        ops.roi_align_nhwc(key_nhwc, zero_rois, layer.output_size, layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                           out_nhwc=True, out=rows[n_key:])
        feats_out = rows[:n_key].clone() if return_feats else None
        cache.mark([key_slot], None)
        assert cache.filled(), 'reference-frame cache has unfilled slots'
        cls_score, bbox_pred = self.bbox_head.forward_cached(rows, n_key, cache, [key_slot])
        out = self.bbox_head.get_bboxes_device(rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=rescale, cfg=self.test_cfg)
        if return_feats:
            return out + (dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=feats_out),)
        return out

    def _simple_test_with_cache(self, x, ref_x, proposals_list, ref_proposals_list, img_metas, ref_img_metas, rescale):
        """simple_test through the reference-frame cache: frames are identified by their img_metas, the cache is brought into the
        order of ``ref_img_metas`` (a FIFO advance is a shift), and only frames it does not hold are computed.  Returns None
        when the situation is outside the cache's fixed-shape contract (the caller then runs the uncached path)."""
        ext = self.bbox_roi_extractor
        if len(x) != 1 or len(ref_x) != 1 or len(proposals_list) != 1 or len(ext.featmap_strides) != 1:
            return None
        T = ref_x[0].shape[0]
        N = ref_proposals_list[0].shape[0] if len(ref_proposals_list) else 0
        if N == 0 or len(ref_proposals_list) != T or len(ref_img_metas) != T or proposals_list[0].shape[0] == 0 or \
                any(p.shape[0] != N for p in ref_proposals_list):
            return None
        keys = [_frame_key(m) for m in ref_img_metas]
        key_key = _frame_key(img_metas[0])
        if len(set(keys)) != T or key_key not in keys:
            return None
        key_slot = keys.index(key_key)
        C, H, W = ref_x[0].shape[1:]
        cache = getattr(self, '_clip_cache', None)
        if cache is None or not cache.compatible(self, T, N, (C, H, W), ref_x[0].device):
            cache = self._clip_cache = self.new_ref_cache(T, N, (C, H, W), ref_x[0].device)
        cache.align(keys)
        todo = [t for t in range(T) if t != key_slot and cache.keys[t] != keys[t]]
        if todo:
            idx = torch.as_tensor(todo, device=ref_x[0].device)
            rr = bbox2roi([ref_proposals_list[t] for t in todo])
            self.update_ref_cache(cache, todo, ref_x[0].index_select(0, idx), rr, keys=[keys[t] for t in todo])
        rois = bbox2roi(proposals_list)
        dets, labels, count = self._cached_step(
            x, rois, bbox2roi([ref_proposals_list[key_slot]]), cache, key_slot, img_metas[0]['img_shape'],
            img_metas[0]['scale_factor'], rescale)
        cache.keys[key_slot] = key_key
        n = int(count)
        return [dets[:n]], [labels[:n]]

    def _graph_step(self, feat, ref_feat, rois, ref_rois, img_shape, scale_factor, rescale):
        """The uncached step (``simple_test_device``) as a replay of the graph captured for these shapes: the caller's maps and
        RoIs are copied into the graph's static inputs, the outputs are copies of its static outputs (``use_cuda_graphs``)."""
        key = ('uncached', tuple(feat.shape), tuple(ref_feat.shape), feat.dtype, rois.shape[0], ref_rois.shape[0], tuple(img_shape),
               tuple(scale_factor), bool(rescale), torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32,
               torch.cuda.current_device(), RefFrameCache._weights_tag(self))
        entry = self._step_graphs.get(key)
        if entry is None:
            if len(self._step_graphs) >= 4:
                self._step_graphs.pop(next(iter(self._step_graphs)))
            st = [t.clone() for t in (feat, ref_feat, rois, ref_rois)]
            graph, outs = self.capture_graph((st[0],), (st[1],), st[2], st[3], img_shape, scale_factor, rescale)
            entry = self._step_graphs[key] = dict(cache=None, graph=graph, outs=outs, st=st)
        for dst, src in zip(entry['st'], (feat, ref_feat, rois, ref_rois)):
            dst.copy_(src)
        entry['graph'].replay()
        return tuple(o.clone() for o in entry['outs'])

    def _cached_step(self, x, rois, key_ref_rois, cache, key_slot, img_shape, scale_factor, rescale):
        """``simple_test_cached_device``, eagerly or -- with ``use_cuda_graphs`` -- as a replay of the step captured for this
        (cache, key slot, proposal count, image geometry, library-math setting): the caller's tensors are copied into the graph's
        static inputs (the key map is 4.9 MB at R-50-DC5 size), the outputs are copies of its static outputs."""
        feat = x[0]
        if not (self.use_cuda_graphs and feat.is_cuda) or torch.cuda.is_current_stream_capturing():
            return self.simple_test_cached_device(x, rois, key_ref_rois, cache, key_slot, img_shape, scale_factor, rescale=rescale)
        key = (id(cache), key_slot, tuple(feat.shape), feat.dtype, rois.shape[0], tuple(img_shape), tuple(scale_factor), bool(rescale),
               torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.cuda.current_device())
        entry = self._step_graphs.get(key)
        if entry is None or entry['cache'] is not cache:
            if len(self._step_graphs) >= 4:                    # a handful of shapes at most: drop the oldest
                self._step_graphs.pop(next(iter(self._step_graphs)))
            st_x, st_rois, st_kr = feat.clone(), rois.clone(), key_ref_rois.clone()
            graph, outs = self.capture_callable(lambda: self.simple_test_cached_device(
                (st_x,), st_rois, st_kr, cache, key_slot, img_shape, scale_factor, rescale=rescale))
            entry = self._step_graphs[key] = dict(cache=cache, graph=graph, outs=outs, x=st_x, rois=st_rois, kr=st_kr)
        entry['x'].copy_(feat)
        entry['rois'].copy_(rois)
        entry['kr'].copy_(key_ref_rois)
        entry['graph'].replay()
        cache.mark([key_slot], None)                           # the host-side bookkeeping the captured call did at capture time
        return tuple(o.clone() for o in entry['outs'])

    @torch.no_grad()
    def simple_test(self, x, ref_x, proposals_list, ref_proposals_list, img_metas, proposals=None, rescale=False,
                    ref_img_metas=None):
        """selsa_roi_head.py:115-145; returns (det_bboxes, det_labels) lists (device tensors; the
        per-class numpy split of bbox2result is left to the caller).

        ``ref_img_metas`` (optional, what SELSA.simple_test already holds, selsa.py:309-312): identifies the reference frames so
        that everything that depends only on a reference frame is taken from the per-clip cache instead of being recomputed."""
        if ref_img_metas is not None and self.use_ref_cache:
            out = self._simple_test_with_cache(x, ref_x, proposals_list, ref_proposals_list, img_metas, ref_img_metas, rescale)
            if out is not None:
                return out
        return self.simple_test_bboxes(x, ref_x, proposals_list, ref_proposals_list, img_metas, self.test_cfg,
                                       rescale=rescale)


def _slot_runs(slots):
    """[(first position in ``slots``, first slot, run length)] of maximal runs of consecutive slots."""
    runs, j = [], 0
    while j < len(slots):
        n = 1
        while j + n < len(slots) and slots[j + n] == slots[j] + n:
            n += 1
        runs.append((j, slots[j], n))
        j += n
    return runs


def _frame_key(meta):
    """Identity of a frame from its img_meta (mmtrack's VideoCollect keys): the file when known, else (video, frame id)."""
    ident = meta.get('filename', None)
    if ident is None:
        ident = (meta.get('video_id', None), meta.get('frame_id', None))
    return (ident, tuple(meta.get('img_shape', ())), bool(meta.get('flip', False)))


class RefFrameCache:
    """Per-clip cache of everything that depends only on a reference frame (see the module docstring), in the ORDER of the
    reference set, so that every reduction over frames / reference proposals runs in the same order as in the uncached step.

      maps [T,H,W,C] fp32, norm [T*H*W], unit [T*H*W, C] bf16   TemporalRoIAlign's search space (only with that extractor)
      K[i] [T*N, D], V[i] = V^T [D, ld] (or V [T*N, D])          aggregator layer i's reference-side projections
    A FIFO advance of the reference set (selsa.py:237-243) is ``shift``: slots move down, in place."""

    def __init__(self, head, T, N, feat_shape, device):
        from .roi_extractors import TemporalRoIAlign
        C, H, W = feat_shape
        ext, bh = head.bbox_roi_extractor, head.bbox_head
        self.T, self.N, self.C, self.H, self.W = T, N, C, H, W
        self.device = torch.device(device)
        ph, pw = ext.roi_layers[0].output_size
        self.row_len = C * ph * pw
        self.keys = [None] * T
        self._filled = [False] * T
        self.maps = self.norm = self.unit = None
        if isinstance(ext, TemporalRoIAlign):
            self.maps = torch.empty((T, H, W, C), dtype=torch.float32, device=device)
            self.norm = torch.empty((T * H * W,), dtype=torch.float32, device=device)
            if C % 64 == 0 and C <= 512:
                self.unit = torch.empty((T * H * W + 4, C), dtype=torch.bfloat16, device=device)[:T * H * W]
        D = bh.shared_fcs[0].out_features
        self.v_transposed, align = bh.aggregator[0].v_layout()
        L = len(bh.shared_fcs)
        self.K = [torch.empty((T * N, D), dtype=torch.float32, device=device) for _ in range(L)]
        if self.v_transposed:
            ld = (T * N + align - 1) // align * align
            self.V = [torch.zeros((D, ld), dtype=torch.float32, device=device) for _ in range(L)]
        else:
            self.V = [torch.empty((T * N, D), dtype=torch.float32, device=device) for _ in range(L)]
        self._tag = self._weights_tag(head)

    @staticmethod
    def _weights_tag(head):
        return tuple((p.data_ptr(), p._version) for p in head.parameters())

    def compatible(self, head, T, N, feat_shape, device):
        return (self.T, self.N, (self.C, self.H, self.W)) == (T, N, tuple(feat_shape)) and self.device == torch.device(device) \
            and self._tag == self._weights_tag(head)

    def filled(self):
        return all(self._filled)

    def mark(self, slots, keys):
        for j, s in enumerate(slots):
            self._filled[s] = True
            self.keys[s] = keys[j] if keys is not None else self.keys[s]

    def reset(self):
        self.keys = [None] * self.T
        self._filled = [False] * self.T

    def shift(self, s):
        """Slots s..T-1 move to 0..T-1-s (the oldest ``s`` frames leave); the last ``s`` slots become unfilled."""
        if s <= 0:
            return
        T, N = self.T, self.N
        if s < T:
            hw = self.H * self.W

            def move(buf, unit, dim=0):
                src = buf.narrow(dim, s * unit, (T - s) * unit).clone()
                buf.narrow(dim, 0, (T - s) * unit).copy_(src)
            if self.maps is not None:
                move(self.maps, 1)
                move(self.norm, hw)
                if self.unit is not None:
                    move(self.unit, hw)
            for i in range(len(self.K)):
                move(self.K[i], N)
                move(self.V[i], N, dim=1 if self.v_transposed else 0)
        self.keys = self.keys[s:] + [None] * min(s, T)
        self._filled = self._filled[s:] + [False] * min(s, T)
        self.keys, self._filled = self.keys[:T], self._filled[:T]

    def align(self, keys):
        """Brings the cache into the order of ``keys``: the shift that preserves the most frames (0 = none); slots whose frame
        is not the wanted one are left for the caller to recompute."""
        best, best_s = -1, 0
        for s in range(self.T + 1):
            hits = sum(1 for t in range(self.T - s) if self.keys[t + s] is not None and self.keys[t + s] == keys[t])
            if hits > best:
                best, best_s = hits, s
        if best <= 0:
            self.reset()
            return
        self.shift(best_s)
        for t in range(self.T):
            if self.keys[t] != keys[t]:
                self._filled[t] = False
