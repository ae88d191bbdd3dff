"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY.

Loads the reference's own hot-path Python files UNMODIFIED -- from
``/root/reference`` in the build container, from the byte-identical staged copy
``oracle/_ref/`` (``oracle/make_ref.py``; git-ignored, travels with gpurun) on the
GPU box -- under small stand-ins for the absent mmcv / mmdet packages
(SURVEY.md Appendix C).  Used to (i) validate the restatement in
``oracle/vod_oracle.py``, (ii) generate the golden vectors committed under
``tests/golden/`` (``tests/golden/make_golden.py``) and (iii) run the reference's
own step as the benchmark comparator (``oracle/ref_step.py``).

Stand-ins (the only arithmetic they carry is the un-vendored mmcv-full ops):
  mmcv.ops.RoIAlign        -> torchvision.ops.roi_align(aligned=True)   (Appendix A.1)
  mmcv.ops.nms.batched_nms -> offset trick + torchvision.ops.nms        (Appendix A.5)
  mmcv.cnn.ConvModule      -> nn.Conv2d (+ReLU), sub-module named .conv
  mmcv.runner.force_fp32   -> identity decorator
  mmcv.utils.Registry      -> name -> class table with register_module/build
  mmcv.ops.modulated_deform_conv2d / ModulatedDeformConv2d -> torchvision.ops.deform_conv2d(mask=...) (same DCN-derived
                              offset layout and border rule; parameters ``weight`` / ``bias`` as mmcv's module)
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

def _default_root():
    if 'VOD_REFERENCE_ROOT' in os.environ:
        return os.environ['VOD_REFERENCE_ROOT']
    if os.path.isdir('/root/reference/mmtracking/mmtrack'):
        return '/root/reference'
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


REF_ROOT = _default_root()
_LOADED = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmtracking', 'mmtrack'))


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError('%s already registered in %s' % (key, self.name))
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    cls = registry.get(args.pop('type'))
    return cls(**args)


class _ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), **kw):
        super().__init__()
        assert norm_cfg is None
        conv_type = (conv_cfg or {}).get('type', 'Conv2d')
        layer = nn.ConvTranspose2d if conv_type == 'deconv' else nn.Conv2d          # mmcv: 'deconv' -> nn.ConvTranspose2d
        self.conv = layer(in_channels, out_channels, kernel_size, stride=stride,
                          padding=padding, bias=(bias == 'auto' or bool(bias)))
        self.with_activation = act_cfg is not None
        if self.with_activation:
            if act_cfg.get('type', 'ReLU') == 'LeakyReLU':
                self.activate = nn.LeakyReLU(negative_slope=act_cfg.get('negative_slope', 0.01), inplace=False)
            else:
                self.activate = nn.ReLU(inplace=False)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def _force_fp32(apply_to=None, out_fp16=False):
    def deco(fn):
        return fn
    return deco


class _RoIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) \
            else tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        assert pool_mode == 'avg'
        self.aligned = aligned

    def forward(self, input, rois):
        from torchvision.ops import roi_align
        return roi_align(input, rois, self.output_size, self.spatial_scale,
                         self.sampling_ratio, self.aligned)


def _batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    from torchvision.ops import nms as tv_nms
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + 1)
        boxes_for_nms = boxes + offsets[:, None]
    nms_cfg_.pop('type', 'nms')
    split_thr = nms_cfg_.pop('split_thr', 10000)
    thr = nms_cfg_.pop('iou_threshold', nms_cfg_.pop('iou_thr', None))
    if boxes_for_nms.shape[0] < split_thr:
        keep = tv_nms(boxes_for_nms, scores, thr)
        boxes = boxes[keep]
        scores = scores[keep]
    else:
        total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
        for id in torch.unique(idxs):
            mask = (idxs == id).nonzero(as_tuple=False).view(-1)
            keep = tv_nms(boxes_for_nms[mask], scores[mask], thr)
            total_mask[mask[keep]] = True
        keep = total_mask.nonzero(as_tuple=False).view(-1)
        keep = keep[scores[keep].argsort(descending=True)]
        boxes = boxes[keep]
        scores = scores[keep]
    return torch.cat([boxes, scores[:, None]], -1), keep


def _modulated_deform_conv2d(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deform_groups=1):
    from torchvision.ops import deform_conv2d
    assert groups == 1 and offset.shape[1] == 2 * deform_groups * weight.shape[2] * weight.shape[3]
    return deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask)


class _ModulatedDeformConv2d(nn.Module):
    """mmcv.ops.ModulatedDeformConv2d's constructor contract (mmcv-full 1.2.x): attributes read by the reference's
    ModulatedDCNPack (in_channels, kernel_size pair, stride, padding, deform_groups, ...), parameters ``weight`` / ``bias``."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deform_groups=1, bias=True):
        super().__init__()
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.Tensor(out_channels)) if bias else None
        n = in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / n ** 0.5
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()


def _constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []  # behave as a package so dotted children resolve
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load(dotted, rel_path):
    path = os.path.join(REF_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(dotted, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = module
    spec.loader.exec_module(module)
    return module


def load():
    """Returns a namespace with the reference's own classes/functions."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    for name in list(sys.modules):
        if name.split('.')[0] in ('mmcv', 'mmdet', 'mmtrack'):
            raise RuntimeError('real %s already imported; shim refuses to shadow it' % name)

    ops_nms = _mod('mmcv.ops.nms', batched_nms=_batched_nms)
    ops = _mod('mmcv.ops', RoIAlign=_RoIAlign, nms=ops_nms, batched_nms=_batched_nms,
               ModulatedDeformConv2d=_ModulatedDeformConv2d, modulated_deform_conv2d=_modulated_deform_conv2d)
    utils = _mod('mmcv.utils', Registry=_Registry, build_from_cfg=_build_from_cfg)
    bricks = _mod('mmcv.cnn.bricks', ConvModule=_ConvModule)
    cnn = _mod('mmcv.cnn', ConvModule=_ConvModule, bricks=bricks, constant_init=_constant_init)
    runner = _mod('mmcv.runner', force_fp32=_force_fp32)
    _mod('mmcv', ops=ops, utils=utils, cnn=cnn, runner=runner)

    ROI_EXTRACTORS = _Registry('roi_extractor')
    AGGREGATORS = _Registry('aggregator')
    _mod('mmdet')
    _mod('mmdet.core')
    _mod('mmdet.core.bbox')
    _mod('mmdet.core.bbox.iou_calculators', bbox_overlaps=None)
    _mod('mmdet.core.post_processing')
    _mod('mmdet.models')
    _mod('mmdet.models.builder', ROI_EXTRACTORS=ROI_EXTRACTORS)
    _mod('mmdet.models.roi_heads')
    rx = _mod('mmdet.models.roi_heads.roi_extractors')
    _mod('mmtrack')
    _mod('mmtrack.core')
    _mod('mmtrack.core.motion')
    _mod('mmtrack.models')
    MOTION = _Registry('motion')
    _mod('mmtrack.models.builder', AGGREGATORS=AGGREGATORS, MOTION=MOTION)
    _mod('mmtrack.models.motion')
    _mod('mmtrack.models.aggregators')
    _mod('mmtrack.models.roi_heads')
    _mod('mmtrack.models.roi_heads.roi_extractors')

    d = 'mmdetection/mmdet/'
    t = 'mmtracking/mmtrack/'
    base = _load('mmdet.models.roi_heads.roi_extractors.base_roi_extractor',
                 d + 'models/roi_heads/roi_extractors/base_roi_extractor.py')
    single = _load('mmdet.models.roi_heads.roi_extractors.single_level_roi_extractor',
                   d + 'models/roi_heads/roi_extractors/single_level_roi_extractor.py')
    rx.BaseRoIExtractor = base.BaseRoIExtractor
    rx.SingleRoIExtractor = single.SingleRoIExtractor
    bbox_nms = _load('mmdet.core.post_processing.bbox_nms', d + 'core/post_processing/bbox_nms.py')
    selsa = _load('mmtrack.models.aggregators.selsa_aggregator',
                  t + 'models/aggregators/selsa_aggregator.py')
    embed = _load('mmtrack.models.aggregators.embed_aggregator',
                  t + 'models/aggregators/embed_aggregator.py')
    troi = _load('mmtrack.models.roi_heads.roi_extractors.temporal_roi_align',
                 t + 'models/roi_heads/roi_extractors/temporal_roi_align.py')
    mm_single = _load('mmtrack.models.roi_heads.roi_extractors.single_level_roi_extractor',
                      t + 'models/roi_heads/roi_extractors/single_level_roi_extractor.py')
    flow = _load('mmtrack.core.motion.flow', t + 'core/motion/flow.py')
    flownet = _load('mmtrack.models.motion.flownet_simple', t + 'models/motion/flownet_simple.py')
    den_path = t + 'models/aggregators/denoising2_aggregator.py'
    denoise = _load('mmtrack.models.aggregators.denoising2_aggregator', den_path) \
        if os.path.exists(os.path.join(REF_ROOT, den_path)) else None

    ns = types.SimpleNamespace(
        SelsaAggregator=selsa.SelsaAggregator,
        EmbedAggregator=embed.EmbedAggregator,
        TemporalRoIAlign=troi.TemporalRoIAlign,
        SingleRoIExtractor=mm_single.SingleRoIExtractor,
        MMDetSingleRoIExtractor=single.SingleRoIExtractor,
        flow_warp_feats=flow.flow_warp_feats,
        FlowNetSimple=flownet.FlowNetSimple,
        Denoising2Aggergator=getattr(denoise, 'Denoising2Aggergator', None),
        TemporalAttentionFusion=getattr(denoise, 'TemporalAttentionFusion', None),
        ModulatedDCNPack=getattr(denoise, 'ModulatedDCNPack', None),
        multiclass_nms=bbox_nms.multiclass_nms,
        batched_nms=_batched_nms,
        RoIAlign=_RoIAlign,
        ROI_EXTRACTORS=ROI_EXTRACTORS,
        AGGREGATORS=AGGREGATORS,
    )
    _LOADED = ns
    return ns


_RPN = None


def load_rpn():
    """The reference's own ``RPNHead._get_bboxes`` and ``DeltaXYWHBBoxCoder`` (files loaded unmodified), for the golden
    vectors of the RPN proposal stage.  Extra stand-ins: mmcv.ConfigDict (attribute dict),
    mmcv.jit / mmcv.cnn.normal_init (no-ops), AnchorHead / RPNTestMixin (empty bases: ``_get_bboxes`` only reads
    ``self.use_sigmoid_cls``, ``self.bbox_coder`` and ``self.test_cfg``)."""
    global _RPN
    if _RPN is not None:
        return _RPN
    load()
    mmcv = sys.modules['mmcv']

    class ConfigDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    mmcv.ConfigDict = ConfigDict
    mmcv.jit = lambda *a, **kw: (lambda fn: fn)
    sys.modules['mmcv.cnn'].normal_init = lambda *a, **kw: None
    HEADS = _Registry('head')
    BBOX_CODERS = _Registry('bbox_coder')
    sys.modules['mmdet.models.builder'].HEADS = HEADS
    _mod('mmdet.core.bbox.builder', BBOX_CODERS=BBOX_CODERS)
    _mod('mmdet.core.bbox.coder')
    _mod('mmdet.models.dense_heads')
    _mod('mmdet.models.dense_heads.anchor_head', AnchorHead=type('AnchorHead', (nn.Module,), {}))
    _mod('mmdet.models.dense_heads.rpn_test_mixin', RPNTestMixin=type('RPNTestMixin', (object,), {}))
    d = 'mmdetection/mmdet/'
    _load('mmdet.core.bbox.coder.base_bbox_coder', d + 'core/bbox/coder/base_bbox_coder.py')
    coder = _load('mmdet.core.bbox.coder.delta_xywh_bbox_coder', d + 'core/bbox/coder/delta_xywh_bbox_coder.py')
    rpn = _load('mmdet.models.dense_heads.rpn_head', d + 'models/dense_heads/rpn_head.py')
    _RPN = types.SimpleNamespace(RPNHead=rpn.RPNHead, DeltaXYWHBBoxCoder=coder.DeltaXYWHBBoxCoder,
                                 delta2bbox=coder.delta2bbox, ConfigDict=ConfigDict)
    return _RPN
