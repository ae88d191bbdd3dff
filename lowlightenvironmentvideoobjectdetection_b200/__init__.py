"""vodagg-b200: B200-native (sm_100a) implementation of the multi-frame feature-aggregation hot path of the
MMTracking / MMDetection video-object-detection stack, behind the reference's own registries and call
signatures.  Host code is Python/PyTorch (device memory + streams only); every op launches hand-written CUDA
kernels from libvodagg.so through a C ABI (include/vodagg.h).  No Triton, no backend dispatch, no CPU fallback.
"""
from . import _lib, ops  # noqa: F401
from ._lib import VodError  # noqa: F401
from .aggregators import EmbedAggregator, SelsaAggregator  # noqa: F401
from .denoise import RDB, Denoising2Aggergator, ModulatedDCNPack, TemporalAttentionFusion  # noqa: F401
from .heads import RefFrameCache, SelsaBBoxHead, SelsaRoIHead, Shared2FCBBoxHead, StandardRoIHead  # noqa: F401
from .motion import DFFFeatureMemo, FlowNetSimple, flow_warp_feats, flow_warp_feats_lowres, flow_warp_feats_shared  # noqa: F401
from .ops import RoIAlign, batched_nms, nms, roi_align  # noqa: F401
from .post_processing import (bbox2roi, delta2bbox, multiclass_nms, rpn_batched_nms, rpn_get_bboxes,  # noqa: F401
                              rpn_get_bboxes_device)
from .registry import (AGGREGATORS, HEADS, MOTION, ROI_EXTRACTORS, ConvModule, Registry, build_aggregator,  # noqa: F401
                       build_from_cfg, build_head, build_motion, build_roi_extractor, force_fp32, register_into_openmmlab)
from .roi_extractors import BaseRoIExtractor, SingleRoIExtractor, TemporalRoIAlign  # noqa: F401

__version__ = '0.1.0'
