"""mmtrack.core.motion mirror: ``flow_warp_feats`` (mmtracking/mmtrack/core/motion/flow.py:4-41)."""
from . import ops


def flow_warp_feats(x, flow):
    """Use flow to warp feature map.

    Args:
        x (Tensor): of shape (N, C, H_x, W_x).
        flow (Tensor): of shape (N, 2, H_f, W_f); channel 0 = x displacement, 1 = y, in flow-image px.

    Returns:
        Tensor: The warped feature map with shape (N, C, H_x, W_x) (a fresh, writable, contiguous
        tensor: FGFA overwrites one slot in place, mmtracking/mmtrack/models/vid/fgfa.py:281).

    Same assertions as the reference (checked by its tests/test_core/test_motion_utils.py:13-29).
    One CUDA kernel: flow resize (bilinear, align_corners=False, x scale), grid build and the
    border-clamped bilinear gather are fused; no grid tensor, no host meshgrid.
    """
    assert len(x.shape) == 4
    assert len(flow.shape) == 4 and flow.shape[1] == 2
    scale_factor = float(x.shape[-1]) / flow.shape[-1]
    # F.interpolate(scale_factor=s) yields floor(in * s); the reference's grid add then requires it to equal x's size
    assert int(flow.shape[-2] * scale_factor) == x.shape[-2] and int(flow.shape[-1] * scale_factor) == x.shape[-1], \
        'resized flow %s does not match the feature map %s' % (tuple(flow.shape[-2:]), tuple(x.shape[-2:]))
    assert len(x) == len(flow), 'x and flow must hold the same number of frames'
    return ops.flow_warp(x, flow).to(x.dtype)


def flow_warp_feats_shared(x, flows):
    """DFF batching: ``flow_warp_feats(x.expand(len(flows), ...), flows)`` in ONE launch without materialising
    the expanded map -- all non-key frames of a key interval warp the same key-frame features
    (mmtracking/mmtrack/models/vid/dff.py:210-216 handles them one frame at a time).  x [1,C,H,W], flows [F,2,Hf,Wf]."""
    assert len(x.shape) == 4 and len(x) == 1
    assert len(flows.shape) == 4 and flows.shape[1] == 2
    scale_factor = float(x.shape[-1]) / flows.shape[-1]
    assert int(flows.shape[-2] * scale_factor) == x.shape[-2] and int(flows.shape[-1] * scale_factor) == x.shape[-1]
    return ops.flow_warp(x, flows).to(x.dtype)


class DFFFeatureMemo:
    """The feature side of ``DFF.extract_feats`` (mmtracking/mmtrack/models/vid/dff.py:184-217): the key frame's feature maps are
    kept, every other frame's features are the key maps warped by that frame's flow.

      memo.set_key(feats)            key frame (frame_id % key_frame_interval == 0): remember its maps (:205-209)
      memo.extract_feats(flow)       a non-key frame: [flow_warp_feats(f, flow) for f in key maps]          (:213-216)
      memo.extract_feats_interval(flows)   all F non-key frames of the interval at once -> per level [F,C,H,W], one launch
    """

    def __init__(self, key_frame_interval=10):
        self.key_frame_interval = key_frame_interval
        self.feats = None

    def is_key_frame(self, frame_id):
        return frame_id % self.key_frame_interval == 0

    def set_key(self, feats):
        self.feats = list(feats) if isinstance(feats, (list, tuple)) else [feats]
        return self.feats

    def extract_feats(self, flow):
        assert self.feats is not None, 'no key frame yet'
        return [flow_warp_feats(f, flow) for f in self.feats]

    def extract_feats_interval(self, flows):
        assert self.feats is not None, 'no key frame yet'
        return [flow_warp_feats_shared(f, flows) for f in self.feats]

    def extract_feats_lowres(self, flow_lr, lowres_info):
        """Same from FlowNetSimple's low-resolution prediction (``forward(..., return_lowres=True)``), for one frame or for all
        non-key frames of the interval: the x8 upsample of the flow is evaluated inside the warp kernel."""
        assert self.feats is not None, 'no key frame yet'
        return [flow_warp_feats_lowres(f, flow_lr, **lowres_info) for f in self.feats]


# ------------------------------------------------------------------------------------------------ FlowNetSimple
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from .registry import MOTION, ConvModule  # noqa: E402

# (name, [(in, out, kernel, stride), ...]) -- the contracting part of FlowNetS (arXiv:1504.06852, table in section 3)
_FLOWNET_ENCODER = (
    ('conv1', ((6, 64, 7, 2),)),
    ('conv2', ((64, 128, 5, 2),)),
    ('conv3', ((128, 256, 5, 2), (256, 256, 3, 1))),
    ('conv4', ((256, 512, 3, 2), (512, 512, 3, 1))),
    ('conv5', ((512, 512, 3, 2), (512, 512, 3, 1))),
    ('conv6', ((512, 1024, 3, 2), (1024, 1024, 3, 1))),
)
# refinement level k consumes cat(conv_k output, deconv_{k} output, upsampled flow): (level, input channels, deconv output channels)
_FLOWNET_DECODER = ((5, 1024, 512), (4, 1026, 256), (3, 770, 128), (2, 386, 64))
_LEAKY = dict(type='LeakyReLU', negative_slope=0.1)


@MOTION.register_module()
class FlowNetSimple(nn.Module):
    """B200-side mirror of mmtrack's FlowNetSimple (mmtracking/mmtrack/models/motion/flownet_simple.py:8-256): same constructor,
    same parameter names (``conv{1-6}.{j}.conv``, ``deconv{2-5}.conv``, ``predict_flow{3-6}.conv``, ``upsample_flow{2-5}.conv``,
    ``predict_flow.conv``), same arithmetic -- the convolutions are library calls here as there (cuDNN; ``.to(torch.bfloat16)`` /
    channels_last are the caller's choice).  What is new is the hand-off to the warp: ``forward(..., return_lowres=True)`` stops
    before the final x(4/img_scale_factor) bilinear upsample and returns the network's low-resolution prediction with the
    upsample parameters, which ``flow_warp_feats_lowres`` evaluates on the fly (the full-resolution flow -- 4.9 MB per frame
    pair, 152 MB for FGFA's 31 pairs -- is then never written)."""

    def __init__(self, img_scale_factor, out_indices=[2, 3, 4, 5, 6], flow_scale_factor=5.0,
                 flow_img_norm_std=[255.0, 255.0, 255.0], flow_img_norm_mean=[0.411, 0.432, 0.450]):
        super().__init__()
        self.img_scale_factor = img_scale_factor
        self.out_indices = out_indices
        self.flow_scale_factor = flow_scale_factor
        self.flow_img_norm_mean = flow_img_norm_mean
        self.flow_img_norm_std = flow_img_norm_std
        for name, convs in _FLOWNET_ENCODER:
            self.add_module(name, nn.ModuleList([
                ConvModule(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=True, conv_cfg=dict(type='Conv'), act_cfg=_LEAKY)
                for cin, cout, k, stride in convs]))
        for level, cin, cout in _FLOWNET_DECODER:
            self.add_module('deconv%d' % level, ConvModule(cin, cout, 4, stride=2, padding=1, bias=False,
                                                           conv_cfg=dict(type='deconv'), act_cfg=_LEAKY))
            self.add_module('predict_flow%d' % (level + 1), ConvModule(cin, 2, 3, stride=1, padding=1, bias=False,
                                                                       conv_cfg=dict(type='Conv'), act_cfg=None))
            self.add_module('upsample_flow%d' % level, ConvModule(2, 2, 4, stride=2, padding=1, bias=False,
                                                                  conv_cfg=dict(type='deconv'), act_cfg=None))
        self.predict_flow = ConvModule(194, 2, 3, stride=1, padding=1, bias=False, conv_cfg=dict(type='Conv'), act_cfg=None)

    def init_weights(self):
        pass

    def prepare_imgs(self, imgs, img_metas):
        """flownet_simple.py:148-191: undo the detector's normalisation, apply FlowNet's, zero the padding, rescale."""
        cfg = img_metas[0]['img_norm_cfg']
        # the four [1,6,1,1] constants are built once per (device, dtype, normalisation) -- the reference caches them on first
        # use too (:166-180) -- so that later calls launch no host-to-device copy and can be captured into a CUDA graph
        key = (imgs.device, imgs.dtype, tuple(cfg['mean']), tuple(cfg['std']))
        consts = getattr(self, '_norm_consts', None)
        if consts is None or consts[0] != key:
            def six(v):
                return torch.tensor(v, device=imgs.device, dtype=imgs.dtype).repeat(2)[None, :, None, None]
            consts = (key, six(cfg['std']), six(cfg['mean']), six(self.flow_img_norm_std), six(self.flow_img_norm_mean))
            self._norm_consts = consts
        _, std, mean, flow_std, flow_mean = consts
        flow_img = imgs * std + mean
        flow_img = flow_img / flow_std - flow_mean
        h, w = img_metas[0]['img_shape'][:2]
        flow_img[:, :, h:, :] = 0.0
        flow_img[:, :, :, w:] = 0.0
        return F.interpolate(flow_img, scale_factor=self.img_scale_factor, mode='bilinear', align_corners=False)

    @staticmethod
    def _crop_like(t, target):
        return t if t.shape[2:] == target.shape[2:] else t[:, :, :target.size(2), :target.size(3)]

    def forward(self, imgs, img_metas, return_lowres=False):
        """imgs [N,6,H,W] image pairs -> flow [N,2,H,W]; with ``return_lowres`` -> (flow_lr [N,2,h,w], dict(up_scale, mult1,
        mult2, full_size)) for ``flow_warp_feats_lowres``."""
        x = self.prepare_imgs(imgs, img_metas)
        feats = {}
        for i, (name, _) in enumerate(_FLOWNET_ENCODER, 1):
            for m in getattr(self, name):
                x = m(x)
            if i in self.out_indices:
                feats[i] = x
        cat = feats[6]
        for level, _, _ in _FLOWNET_DECODER:                       # flownet_simple.py:212-227
            skip = feats[level]
            flow = getattr(self, 'predict_flow%d' % (level + 1))(cat)
            up = self._crop_like(getattr(self, 'upsample_flow%d' % level)(flow), skip)
            dec = self._crop_like(getattr(self, 'deconv%d' % level)(cat), skip)
            cat = torch.cat((skip, dec, up), dim=1)
        flow = self.predict_flow(cat)
        up_scale = 4 / self.img_scale_factor
        if return_lowres:
            full = (int(flow.shape[2] * up_scale), int(flow.shape[3] * up_scale))
            return flow, dict(up_scale=up_scale, mult1=up_scale, mult2=self.flow_scale_factor, full_size=full)
        flow = F.interpolate(flow, scale_factor=up_scale, mode='bilinear', align_corners=False)   # :229-233
        flow *= up_scale
        flow *= self.flow_scale_factor
        return flow


def flow_warp_feats_lowres(x, flow_lr, up_scale, mult1, mult2, full_size):
    """``flow_warp_feats(x, mult2 * (mult1 * interpolate(flow_lr, scale_factor=up_scale)))`` without materialising the
    full-resolution flow (x may hold one map shared by all flows, as in DFF).  Arguments after ``flow_lr`` are the dict
    ``FlowNetSimple.forward(..., return_lowres=True)`` returns."""
    assert len(x.shape) == 4 and len(flow_lr.shape) == 4 and flow_lr.shape[1] == 2
    assert len(x) in (1, len(flow_lr))
    scale_factor = float(x.shape[-1]) / full_size[1]
    assert int(full_size[0] * scale_factor) == x.shape[-2] and int(full_size[1] * scale_factor) == x.shape[-1], \
        'resized flow %s does not match the feature map %s' % (tuple(full_size), tuple(x.shape[-2:]))
    return ops.flow_warp_lowres(x, flow_lr, full_size, up_scale, mult1, mult2).to(x.dtype)
