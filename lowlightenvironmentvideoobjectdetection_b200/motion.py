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
