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
