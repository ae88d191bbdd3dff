"""Golden vectors of the RPN proposal stage, produced by the reference's own
RPNHead._get_bboxes + DeltaXYWHBBoxCoder (mmdetection/mmdet/models/dense_heads/rpn_head.py:82-236, loaded unmodified
under oracle/ref_shim.load_rpn) -- build container only.  Run:  python tests/golden/make_rpn_golden.py"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def grid_anchors(H, W, stride=16, scales=(4, 8, 16), ratios=(0.5, 1.0, 2.0)):
    """Anchors in (location-major, anchor-minor) order, the order cls_score.permute(0, 2, 3, 1).reshape(-1) produces."""
    base = []
    for r in ratios:
        for s in scales:
            w, h = stride * s / r ** 0.5, stride * s * r ** 0.5
            base.append([-w / 2, -h / 2, w / 2, h / 2])
    base = torch.tensor(base)
    ys, xs = torch.meshgrid(torch.arange(H) * float(stride), torch.arange(W) * float(stride), indexing='ij')
    shift = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4)
    return (shift + base[None]).reshape(-1, 4)


def main():
    R = ref_shim.load_rpn()
    g = torch.Generator().manual_seed(20261018)
    B, H, W = 3, 12, 20
    anchors = grid_anchors(H, W)
    Ap = anchors.shape[0] // (H * W)
    cls = torch.randn(B, Ap, H, W, generator=g) * 2.5
    reg = torch.randn(B, Ap * 4, H, W, generator=g) * 0.4
    reg[0, :, 0, 0] = 9.0          # exercises the wh_ratio clip
    img_shape = (H * 16 - 5, W * 16 - 3, 3)
    head = object.__new__(R.RPNHead)
    torch.nn.Module.__init__(head)
    head.use_sigmoid_cls = True
    head.bbox_coder = R.DeltaXYWHBBoxCoder(target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.))
    arrays = dict(cls=cls.numpy(), reg=reg.numpy(), anchors=anchors.numpy(), img_shape=np.array(img_shape))
    for tag, (nms_pre, thr, mx) in dict(a=(600, 0.7, 100), b=(5000, 0.5, 40)).items():   # b: fewer anchors than nms_pre
        cfg = R.ConfigDict(nms_pre=nms_pre, nms=R.ConfigDict(type='nms', iou_threshold=thr), max_per_img=mx, min_bbox_size=0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = head._get_bboxes([cls], [reg], [anchors], [img_shape] * B, None, cfg)
        arrays['cfg_' + tag] = np.array([nms_pre, thr, mx], dtype=np.float64)
        for b in range(B):
            arrays['dets_%s_%d' % (tag, b)] = res[b].numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'rpn_golden.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
