"""Generates tests/golden/hotpath_golden.npz by running the reference's own
UNMODIFIED hot-path Python files (loaded from /root/reference under the shims
of oracle/ref_shim.py) on fixed-seed inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures hold inputs, weights and the reference's outputs, so the parity
tests on the GPU box (where /root/reference does not exist) need nothing else.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

warnings.filterwarnings('ignore')


def boxes_clustered(g, n, n_centres, img_w=1000., img_h=600., jitter=6.0):
    """RPN-like boxes clustered around a few centres so that NMS really suppresses."""
    c = torch.rand(n_centres, 2, generator=g) * torch.tensor([img_w, img_h])
    wh = torch.exp(torch.rand(n_centres, 2, generator=g) * 2.5 + 3.0)
    which = torch.randint(0, n_centres, (n,), generator=g)
    ctr = c[which] + torch.randn(n, 2, generator=g) * jitter
    sz = wh[which] * torch.exp(torch.randn(n, 2, generator=g) * 0.15)
    b = torch.cat([ctr - sz / 2, ctr + sz / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, img_w)
    b[:, 1::2] = b[:, 1::2].clamp(0, img_h)
    return b


def main():
    R = ref_shim.load()
    out = {}
    g = torch.Generator().manual_seed(20261018)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731

    with torch.no_grad():
        # --- SelsaAggregator (selsa_aggregator.py:29-73): D=128, 16 heads (d=8)
        torch.manual_seed(1)
        m = R.SelsaAggregator(128, 16)
        x, r = rn(20, 128), rn(70, 128)
        out['selsa_x'], out['selsa_ref_x'] = x, r
        for k, v in m.state_dict().items():
            out['selsa_p.' + k] = v
        out['selsa_out'] = m(x, r)
        # d=64 heads (the tcgen05 shape): D=128, 2 heads
        torch.manual_seed(2)
        m = R.SelsaAggregator(128, 2)
        x, r = rn(37, 128), rn(300, 128)
        out['selsa64_x'], out['selsa64_ref_x'] = x, r
        for k, v in m.state_dict().items():
            out['selsa64_p.' + k] = v
        out['selsa64_out'] = m(x, r)

        # --- flow_warp_feats (flow.py:4-41): stride-16 geometry and the reference test's 10->32
        x, f = rn(2, 8, 12, 20), rn(2, 2, 192, 320) * 8.0
        out['warp_x'], out['warp_flow'], out['warp_out'] = x, f, R.flow_warp_feats(x, f)
        x, f = rn(2, 8, 32, 32), rn(2, 2, 10, 10)
        out['warp2_x'], out['warp2_flow'], out['warp2_out'] = x, f, R.flow_warp_feats(x, f)

        # --- EmbedAggregator (embed_aggregator.py:50-81): 2 convs, C=16, T=3
        torch.manual_seed(3)
        m = R.EmbedAggregator(num_convs=2, channels=16, kernel_size=3)
        x, r = rn(1, 16, 10, 14), rn(3, 16, 10, 14)
        out['embed_x'], out['embed_ref_x'] = x, r
        for k, v in m.state_dict().items():
            out['embed_p.' + k] = v
        out['embed_out'] = m(x, r)

        # --- RoIAlign + TemporalRoIAlign (temporal_roi_align.py:183-207): C=64, 12x20 map, T=3
        torch.manual_seed(4)
        m = R.TemporalRoIAlign(num_most_similar_points=2, num_temporal_attention_blocks=4,
                               roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                               out_channels=64, featmap_strides=[16])
        feat, ref = torch.relu(rn(1, 64, 12, 20)), torch.relu(rn(3, 64, 12, 20))
        rois = torch.tensor([[0, 10., 20., 200., 150.], [0, 100., 50., 300., 180.],
                             [0, 0., 0., 319., 191.], [0, 150.5, 80.25, 170.75, 100.5],
                             [0, -20., -10., 40., 30.], [0, 250., 120., 400., 260.]])
        out['troi_feat'], out['troi_ref'], out['troi_rois'] = feat, ref, rois
        for k, v in m.state_dict().items():
            out['troi_p.' + k] = v
        roi_feats = R.RoIAlign(7, 1 / 16, 2)(feat, rois)
        out['roialign_out'] = roi_feats
        out['msra_out'] = m.most_similar_roi_align(roi_feats, ref).contiguous()
        out['troi_out'] = m((feat,), rois, ref_feats=(ref,))
        ref_rois = torch.cat([torch.randint(0, 3, (9, 1), generator=g).float(),
                              boxes_clustered(g, 9, 3, 320., 192.)], 1)
        out['troi_ref_rois'] = ref_rois
        out['troi_ref_out'] = m((ref,), ref_rois)
        m0 = R.TemporalRoIAlign(num_most_similar_points=2, num_temporal_attention_blocks=0,
                                roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                out_channels=64, featmap_strides=[16])
        out['troi_mean_out'] = m0((feat,), rois, ref_feats=(ref,))

        # --- multiclass_nms (bbox_nms.py:7-93): 120 proposals x 6 classes, clustered
        n, ncls = 120, 6
        base = boxes_clustered(g, n, 7)
        multi = (base[:, None, :] + rn(n, ncls, 4) * 3.0).reshape(n, ncls * 4)
        scores = torch.softmax(rn(n, ncls + 1) * 2.0, dim=1)
        cfg = dict(type='nms', iou_threshold=0.5)
        d, l, k = R.multiclass_nms(multi, scores, 0.05, cfg, 100, return_inds=True)
        out['mcnms_bboxes'], out['mcnms_scores'] = multi, scores
        out['mcnms_dets'], out['mcnms_labels'], out['mcnms_keep'] = d, l, k

        # --- batched_nms (mmcv 1.2.x semantics; rpn_head.py:233-235 style, single class, thr 0.7)
        n = 1500
        b = boxes_clustered(g, n, 40)
        s = torch.rand(n, generator=g)
        d, k = R.batched_nms(b, s, torch.zeros(n, dtype=torch.long), dict(type='nms', iou_threshold=0.7))
        out['rpnnms_boxes'], out['rpnnms_scores'], out['rpnnms_dets'], out['rpnnms_keep'] = b, s, d, k
        # multi-class, both below and above split_thr
        ids = torch.randint(0, 5, (n,), generator=g)
        d, k = R.batched_nms(b, s, ids, dict(type='nms', iou_threshold=0.5))
        out['bnms_ids'], out['bnms_dets'], out['bnms_keep'] = ids, d, k
        d, k = R.batched_nms(b, s, ids, dict(type='nms', iou_threshold=0.5, split_thr=1000))
        out['bnms_split_dets'], out['bnms_split_keep'] = d, k

    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
              for k, v in out.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hotpath_golden.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024), 'torch', torch.__version__)
    print('kept: mcnms', len(out['mcnms_keep']), 'rpn', len(out['rpnnms_keep']),
          'bnms', len(out['bnms_keep']), 'split', len(out['bnms_split_keep']))


if __name__ == '__main__':
    main()
