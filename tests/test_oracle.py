"""CPU tests of the oracle itself: pinned against the golden vectors produced by the reference's own
unmodified files (tests/golden/make_golden.py), against torchvision (the executable stand-in for
mmcv-full's ops) and, when /root/reference is present (build container), against the reference live."""
import pytest
import torch

from oracle import ref_shim
from oracle import vod_oracle as O

from conftest import params, rel_err
from helpers import clustered_boxes, rpn_like_rois


def test_selsa_matches_golden(golden):
    for pre, heads in (('selsa', 16), ('selsa64', 2)):
        out = O.selsa_aggregate(golden[pre + '_x'], golden[pre + '_ref_x'], params(golden, pre + '_p.'), heads)
        assert rel_err(out, golden[pre + '_out']) < 1e-5


def test_flow_warp_matches_golden(golden):
    for p in ('warp', 'warp2'):
        out = O.flow_warp_feats(golden[p + '_x'], golden[p + '_flow'])
        # closed form vs ATen's normalise/un-normalise round trip: ~1e-5 abs on N(0,1) features
        assert (out - golden[p + '_out']).abs().max() < 1e-4


def test_embed_aggregator_matches_golden(golden):
    p = params(golden, 'embed_p.')
    convs = [(p['embed_convs.0.conv.weight'], p['embed_convs.0.conv.bias'], True),
             (p['embed_convs.1.conv.weight'], p['embed_convs.1.conv.bias'], False)]
    out = O.embed_aggregate(golden['embed_x'], golden['embed_ref_x'], convs)
    assert rel_err(out, golden['embed_out']) < 1e-5


def test_roi_align_and_troi_match_golden(golden):
    ra = O.roi_align(golden['troi_feat'], golden['troi_rois'], 7, 1 / 16, 2, True)
    assert rel_err(ra, golden['roialign_out']) < 1e-6
    msra = O.most_similar_roi_align(ra, golden['troi_ref'], 2)
    assert rel_err(msra, golden['msra_out']) < 1e-5
    p = params(golden, 'troi_p.')
    out = O.temporal_roi_align(golden['troi_feat'], golden['troi_rois'], golden['troi_ref'],
                               p['embed_network.conv.weight'], p['embed_network.conv.bias'], 2, 4)
    assert rel_err(out, golden['troi_out']) < 1e-5
    out0 = O.temporal_roi_align(golden['troi_feat'], golden['troi_rois'], golden['troi_ref'], None, None, 2, 0)
    assert rel_err(out0, golden['troi_mean_out']) < 1e-5
    assert rel_err(O.roi_align(golden['troi_ref'], golden['troi_ref_rois'], 7, 1 / 16, 2, True), golden['troi_ref_out']) < 1e-6


def test_nms_matches_golden(golden):
    b, s = golden['rpnnms_boxes'], golden['rpnnms_scores']
    d, k = O.batched_nms(b, s, torch.zeros(len(b), dtype=torch.long), dict(type='nms', iou_threshold=0.7))
    assert torch.equal(k, golden['rpnnms_keep']) and torch.equal(d, golden['rpnnms_dets'])
    d, k = O.batched_nms(b, s, golden['bnms_ids'], dict(type='nms', iou_threshold=0.5))
    assert torch.equal(k, golden['bnms_keep']) and torch.equal(d, golden['bnms_dets'])
    d, k = O.batched_nms(b, s, golden['bnms_ids'], dict(type='nms', iou_threshold=0.5, split_thr=1000))
    assert torch.equal(k, golden['bnms_split_keep']) and torch.equal(d, golden['bnms_split_dets'])
    d, l, k = O.multiclass_nms(golden['mcnms_bboxes'], golden['mcnms_scores'], 0.05, dict(type='nms', iou_threshold=0.5),
                               100, return_inds=True)
    assert torch.equal(k, golden['mcnms_keep']) and torch.equal(l, golden['mcnms_labels'])
    assert torch.equal(d, golden['mcnms_dets'])


def test_c_oracle_vs_torchvision():
    """mmcv-full's RoIAlign / nms are un-vendored; torchvision is the executable truth."""
    tv = pytest.importorskip('torchvision')
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 24, 19, 31, generator=g)
    rois = rpn_like_rois(g, 40, 2, 31 * 16., 19 * 16.)
    rois[0, 1:] = torch.tensor([-50., -30., 20., 10.])
    for aligned, sr, osz in ((True, 2, (7, 7)), (False, 2, (7, 7)), (True, 0, (5, 6))):
        ref = tv.ops.roi_align(feat, rois, osz, 1 / 16, sr, aligned)
        assert rel_err(O.roi_align(feat, rois, osz, 1 / 16, sr, aligned), ref) < 1e-6
    boxes = clustered_boxes(g, 4000, 60)
    scores = torch.rand(4000, generator=g)
    for thr in (0.3, 0.5, 0.7):
        assert torch.equal(O.nms(boxes, scores, thr)[1], tv.ops.nms(boxes, scores, thr))


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present (GPU box)')
def test_oracle_vs_live_reference():
    R = ref_shim.load()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        m = R.SelsaAggregator(256, 4)
        x, r = torch.randn(33, 256, generator=g), torch.randn(120, 256, generator=g)
        assert rel_err(O.selsa_aggregate(x, r, dict(m.state_dict()), 4), m(x, r)) < 1e-6
        x, f = torch.randn(3, 16, 38, 63, generator=g), torch.randn(3, 2, 608, 1008, generator=g) * 8
        assert (O.flow_warp_feats(x, f) - R.flow_warp_feats(x, f)).abs().max() < 1e-4
        m = R.TemporalRoIAlign(2, 4, roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2), out_channels=32,
                               featmap_strides=[16])
        feat, ref = torch.relu(torch.randn(1, 32, 20, 30, generator=g)), torch.relu(torch.randn(4, 32, 20, 30, generator=g))
        rois = rpn_like_rois(g, 12, 1, 480., 320.)
        a = m((feat,), rois, ref_feats=(ref,))
        b = O.temporal_roi_align(feat, rois, ref, m.embed_network.conv.weight, m.embed_network.conv.bias, 2, 4)
        assert rel_err(b, a) < 1e-5
        n = 5000
        boxes, scores = clustered_boxes(g, n, 80), torch.rand(n, generator=g)
        ids = torch.randint(0, 30, (n,), generator=g)
        for cfg in (dict(type='nms', iou_threshold=0.5), dict(type='nms', iou_threshold=0.5, split_thr=3000)):
            da, ka = R.batched_nms(boxes, scores, ids, cfg)
            db, kb = O.batched_nms(boxes, scores, ids, cfg)
            assert torch.equal(ka, kb) and torch.equal(da, db)


def test_rpn_get_bboxes_matches_reference_golden(rpn_golden):
    """oracle restatement of RPNHead._get_bboxes == the reference's own method (unmodified file) on the committed vectors."""
    G = rpn_golden
    img_shape = tuple(int(v) for v in G['img_shape'])
    for tag in ('a', 'b'):
        nms_pre, thr, mx = G['cfg_' + tag].tolist()
        for b in range(G['cls'].shape[0]):
            d = O.rpn_get_bboxes(G['cls'][b], G['reg'][b], G['anchors'], img_shape, int(nms_pre), thr, int(mx))
            assert torch.equal(d, G['dets_%s_%d' % (tag, b)])


def test_flownet_simple_mirror_matches_the_reference():
    """our FlowNetSimple (motion.py) has the reference's parameter names and, with the reference's weights,
    gives the reference's flow bit for bit on the CPU (reference class loaded unmodified through oracle/ref_shim.py)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference files not present (neither /root/reference nor oracle/_ref)')
    import lowlightenvironmentvideoobjectdetection_b200 as vod
    ns = ref_shim.load()
    torch.manual_seed(0)
    ref = ns.FlowNetSimple(img_scale_factor=0.5)
    ours = vod.build_motion(dict(type='FlowNetSimple', img_scale_factor=0.5))
    assert list(ours.state_dict()) == list(ref.state_dict())
    ours.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(3)
    imgs = torch.randn(2, 6, 96, 128, generator=g)
    metas = [dict(img_shape=(90, 120, 3), img_norm_cfg=dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375]))]
    with torch.no_grad():
        a, b = ref(imgs.clone(), metas), ours(imgs.clone(), metas)
        lr, info = ours(imgs.clone(), metas, return_lowres=True)
    assert a.shape == (2, 2, 96, 128) and torch.equal(a, b)
    assert lr.shape == (2, 2, 12, 16) and info == dict(up_scale=8.0, mult1=8.0, mult2=5.0, full_size=(96, 128))


def test_modulated_deform_conv_restatement_vs_torchvision():
    """mmcv-full's modulated_deform_conv2d is not vendored: the oracle restates the published DCNv2 algorithm; pinned here
    against torchvision.ops.deform_conv2d (same offset layout and border rule) incl. stride 2 and samples outside the map."""
    from torchvision.ops import deform_conv2d
    g = torch.Generator().manual_seed(0)
    for (B, C, H, W, G, Co, stride) in [(2, 16, 9, 11, 4, 8, 1), (1, 8, 10, 7, 2, 5, 2)]:
        Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
        x = torch.randn(B, C, H, W, generator=g)
        off = torch.randn(B, 2 * G * 9, Ho, Wo, generator=g) * 3
        m = torch.rand(B, G * 9, Ho, Wo, generator=g)
        w, b = torch.randn(Co, C, 3, 3, generator=g), torch.randn(Co, generator=g)
        want = deform_conv2d(x, off, w, b, stride=stride, padding=1, mask=m)
        got = O.modulated_deform_conv2d(x, off, m, w, b, stride, 1, 1, 1, G)
        assert (got - want).abs().max() <= 1e-5 * want.abs().max()


def test_temporal_attention_fusion_restatement_vs_reference_file():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference files not staged')
    R = ref_shim.load()
    torch.manual_seed(0)
    taf = R.TemporalAttentionFusion(16, 8, emb_nums=3).eval()
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.weight, 0, 0.05)
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.bias, 0, 0.5)
    x = torch.randn(4, 16, 10, 13)
    with torch.no_grad():
        want = taf(x.clone())
    got = O.temporal_attention_fusion(x, dict(taf.state_dict()))
    assert (got - want).abs().max() <= 1e-5 * want.abs().max()


def test_denoising_aggregator_mirror_has_the_reference_state_dict():
    from oracle import ref_shim
    import lowlightenvironmentvideoobjectdetection_b200 as vod
    if not ref_shim.available():
        pytest.skip('reference files not staged')
    R = ref_shim.load()
    ref, ours = R.Denoising2Aggergator(), vod.build_aggregator(dict(type='Denoising2Aggergator'))
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a) == list(b) and all(a[k].shape == b[k].shape for k in a)


def _denoise_golden():
    import numpy as np
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'denoise_golden.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_temporal_attention_fusion_restatement_vs_golden():
    """Committed fixture from the reference's own denoising2_aggregator.py (tests/golden/make_denoise_golden.py): pins the
    oracle's restatement (incl. its modulated deformable conv) without needing the reference tree."""
    gd = _denoise_golden()
    p = {k[len('taf_p.'):]: v for k, v in gd.items() if k.startswith('taf_p.')}
    got = O.temporal_attention_fusion(gd['taf_x'], p)
    assert got.shape == gd['taf_out'].shape
    assert (got - gd['taf_out']).abs().max() <= 1e-5 * gd['taf_out'].abs().max()
