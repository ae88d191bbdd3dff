"""Round-2 widening: the FGFA / DFF callers and the DFF batched warp, plus the execution variants added
this round (two-stream overlap, few-RoI channel slabs, split-K first FC) against the oracle or the single-stream path."""
import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops
from oracle import vod_oracle as O

from conftest import rel_err
from helpers import rpn_like_rois

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_flow_warp_shared_map_is_the_expanded_warp():
    """DFF batching (dff.py:210-216 handles one non-key frame at a time): all flows of a key interval warp the SAME key map in
    one launch; bit-identical to warping an expanded copy, and to the oracle."""
    g = torch.Generator().manual_seed(3)
    C, H, W, F = 96, 13, 21, 9
    key = torch.relu(torch.randn(1, C, H, W, generator=g)) * 0.25            # low-light scaling
    flows = torch.randn(F, 2, H * 16, W * 16, generator=g) * 6
    shared = vod.flow_warp_feats_shared(key.to(DEV), flows.to(DEV))
    expanded = vod.flow_warp_feats(key.to(DEV).expand(F, C, H, W), flows.to(DEV))
    assert shared.shape == (F, C, H, W) and torch.equal(shared, expanded)
    assert rel_err(shared, O.flow_warp_feats(key.expand(F, C, H, W).contiguous(), flows)) < 2e-5
    memo = vod.DFFFeatureMemo(key_frame_interval=10)
    assert memo.is_key_frame(0) and memo.is_key_frame(20) and not memo.is_key_frame(7)
    memo.set_key((key.to(DEV),))
    per_frame = torch.cat([memo.extract_feats(flows[f:f + 1].to(DEV))[0] for f in range(F)], 0)
    assert torch.equal(memo.extract_feats_interval(flows.to(DEV))[0], per_frame)
    with pytest.raises(AssertionError):
        vod.flow_warp_feats_shared(key.to(DEV).expand(2, C, H, W), flows.to(DEV))


def test_standard_roi_head_vs_oracle():
    """StandardRoIHead + Shared2FCBBoxHead (the FGFA / DFF detectors' RoI head): RoIAlign -> 2 FC + ReLU ->
    fc_cls / fc_reg -> get_bboxes + multiclass NMS, against the oracle's pieces."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(4)
    C, H, W, D, classes, N = 64, 12, 20, 128, 6, 40
    head = vod.StandardRoIHead(
        bbox_roi_extractor=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                out_channels=C, featmap_strides=[16]),
        bbox_head=dict(type='Shared2FCBBoxHead', in_channels=C, fc_out_channels=D, num_classes=classes)).to(DEV).eval()
    torch.nn.init.normal_(head.bbox_head.fc_cls.weight, 0, 0.3)
    torch.nn.init.normal_(head.bbox_head.fc_reg.weight, 0, 0.05)
    assert list(head.bbox_head.state_dict()) == ['shared_fcs.0.weight', 'shared_fcs.0.bias', 'shared_fcs.1.weight', 'shared_fcs.1.bias',
                                                'fc_cls.weight', 'fc_cls.bias', 'fc_reg.weight', 'fc_reg.bias']
    x = torch.relu(torch.randn(1, C, H, W, generator=g))
    props = rpn_like_rois(g, N, 1, W * 16., H * 16.)[:, 1:]
    img = (H * 16, W * 16, 3)
    dets, labels = head.simple_test((x.to(DEV),), [props.to(DEV)], [dict(img_shape=img, scale_factor=(1., 1., 1., 1.))])
    sd = {k: v.detach().cpu() for k, v in head.bbox_head.state_dict().items()}
    rois = vod.bbox2roi([props])
    a = O.roi_align(x, rois, 7, 1 / 16, 2, True).flatten(1)
    for i in range(2):
        a = torch.relu(torch.nn.functional.linear(a, sd['shared_fcs.%d.weight' % i], sd['shared_fcs.%d.bias' % i]))
    cls = torch.nn.functional.linear(a, sd['fc_cls.weight'], sd['fc_cls.bias'])
    reg = torch.nn.functional.linear(a, sd['fc_reg.weight'], sd['fc_reg.bias'])
    d0, l0 = O.get_bboxes(rois, cls, reg, img, (1., 1., 1., 1.), False, 0.0001, dict(type='nms', iou_threshold=0.5), 100)
    assert dets[0].shape == d0.shape and torch.equal(labels[0].cpu(), l0)
    assert (dets[0].cpu() - d0).abs().max() < 1e-3


@pytest.mark.parametrize('K', [40, 700])
def test_roi_align_few_and_many_rois_vs_oracle(K):
    """K = 40: the launch is split into 128-channel slabs (few RoIs would leave most SMs idle); K = 700: one CTA per RoI."""
    g = torch.Generator().manual_seed(K)
    C, H, W = 512, 38, 63
    feat = torch.randn(2, C, H, W, generator=g)
    rois = rpn_like_rois(g, K // 2, 2)
    want = O.roi_align(feat, rois, 7, 1 / 16, 2, True)
    for cl in (False, True):
        got = ops.roi_align(feat.to(DEV), rois.to(DEV), 7, 1 / 16, 2, channels_last_out=cl)
        assert rel_err(got, want) < 2e-5


def test_two_stream_overlap_matches_single_stream():
    """The reference branch (RoIAlign of the reference RoIs + first FC) and the key-slot embed conv / G product run on a side
    stream next to the key branch; same kernels on the same data, so the step's outputs are those of the one-stream run."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(6)
    C, H, W, D, N, T = 64, 12, 20, 128, 24, 9
    head = vod.SelsaRoIHead(
        bbox_roi_extractor=dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2), out_channels=C, featmap_strides=[16]),
        bbox_head=dict(type='SelsaBBoxHead', num_shared_fcs=2, in_channels=C, fc_out_channels=D, num_classes=5,
                       aggregator=dict(type='SelsaAggregator', in_channels=D, num_attention_blocks=2))).to(DEV).eval()
    ref_x = torch.relu(torch.randn(T, C, H, W, generator=g)).to(DEV)
    rois = rpn_like_rois(g, N, 1, W * 16., H * 16.).to(DEV)
    ref_rois = rpn_like_rois(g, N, T, W * 16., H * 16.).to(DEV)
    outs = {}
    for overlap in (True, False):
        head.overlap = overlap
        head.bbox_roi_extractor.overlap = overlap
        for _ in range(3):      # repeated: a missing stream dependency would show up as run-to-run differences
            res = head._bbox_forward((ref_x[T - 1:],), (ref_x,), rois, ref_rois)
            torch.cuda.synchronize()
            outs.setdefault(overlap, []).append({k: v.clone() for k, v in res.items()})
    for k in ('bbox_feats', 'cls_score', 'bbox_pred'):
        for o in outs[True] + outs[False]:
            assert rel_err(o[k], outs[False][0][k]) < 1e-6, k


def test_split_k_first_fc_matches_linear():
    """SelsaBBoxHead._linear_few_rows (the cached step's first FC over 600 rows): batched split-K == F.linear to fp32 rounding."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(600, 25088, device=DEV, generator=g)
    w = torch.randn(1024, 25088, device=DEV, generator=g) * 0.01
    b = torch.randn(1024, device=DEV, generator=g)
    got = vod.SelsaBBoxHead._linear_few_rows(x, w, b)
    want = torch.nn.functional.linear(x.double(), w.double(), b.double()).float()
    assert rel_err(got, want) < 1e-5
    big = torch.randn(2000, 64, device=DEV, generator=g)     # outside the few-rows regime: plain F.linear
    assert torch.equal(vod.SelsaBBoxHead._linear_few_rows(big, w[:, :64].contiguous(), b), torch.nn.functional.linear(big, w[:, :64].contiguous(), b))


def test_flow_warp_from_lowres_flow_equals_materialised_flow():
    """FlowNetSimple ends with a x8 bilinear upsample + two scalar multiplies (flownet_simple.py:229-236) whose
    output flow_warp_feats immediately shrinks by 1/16.  vod_flow_warp_lowres evaluates the upsample at the 4 of every 256
    values the warp reads: same result as warping with the materialised full-resolution flow (oracle on the CPU)."""
    g = torch.Generator().manual_seed(12)
    C, H, W, F_ = 64, 12, 20, 5
    x = torch.relu(torch.randn(F_, C, H, W, generator=g))
    flow_lr = torch.randn(F_, 2, H * 2, W * 2, generator=g) * 0.3                 # FlowNet predicts at 1/8 of the image
    info = dict(up_scale=8.0, mult1=8.0, mult2=5.0, full_size=(H * 16, W * 16))
    full = torch.nn.functional.interpolate(flow_lr, scale_factor=8.0, mode='bilinear', align_corners=False)
    full = full * 8.0
    full = full * 5.0
    want = O.flow_warp_feats(x, full)
    got = vod.flow_warp_feats_lowres(x.to(DEV), flow_lr.to(DEV), **info)
    assert got.shape == want.shape and rel_err(got, want) < 2e-5
    # ... and bit-identical to our own warp of the flow materialised on the device with the same arithmetic order
    got_full = vod.flow_warp_feats(x.to(DEV), full.to(DEV))
    assert rel_err(got, got_full) < 2e-6
    # one key map shared by all flows (DFF interval), through the feature memo
    memo = vod.DFFFeatureMemo(10)
    memo.set_key((x[:1].to(DEV),))
    shared = memo.extract_feats_lowres(flow_lr.to(DEV), info)[0]
    assert rel_err(shared, O.flow_warp_feats(x[:1].expand(F_, C, H, W).contiguous(), full)) < 2e-5


def test_flownet_simple_mirror_on_device():
    """The FlowNetSimple mirror (library convs) on the GPU in fp32 and bf16 against itself on the CPU, and its low-resolution
    hand-off: warp(full-res flow) == warp_lowres(low-res prediction)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = vod.build_motion(dict(type='FlowNetSimple', img_scale_factor=0.5)).eval()
    g = torch.Generator().manual_seed(13)
    imgs = torch.randn(2, 6, 192, 320, generator=g)
    metas = [dict(img_shape=(180, 310, 3), img_norm_cfg=dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375]))]
    with torch.no_grad():
        want = net(imgs.clone(), metas)
        net = net.to(DEV)
        got = net(imgs.to(DEV), metas)
        assert got.shape == (2, 2, 192, 320) and rel_err(got, want) < 1e-3
        lr, info = net(imgs.to(DEV), metas, return_lowres=True)
        assert info['full_size'] == (192, 320) and lr.shape == (2, 2, 24, 40)
        x = torch.relu(torch.randn(2, 32, 12, 20, generator=g)).to(DEV)
        assert rel_err(vod.flow_warp_feats_lowres(x, lr, **info), vod.flow_warp_feats(x, got)) < 1e-5
        half = net.to(torch.bfloat16)(imgs.to(DEV).bfloat16(), metas)
        assert half.dtype == torch.bfloat16 and rel_err(half.float(), want) < 0.1        # bf16 convs: stated separately
