"""Reference-frame cache: a key-frame step through the cache gives the results of the uncached step.

The reference recomputes, on every key frame, the RoI features, shared FCs and K/V projections of ALL reference frames
(mmtracking/mmtrack/models/roi_heads/selsa_roi_head.py:83-93, roi_heads/bbox_heads/selsa_bbox_head.py:53-58) although its frame
memory only turns over slowly (mmtrack/models/vid/selsa.py:207-249).  Everything our own kernels produce must be IDENTICAL with
and without the cache (same kernels, same per-RoI arithmetic, same reduction order because the cache keeps the reference
order); the library GEMMs see other batch shapes (300 new rows instead of 4500), so cuBLAS may pick another kernel and the
head's outputs agree to fp32 rounding, not bit for bit -- stated as a tolerance below.
"""
import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod

from helpers import rpn_like_rois

pytestmark = pytest.mark.gpu
DEV = 'cuda'
C, H, W, D, CLASSES = 64, 12, 20, 128, 6
IMG = (H * 16, W * 16, 3)


def _head(troi=True, fcs=2):
    torch.manual_seed(0)
    ext = dict(type='TemporalRoIAlign' if troi else 'SingleRoIExtractor',
               roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2), out_channels=C, featmap_strides=[16])
    if troi:
        ext.update(num_most_similar_points=2, num_temporal_attention_blocks=4)
    head = vod.SelsaRoIHead(bbox_roi_extractor=ext,
                            bbox_head=dict(type='SelsaBBoxHead', num_shared_fcs=fcs, in_channels=C, fc_out_channels=D,
                                           num_classes=CLASSES, aggregator=dict(type='SelsaAggregator', in_channels=D,
                                                                                num_attention_blocks=2))).to(DEV).eval()
    torch.nn.init.normal_(head.bbox_head.fc_cls.weight, 0, 0.3)
    torch.nn.init.normal_(head.bbox_head.fc_reg.weight, 0, 0.05)
    return head


def _clip(n_frames, N, seed):
    g = torch.Generator().manual_seed(seed)
    maps = torch.relu(torch.randn(n_frames, C, H, W, generator=g)).to(DEV)
    props = [rpn_like_rois(g, N, 1, W * 16., H * 16.)[:, 1:].to(DEV) for _ in range(n_frames)]
    metas = [dict(video_id=seed, frame_id=f, img_shape=IMG, scale_factor=(1., 1., 1., 1.)) for f in range(n_frames)]
    return maps, props, metas


def _compare(a, b, tol=2e-4):
    (da, la), (db, lb) = a, b
    assert da[0].shape == db[0].shape and torch.equal(la[0], lb[0])
    assert (da[0] - db[0]).abs().max() < tol, float((da[0] - db[0]).abs().max())      # px / score


@pytest.mark.parametrize('troi,T,mode', [(True, 5, 'fifo'), (True, 9, 'fifo'), (True, 5, 'adaptive'), (False, 5, 'fifo'),
                                         (True, 9, 'adaptive')])
def test_cached_simple_test_matches_uncached_over_a_clip(troi, T, mode):
    """40 key frames.  'fifo': the reference set is the window [k-l, k+r] sliding by one frame per key frame (fixed stride 1:
    one frame enters, one leaves, the key frame's slot sits in the middle); 'adaptive': T-1 fixed memory frames + the key frame
    last (selsa.py:207-225).  T = 9 stacks 10 frames, i.e. TemporalRoIAlign's key-projected logits path."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    N, steps = 20, 40
    head = _head(troi)
    maps, props, metas = _clip(steps + T, N, 7 + T)
    computed = []
    inner = head.update_ref_cache
    head.update_ref_cache = lambda cache, slots, *a, **k: (computed.append(len(slots)), inner(cache, slots, *a, **k))[1]
    left = T // 2
    for step in range(steps):
        if mode == 'fifo':
            k = left + step
            ids = list(range(k - left, k - left + T))
        else:
            k = T - 1 + step
            ids = list(range(T - 1)) + [k]
        x = maps[k:k + 1]
        ref_x = maps[ids]
        args = ((x,), (ref_x,), [props[k]], [props[i] for i in ids], [metas[k]])
        head.use_ref_cache = True
        got = head.simple_test(*args, ref_img_metas=[metas[i] for i in ids])
        head.use_ref_cache = False
        want = head.simple_test(*args, ref_img_metas=[metas[i] for i in ids])
        _compare(got, want)
    # the cache did its job: after the first step only the frame that entered the window is computed (none in adaptive mode)
    assert computed[0] == T - 1
    assert computed[1:] == ([1] * (steps - 1) if mode == 'fifo' else [])


@pytest.mark.parametrize('mode,T', [('fifo', 5), ('adaptive', 9)])
def test_cached_simple_test_with_cuda_graphs_matches_eager(mode, T):
    """``head.use_cuda_graphs = True``: the cached drop-in call replays one captured key-frame step per call (the frames that
    enter the reference set are still processed eagerly).  Over a clip with memory turnover the detections equal the eager cached
    call's -- same kernels on the same inputs, so bit for bit -- one graph serves the whole clip, and the caller's outputs are
    not aliased by the next call."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    N, steps = 20, 12
    eager, graphed = _head(True), _head(True)
    graphed.load_state_dict(eager.state_dict())
    graphed.use_cuda_graphs = True
    maps, props, metas = _clip(steps + T, N, 31 + T)
    left = T // 2
    kept = []
    for step in range(steps):
        if mode == 'fifo':
            k = left + step
            ids = list(range(k - left, k - left + T))
        else:
            k = T - 1 + step
            ids = list(range(T - 1)) + [k]
        args = ((maps[k:k + 1],), (maps[ids],), [props[k]], [props[i] for i in ids], [metas[k]])
        want = eager.simple_test(*args, ref_img_metas=[metas[i] for i in ids])
        got = graphed.simple_test(*args, ref_img_metas=[metas[i] for i in ids])
        assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[1][0], want[1][0])
        kept.append((got[0][0], want[0][0].clone()))
    assert len(graphed._step_graphs) == 1
    assert all(torch.equal(a, b) for a, b in kept)          # earlier results were not overwritten by later replays


def test_uncached_simple_test_with_cuda_graphs_matches_eager():
    """``head.use_cuda_graphs = True`` without ``ref_img_metas``: the plain (uncached) drop-in call replays the captured
    ``simple_test_device`` step; same detections as the eager call for every frame of a clip, one graph for the clip, a second one
    when the proposal count changes."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, N, steps = 5, 20, 6
    eager, graphed = _head(True), _head(True)
    graphed.load_state_dict(eager.state_dict())
    graphed.use_cuda_graphs = True
    maps, props, metas = _clip(steps + T, N, 57)
    for step in range(steps):
        ids = list(range(step, step + T))
        k = ids[-1]
        args = ((maps[k:k + 1],), (maps[ids],), [props[k]], [props[i] for i in ids], [metas[k]])
        want, got = eager.simple_test(*args), graphed.simple_test(*args)
        assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[1][0], want[1][0])
    assert len(graphed._step_graphs) == 1
    args = ((maps[k:k + 1],), (maps[ids],), [props[k][:11]], [props[i] for i in ids], [metas[k]])
    want, got = eager.simple_test(*args), graphed.simple_test(*args)
    assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[1][0], want[1][0])
    assert len(graphed._step_graphs) == 2


def test_cached_device_step_our_kernels_bit_identical():
    """Device-level API: fill the cache once, then every key frame's RoI features (RoIAlign, most-similar sampling, TAFA: our
    kernels + the same-shape key-slot conv) are bit-identical to the uncached extractor; the head's scores agree to fp32 GEMM
    rounding.  Also: a weight change invalidates the cache."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, N = 9, 24
    head = _head(True, fcs=3)
    maps, props, metas = _clip(T + 6, N, 3)
    cache = head.new_ref_cache(T, N, (C, H, W))
    memo = list(range(T - 1))
    head.update_ref_cache(cache, memo, maps[memo], vod.bbox2roi([props[i] for i in memo]), keys=memo)
    for k in range(T - 1, T + 5):
        rois = vod.bbox2roi([props[k]])
        dets, labels, count, mid = head.simple_test_cached_device((maps[k:k + 1],), rois, rois, cache, T - 1, IMG, (1., 1., 1., 1.),
                                                                  return_feats=True)
        ids = memo + [k]
        ref_rois = vod.bbox2roi([props[i] for i in ids])
        res = head._bbox_forward((maps[k:k + 1],), (maps[ids],), rois, ref_rois)
        want_feats = res['bbox_feats'].permute(0, 2, 3, 1).reshape(N, -1)            # channels_last rows, as the cache path holds them
        assert torch.equal(mid['bbox_feats'], want_feats)
        assert (mid['cls_score'] - res['cls_score']).abs().max() < 1e-4 * res['cls_score'].abs().max()
        assert (mid['bbox_pred'] - res['bbox_pred']).abs().max() < 1e-4 * res['bbox_pred'].abs().max()
        d2, l2, c2 = head.bbox_head.get_bboxes_device(rois, res['cls_score'], res['bbox_pred'], IMG, (1., 1., 1., 1.), cfg=head.test_cfg)
        assert int(count) == int(c2) and torch.equal(labels, l2) and (dets - d2).abs().max() < 2e-4
    assert cache.compatible(head, T, N, (C, H, W), maps.device)
    with torch.no_grad():
        head.bbox_head.shared_fcs[1].weight.add_(0.01)
    assert not cache.compatible(head, T, N, (C, H, W), maps.device)


def test_cached_step_and_cache_fill_inside_cuda_graphs():
    """What bench.py's `cached` loop replays: graph A = fill the T-1 memory slots from static map / RoI buffers (clip start),
    graph B = one key-frame step against the cache.  Replayed over two clips, compared with the eager uncached step."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, N = 9, 24
    head = _head(True, fcs=2)
    cache = head.new_ref_cache(T, N, (C, H, W))
    st_memo = torch.zeros(T - 1, C, H, W, device=DEV)
    st_memo_rois = torch.zeros((T - 1) * N, 5, device=DEV)
    st_memo_rois[:, 0] = torch.arange(T - 1, device=DEV, dtype=torch.float32).repeat_interleave(N)
    st_key = torch.zeros(1, C, H, W, device=DEV)
    st_rois = torch.zeros(N, 5, device=DEV)
    maps0, props0, _ = _clip(T + 3, N, 11)
    st_memo.copy_(maps0[:T - 1]); st_memo_rois[:, 1:] = torch.cat(props0[:T - 1], 0)
    st_key.copy_(maps0[T - 1:T]); st_rois[:, 1:] = props0[T - 1]
    slots = list(range(T - 1))
    g_fill, _ = head.capture_callable(lambda: head.update_ref_cache(cache, slots, st_memo, st_memo_rois))
    g_step, (dets, labels, count) = head.capture_callable(
        lambda: head.simple_test_cached_device((st_key,), st_rois, st_rois, cache, T - 1, IMG, (1., 1., 1., 1.)))
    for seed in (12, 13):
        maps, props, _ = _clip(T + 3, N, seed)
        st_memo.copy_(maps[:T - 1]); st_memo_rois[:, 1:] = torch.cat(props[:T - 1], 0)
        g_fill.replay()
        for k in range(T - 1, T + 2):
            st_key.copy_(maps[k:k + 1]); st_rois[:, 1:] = props[k]
            g_step.replay()
            torch.cuda.synchronize()
            ids = slots + [k]
            ref_rois = vod.bbox2roi([props[i] for i in ids])
            de, le, ce = head.simple_test_device((maps[k:k + 1],), (maps[ids],), st_rois, ref_rois, IMG, (1., 1., 1., 1.))
            n = int(ce)
            assert int(count) == n and torch.equal(labels[:n], le[:n]) and (dets[:n] - de[:n]).abs().max() < 2e-4
