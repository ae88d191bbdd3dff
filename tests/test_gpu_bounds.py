"""Guard-band check of every device buffer the host wrappers hand to the kernels (outputs, workspaces, intermediates).

compute-sanitizer is not available on the GPU pool, so out-of-bounds WRITES are looked for the plain way: while an operator runs,
``torch.empty`` is replaced by an allocator that puts 4 KB of a known byte pattern on both sides of every CUDA buffer; afterwards
every band must still hold the pattern.  Shapes are deliberately ragged (odd RoI / proposal counts, channel counts that exercise
the scalar, the 128-bit and the tensor-core paths, partial tiles).  Each case also runs once on ordinary buffers first and the
two results must be identical, so a kernel that leaves part of its output unwritten (the guarded buffers start as the byte
pattern, ordinary ones as whatever the allocator held) shows up as well."""
import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops

from helpers import clustered_boxes, rpn_like_rois

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GUARD = 4096
PATTERN = 0xA5


class GuardedEmpty:
    """Context manager: torch.empty(...) on a CUDA device returns a view into a larger byte buffer with guard bands."""

    def __enter__(self):
        self.real = torch.empty
        self.bufs = []
        torch.empty = self.alloc
        return self

    def __exit__(self, *exc):
        torch.empty = self.real

    def alloc(self, *size, **kw):
        dev = kw.get('device')
        extra = set(kw) - {'dtype', 'device'}
        if dev is None or torch.device(dev).type != 'cuda' or extra or kw.get('out') is not None:
            return self.real(*size, **kw)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dtype = kw.get('dtype') or torch.get_default_dtype()
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 16
        raw = self.real(GUARD + nbytes + pad + GUARD, dtype=torch.uint8, device=dev)
        raw.fill_(PATTERN)
        self.bufs.append((raw, nbytes))
        return raw[GUARD:GUARD + nbytes].view(dtype).view(shape)

    def check(self, what):
        torch.cuda.synchronize()
        assert self.bufs, 'no device buffer was allocated through torch.empty'
        for i, (raw, nbytes) in enumerate(self.bufs):
            lo, hi = raw[:GUARD], raw[GUARD + nbytes:]
            assert bool((lo == PATTERN).all()), '%s: buffer %d (%d bytes): bytes BEFORE it were written' % (what, i, nbytes)
            assert bool((hi == PATTERN).all()), '%s: buffer %d (%d bytes): bytes AFTER it were written' % (what, i, nbytes)


def _same(a, b):
    if isinstance(a, (tuple, list)):
        return all(_same(x, y) for x, y in zip(a, b))
    if not torch.is_tensor(a):
        return a == b
    return a.shape == b.shape and bool(((a == b) | ((a != a) & (b != b))).all())


def _run(what, fn):
    """fn() on ordinary buffers, then under guard bands: same results, bands intact."""
    vod._lib.load()
    want = fn()
    torch.cuda.synchronize()
    ops._ws.__init__()                  # drop the cached workspaces: they are re-allocated under the guard
    with GuardedEmpty() as g:
        got = fn()
        g.check(what)
    ops._ws.__init__()
    assert _same(got, want), '%s: results differ between ordinary and guarded buffers' % what


def test_roi_align_bounds():
    g = torch.Generator().manual_seed(1)
    for C, K, B in ((36, 37, 2), (192, 5, 1), (512, 301, 3), (64, 1, 1)):
        x = torch.randn(B, C, 13, 21, generator=g).to(DEV)
        rois = rpn_like_rois(g, K, 1, 21 * 16., 13 * 16.)
        rois[:, 0] = torch.randint(0, B, (K,), generator=g).float()
        rois = rois.to(DEV)
        _run('roi_align C=%d K=%d' % (C, K), lambda: ops.roi_align(x, rois, 7, 1 / 16., 2, 'avg', True))
        _run('roi_align channels_last C=%d K=%d' % (C, K), lambda: ops.roi_align(x, rois, 7, 1 / 16., 2, 'avg', True, channels_last_out=True).contiguous())
        _run('roi_align adaptive C=%d K=%d' % (C, K), lambda: ops.roi_align(x, rois, (3, 5), 1 / 16., 0, 'avg', False))


def test_warp_and_embed_bounds():
    g = torch.Generator().manual_seed(2)
    for T, C, H, W in ((3, 20, 9, 11), (5, 64, 38, 63), (1, 7, 5, 3)):
        x = torch.randn(T, C, H, W, generator=g).to(DEV)
        flow = (torch.randn(T, 2, H * 16, W * 16, generator=g) * 9).to(DEV)
        key = torch.randn(1, C, H, W, generator=g).to(DEV)
        _run('flow_warp', lambda: ops.flow_warp(x, flow))
        ke, re_ = torch.randn(1, C, H, W, generator=g).to(DEV), torch.randn(T, C, H, W, generator=g).to(DEV)
        _run('embed_weighted_sum', lambda: ops.embed_weighted_sum(ke, re_, x))
        _run('fgfa_warp_weighted_sum', lambda: ops.fgfa_warp_weighted_sum(ke, re_, x, flow, key_x=key, key_slot=T - 1))


def test_selsa_bounds():
    g = torch.Generator().manual_seed(3)
    for N, M, heads, d in ((37, 115, 16, 64), (300, 4501, 16, 64), (5, 3, 2, 8), (129, 64, 4, 64)):
        q, k, v = (torch.randn(n, heads * d, generator=g).to(DEV) for n in (N, M, M))
        _run('selsa N=%d M=%d d=%d' % (N, M, d), lambda: ops.selsa_attention(q, k, v, heads))
        if d == 64:
            ld = (M + 3) // 4 * 4
            vt = torch.zeros(heads * d, ld, device=DEV)
            vt[:, :M] = v.t()
            _run('selsa V^T N=%d M=%d' % (N, M), lambda: ops.selsa_attention(q, k, vt, heads, v_transposed=True))
            _run('selsa bf16 N=%d M=%d' % (N, M), lambda: ops.selsa_attention(q.bfloat16(), k.bfloat16(), v.bfloat16(), heads))


def test_nms_and_decode_bounds():
    g = torch.Generator().manual_seed(4)
    for n in (1, 63, 65, 1000, 9001):
        boxes = clustered_boxes(g, n, max(1, n // 7)).to(DEV)
        scores = torch.rand(n, generator=g).to(DEV)
        ids = torch.randint(0, 30, (n,), generator=g).to(DEV)
        _run('batched_nms n=%d' % n, lambda: ops.batched_nms(boxes, scores, ids, dict(type='nms', iou_threshold=0.5)))
        _run('nms n=%d' % n, lambda: ops.nms(boxes, scores, 0.7, max_num=min(n, 300)))


def test_temporal_roi_align_and_head_bounds():
    """The whole SELSA + TemporalRoIAlign step (msra GEMM / re-score / overflow re-scan, TAFA logits + weighting in both modes,
    RoIAlign, SELSA layers with the fused tail, decode + NMS) at ragged sizes, uncached and through the reference-frame cache."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    for C, T, N, heads in ((64, 3, 37, 4), (128, 9, 21, 4), (48, 3, 10, 4)):
        H, W, D = 11, 19, 128
        torch.manual_seed(0)
        head = vod.SelsaRoIHead(
            bbox_roi_extractor=dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=heads,
                                    roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2), out_channels=C, featmap_strides=[16]),
            bbox_head=dict(type='SelsaBBoxHead', num_shared_fcs=2, in_channels=C, fc_out_channels=D, num_classes=6,
                           aggregator=dict(type='SelsaAggregator', in_channels=D, num_attention_blocks=2)),
            test_cfg=dict(score_thr=0.0001, nms=dict(type='nms', iou_threshold=0.5), max_per_img=100)).to(DEV).eval()
        ref_x = torch.relu(torch.randn(T, C, H, W, generator=g)).to(DEV)
        x = ref_x[T - 1:T].clone()
        props = [rpn_like_rois(g, N, 1, W * 16., H * 16.)[:, 1:].to(DEV)]
        ref_props = [rpn_like_rois(g, N, 1, W * 16., H * 16.)[:, 1:].to(DEV) for _ in range(T)]
        metas = [dict(img_shape=(H * 16, W * 16, 3), scale_factor=(1., 1., 1., 1.))]
        _run('SelsaRoIHead.simple_test C=%d T=%d N=%d' % (C, T, N),
             lambda: head.simple_test((x,), (ref_x,), props, ref_props, metas))
        ref_metas = [dict(video_id=1, frame_id=t, img_shape=(H * 16, W * 16, 3), scale_factor=(1., 1., 1., 1.)) for t in range(T)]

        def cached():
            head._clip_cache = None
            outs = [head.simple_test((x,), (ref_x,), props, ref_props, [ref_metas[T - 1]], ref_img_metas=ref_metas) for _ in range(2)]
            return outs[1]
        _run('SelsaRoIHead.simple_test with the cache C=%d T=%d N=%d' % (C, T, N), cached)
        # low-contrast maps: every (row, frame) pair goes through the overflow re-scan lists
        flat = (ref_x * 0.02 + 0.5)
        _run('TemporalRoIAlign on low-contrast maps C=%d' % C,
             lambda: head.bbox_roi_extractor((flat[T - 1:T],), vod.bbox2roi(props), ref_feats=(flat,)).contiguous())


def test_denoise_kernels_bounds():
    g = torch.Generator().manual_seed(6)
    taf = vod.TemporalAttentionFusion(16, 8, emb_nums=2).to(DEV).eval()
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.weight, 0, 0.05)
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.bias, 0, 0.5)
    for T, H, W in ((3, 9, 11), (2, 17, 30)):
        xt = torch.randn(T, 16, H, W, generator=g).to(DEV)
        with torch.no_grad():
            _run('TemporalAttentionFusion T=%d %dx%d' % (T, H, W), lambda: taf(xt))


def test_ops_concurrent_on_two_streams():
    """SURVEY 8b threading contract: the entry points keep no device-global state and take their stream explicitly, so the same
    operator may run on two streams at once (per-stream workspaces, `_lib.Workspace`).  Each operator is issued alternately on
    two streams with two different inputs, several rounds deep, and every result must equal the one computed alone."""
    g = torch.Generator().manual_seed(9)
    T, C, H, W, N = 3, 64, 12, 20, 37
    cases = []
    for seed in range(2):
        maps = torch.relu(torch.randn(T, C, H, W, generator=g)).to(DEV)
        nhwc, norm, unit = ops._to_nhwc(maps, True, True)
        rois = rpn_like_rois(g, N, 1, W * 16., H * 16.).to(DEV)
        rows = ops.roi_align_nhwc(nhwc[T - 1:T].contiguous(), rois, 7, 1 / 16., 2, True, out_nhwc=True).view(N * 49, C)
        q, k, v = (torch.randn(n, 1024, generator=g).to(DEV) for n in (150, 700, 700))
        boxes, scores = clustered_boxes(g, 3000, 300).to(DEV), torch.rand(3000, generator=g).to(DEV)
        ids = torch.randint(0, 30, (3000,), generator=g).to(DEV)
        x_all = torch.randn(T + 1, N, 49, C, generator=g).to(DEV)
        emb = torch.randn(T + 1, N, 49, C, generator=g).to(DEV)
        cases.append(dict(nhwc=nhwc.contiguous(), norm=norm, unit=unit, rois=rois, rows=rows, q=q, k=k, v=v, boxes=boxes,
                          scores=scores, ids=ids, x_all=x_all, emb=emb))
    fns = {
        'roi_align': lambda c: ops.roi_align_nhwc(c['nhwc'], torch.cat([c['rois']] * 8), 7, 1 / 16., 2, True),
        'msra_topk_sample': lambda c: ops.msra_topk_sample(c['rows'], c['nhwc'], 2, ref_norm=c['norm'], ref_unit=c['unit']),
        'selsa_attention': lambda c: ops.selsa_attention(c['q'], c['k'], c['v'], 16),
        'batched_nms': lambda c: ops.batched_nms(c['boxes'], c['scores'], c['ids'], dict(type='nms', iou_threshold=0.5))[1],
        'tafa_weighted_sum': lambda c: ops.tafa_weighted_sum(c['x_all'], c['emb'], 4),
    }
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for name, fn in fns.items():
        want = [fn(c) for c in cases]
        torch.cuda.synchronize()
        got = [[], []]
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for rnd in range(6):
            for lane in (0, 1):
                with torch.cuda.stream(streams[lane]):
                    got[lane].append(fn(cases[lane]))
        torch.cuda.synchronize()
        for lane in (0, 1):
            for out in got[lane]:
                assert _same(out, want[lane]), '%s: result on stream %d differs from the result computed alone' % (name, lane)
