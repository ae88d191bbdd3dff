"""Parity at BASELINE.json's FULL sizes against the CPU oracle (round 2):

* the cfg-3 key-frame step (N=300, T=15, C=512, D=1024, 3 layers) in exactly the configuration ``bench.py`` times -- built by
  ``bench.build_head`` / ``bench.make_inputs``, library GEMMs/convs in tf32 -- and again with fp32 library math; tolerances are
  stated per tensor below and the measured errors are written to ``gpurun_out/parity_fullsize.json``;
* the cfg-5 sweep extreme (N=1000, T=31) on slices the oracle finishes in seconds: SELSA attention rows and TemporalRoIAlign
  RoIs, instead of comparing the repo with itself.
"""
import json
import os

import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops
from oracle import vod_oracle as O

from conftest import ROOT, rel_err
from helpers import rpn_like_rois

pytestmark = pytest.mark.gpu
DEV = 'cuda'

# stated tolerances of the full-size step, max|a-b| / max|b| against the fp32 CPU oracle, over the RoIs whose sampled
# locations are the oracle's (a pick that differs must be an fp32 TIE, verified below, and is then excluded)
#   fp32 library math: only our own tcgen05 kernels deviate from fp32 (tf32 SELSA attention, north_star bar 1e-3 per layer)
#   tf32 library math (the benchmarked configuration): fc_0 (K = 25088), the per-layer projections, the key-slot embed conv and
#   the G product additionally round their operands to tf32 (10-bit mantissa, relative 4.9e-4 per operand)
TOL = {
    False: dict(bbox_feats=1e-5, cls_score=1e-4, bbox_pred=1e-4, det_match=0.97),   # measured r02: 4.7e-7 / 6.8e-6 / 5.7e-6 / 1.0
    True: dict(bbox_feats=1e-3, cls_score=5e-3, bbox_pred=5e-3, det_match=0.9),       # measured r02: 9.2e-5 / 9.7e-4 / 8.3e-4 / 1.0
}
TIE = 2e-6          # two similarities closer than this are a tie in fp32 (512-term dot products of magnitude <= 1)
_cache = {}


def _record(name, value):
    out = os.path.join(ROOT, 'gpurun_out')
    if not os.path.isdir(out):
        return
    path = os.path.join(out, 'parity_fullsize.json')
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = value
    json.dump(data, open(path, 'w'), indent=1, sort_keys=True)


def _oracle_top3(roi_feats, ref):
    """The oracle's similarities (temporal_roi_align.py:127-145) reduced frame by frame to the top-3 values / locations, and a
    closure that evaluates the oracle's similarity of arbitrary (row, frame, location) picks."""
    roi_e = roi_feats / roi_feats.norm(p=2, dim=1, keepdim=True)
    ref_e = ref / ref.norm(p=2, dim=1, keepdim=True)
    C = roi_e.shape[1]
    a = roi_e.permute(0, 2, 3, 1).reshape(-1, C)
    vals, idxs = [], []
    for t in range(ref.shape[0]):
        sim = a.mm(ref_e[t].reshape(C, -1))
        v, i = sim.topk(3, dim=1)
        vals.append(v); idxs.append(i)

    def sim_of(rows, frames, locs):
        b = ref_e.reshape(ref.shape[0], C, -1)[frames, :, locs]            # [m, C]
        return (a[rows] * b).sum(dim=1)
    return torch.stack(vals, 1), torch.stack(idxs, 1), sim_of


def _cfg3_oracle():
    if 'cpu' not in _cache:
        import bench
        cfg = bench.CONFIGS['cfg3']
        ref_x, props_all = bench.make_inputs(cfg, 0)
        sd = bench.cpu_head_state(cfg)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            want = bench.cpu_step(cfg, sd, ref_x, props_all, return_all=True)
            rois, _ = bench.step_rois(cfg, props_all)
            key_feats = O.roi_align(ref_x[cfg['T'] - 1:], rois, 7, 1 / 16, 2, True)
            want['top3_val'], want['top3_idx'], want['sim_of'] = _oracle_top3(key_feats, ref_x)
        _cache['cpu'] = (cfg, ref_x, props_all, sd, want)
    return _cache['cpu']


def _tie_explained_rois(head, cfg, ref_x, props_all, want):
    """Runs our most-similar search, and proves that every (row, frame) whose selected locations differ from the oracle's is an
    fp32 tie: by the ORACLE's own similarities, our picks are within TIE of its picks.  Returns the RoIs that contain one."""
    rois, _ = bench_mod().step_rois(cfg, props_all, DEV)
    st = ref_x.to(DEV)
    x_all, idx, val = head.bbox_roi_extractor._stack_key_and_refs(st[cfg['T'] - 1:], rois, st, return_indices=True)
    idx1 = idx.cpu().long()                                               # [NP, T, 2] descending similarity
    idx0 = want['top3_idx'][:, :, :2]
    differs = (idx1 != idx0).any(dim=2)
    rows, frames = differs.nonzero(as_tuple=True)
    for j in range(2):
        ours = want['sim_of'](rows, frames, idx1[rows, frames, j])
        theirs = want['top3_val'][rows, frames, j]
        worst = float((ours - theirs).abs().max()) if len(rows) else 0.0
        assert worst <= TIE, 'a differing pick is not an fp32 tie: oracle similarity differs by %.3e' % worst
    return sorted(set((rows // 49).tolist())), int(differs.sum()), int(differs.numel())


def bench_mod():
    import bench
    return bench


@pytest.mark.parametrize('tf32', [True, False])
def test_cfg3_full_size_step_vs_oracle(tf32):
    """mmtracking/mmtrack/models/roi_heads/selsa_roi_head.py:80-97,147-187 at N=300, T=15 -- the step bench.py times."""
    import bench
    cfg, ref_x, props_all, sd, want = _cfg3_oracle()
    head = bench.build_head(cfg, torch.device(DEV))
    # same random-init weights on both sides (bench.build_head seeds the generator)
    for k, v in head.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    with bench.library_math(tf32):
        ours = bench.gpu_step_outputs(head, cfg, ref_x.to(DEV), props_all.to(DEV))
        # ... and the captured graph the timed loop replays gives the same detections as the eager call
        rois, ref_rois = bench.step_rois(cfg, props_all, DEV)
        st_ref = ref_x.to(DEV)
        graph, (g_dets, g_labels, g_count) = head.capture_graph((st_ref[cfg['T'] - 1:],), (st_ref,), rois, ref_rois, bench.IMG_SHAPE,
                                                                (1., 1., 1., 1.))
        graph.replay()
        torch.cuda.synchronize()
        tie_rois, n_diff, n_pairs = _tie_explained_rois(head, cfg, ref_x, props_all, want)
    n = int(g_count.item())
    raw = bench.parity_report(ours, want)                               # what bench.py prints (no exclusions)
    rep = bench.parity_report(ours, want, exclude_rois=tie_rois)
    rep['pairs_with_tied_picks'] = [n_diff, n_pairs]
    rep['graph_vs_eager_det_match'] = bench.match_detections(g_dets[:n].cpu(), g_labels[:n].cpu(), ours['dets'], ours['labels'],
                                                             box_tol=0.05, score_tol=1e-4)
    _record('cfg3_step_library_%s' % ('tf32' if tf32 else 'fp32'), dict(excluding_verified_ties=rep, all_rois=raw))
    print('cfg3 full-size parity (library math %s):' % ('tf32' if tf32 else 'fp32'), rep, raw)
    tol = TOL[tf32]
    assert n_diff <= 0.002 * n_pairs, (n_diff, n_pairs)                 # ties are rare even on iid noise
    assert rep['bbox_feats_rel_err'] < tol['bbox_feats'], rep
    assert rep['cls_score_rel_err'] < tol['cls_score'], rep
    assert rep['bbox_pred_rel_err'] < tol['bbox_pred'], rep
    assert rep['det_match'] >= tol['det_match'], rep
    assert rep['graph_vs_eager_det_match'] >= 0.95, rep
    assert not any(k.endswith('nan_mismatch') for k in rep)


def test_selsa_sweep_size_rows_vs_oracle():
    """cfg-5 sweep extreme, N=1000 proposals x T=31 frames (M=31000 reference proposals), 16 heads: two 128-row slices of
    the tensor-core kernel's output against the CPU oracle (selsa_aggregator.py:51-70)."""
    g = torch.Generator().manual_seed(42)
    N, M, heads = 1000, 31000, 16
    q = torch.randn(N, 1024, generator=g)
    k = torch.randn(M, 1024, generator=g)
    v = torch.randn(M, 1024, generator=g)
    out = ops.selsa_attention(q.to(DEV), k.to(DEV), v.to(DEV), heads, impl=ops.IMPL_TC).cpu()
    errs = []
    for lo in (0, 872):
        want = O.selsa_attention(q[lo:lo + 128], k, v, heads)
        errs.append(rel_err(out[lo:lo + 128], want))
    _record('selsa_sweep_rows', errs)
    assert max(errs) < 1e-3, errs           # north_star: aggregated features within 1e-3 (fp32 I/O, tf32 tensor-core math)
    out_s = ops.selsa_attention(q.to(DEV), k.to(DEV), v.to(DEV), heads, impl=ops.IMPL_SIMT).cpu()
    assert rel_err(out_s[:128], O.selsa_attention(q[:128], k, v, heads)) < 2e-5


def test_temporal_roi_align_sweep_size_rois_vs_oracle():
    """N=1000 proposals, T=31 reference maps: 32 of the RoIs against the oracle's TemporalRoIAlign
    (temporal_roi_align.py:99-207; RoIs are independent, so the oracle runs on the 32 alone)."""
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(43)
    m = vod.build_roi_extractor(dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                     roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=512, featmap_strides=[16])).to(DEV)
    T, N = 31, 1000
    ref = torch.relu(torch.randn(T, 512, 38, 63, generator=g))
    rois = rpn_like_rois(g, N, 1)
    pick = torch.cat([torch.arange(0, 16), torch.arange(N - 16, N)])
    w, b = m.embed_network.conv.weight.detach().cpu(), m.embed_network.conv.bias.detach().cpu()
    want = O.temporal_roi_align(ref[T - 1:T], rois[pick], ref, w, b, 2, 4)
    import bench
    errs = {}
    for tf32 in (False, True):
        with bench.library_math(tf32):
            out = m((ref[T - 1:T].to(DEV),), rois.to(DEV), ref_feats=(ref.to(DEV),))
        assert out.shape == (N, 512, 7, 7)
        errs['tf32' if tf32 else 'fp32'] = rel_err(out[pick.to(DEV)], want)
    _record('troi_sweep_rois', errs)
    assert errs['fp32'] < 1e-4, errs
    assert errs['tf32'] < 3e-3, errs
    # reference RoIs of all 31 frames: plain RoIAlign rows against the oracle
    ref_rois = rpn_like_rois(g, N, T)
    rf = m((ref.to(DEV),), ref_rois.to(DEV))
    sel = torch.tensor([0, 999, 15000, 30999])
    assert rel_err(rf[sel.to(DEV)], O.roi_align(ref, ref_rois[sel], 7, 1 / 16, 2, True)) < 2e-5


def test_nms_nan_scores_do_not_corrupt_the_sort():
    """a NaN score compares false everywhere; the counting sort must still produce a permutation.  NaN ranks first
    (torch.sort(descending) order); the finite boxes behave exactly as without the NaN boxes when nothing overlaps them."""
    g = torch.Generator().manual_seed(5)
    n = 500
    boxes = torch.rand(n, 2, generator=g) * 400
    boxes = torch.cat([boxes, boxes + torch.rand(n, 2, generator=g) * 60 + 4], 1)
    scores = torch.rand(n, generator=g)
    far = torch.tensor([[5000., 5000., 5010., 5010.], [6000., 6000., 6010., 6010.], [7000., 7000., 7010., 7010.]])
    b2 = torch.cat([boxes[:100], far[:1], boxes[100:], far[1:]], 0)
    s2 = torch.cat([scores[:100], torch.tensor([float('nan')]), scores[100:], torch.tensor([float('nan'), float('inf')])], 0)
    _, keep0 = O.batched_nms(boxes, scores, torch.zeros(n, dtype=torch.long), dict(type='nms', iou_threshold=0.5), class_agnostic=True)
    for _ in range(2):
        dets, keep = vod.nms(b2.to(DEV), s2.to(DEV), 0.5)
        keep = keep.cpu()
        assert keep.unique().numel() == keep.numel() and int(keep.min()) >= 0 and int(keep.max()) < n + 3
        assert keep[:3].tolist() == [100, n + 1, n + 2]            # NaN (index order), then +inf
        rest = keep[3:]
        mapped = torch.where(rest > 100, rest - 1, rest)            # indices of the finite boxes in the original numbering
        assert torch.equal(mapped, keep0)
    # all-NaN input: every box kept or suppressed deterministically, nothing out of range
    dets, keep = vod.nms(b2.to(DEV), torch.full((n + 3,), float('nan'), device=DEV), 0.5)
    assert int(keep.min()) >= 0 and int(keep.max()) < n + 3 and keep.unique().numel() == keep.numel()
