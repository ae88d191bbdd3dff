"""tcgen05 / TMEM / TMA building blocks in isolation (run first: everything tensor-core depends on them)."""
import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(128, 128, 64), (256, 384, 512), (300, 200, 96), (77, 130, 40)])
def test_gemm_bf16(shape):
    M, N, K = shape
    g = torch.Generator(device='cuda').manual_seed(1)
    a = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    b = torch.randn(N, K, device='cuda', generator=g).bfloat16()
    d = ops.test_gemm_nt(a, b)
    ref = a.float() @ b.float().t()
    err = (d - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err   # bf16 products are exact in fp32; only accumulation order differs


@pytest.mark.parametrize('shape', [(128, 128, 64), (256, 384, 512), (300, 200, 128), (77, 130, 192)])
def test_gemm_bf16_a_in_tmem(shape):
    """TS-form MMA: A rows written to TMEM (lane = row, two bf16 per column) with tcgen05.st."""
    M, N, K = shape
    g = torch.Generator(device='cuda').manual_seed(4)
    a = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    b = torch.randn(N, K, device='cuda', generator=g).bfloat16()
    d = ops.test_gemm_nt(a, b, a_in_tmem=True)
    ref = a.float() @ b.float().t()
    err = (d - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err


@pytest.mark.parametrize('shape', [(128, 128, 32), (256, 256, 64), (300, 200, 100)])
def test_gemm_tf32(shape):
    M, N, K = shape
    g = torch.Generator(device='cuda').manual_seed(2)
    a = torch.randn(M, K, device='cuda', generator=g)
    b = torch.randn(N, K, device='cuda', generator=g)
    d = ops.test_gemm_nt(a, b)
    ref = a.double() @ b.double().t()
    err = (d.double() - ref).abs().max().item() / ref.abs().max().item()
    bias = ((d.double() - ref) * ref.sign()).mean().item() / ref.abs().mean().item()
    print('tf32 gemm', shape, 'max rel err %.3e' % err, 'signed bias %.3e' % bias)
    assert err < 2e-3, err   # tf32: 10-bit mantissa operands, fp32 accumulate
    assert abs(bias) < 1e-4, bias   # operands are ROUNDED to tf32 by the TMA unit, not truncated


# ------------------------------------------------------------------------------------------ (3) SELSA on tcgen05
from oracle import vod_oracle as O  # noqa: E402
from conftest import params, rel_err  # noqa: E402

SELSA_TF32_TOL = 1e-3    # north_star bar: aggregated features within 1e-3 relative error in fp32 I/O
SELSA_BF16_TOL = 2e-2    # bf16 operands (8-bit mantissa): stated separately


# M = 20 / 33: the second softmax group of the kernel sees no (or one) reference row; M = 70: a 6-row tail chunk
@pytest.mark.parametrize('N,M,heads', [(37, 300, 2), (128, 64, 1), (300, 900, 16), (300, 4500, 16), (130, 1000, 4), (1, 70, 3),
                                       (40, 20, 2), (129, 33, 1)])
def test_selsa_attention_tc_vs_oracle(N, M, heads):
    g = torch.Generator().manual_seed(N + M)
    D = heads * 64
    q, k, v = (torch.randn(n, D, generator=g) for n in (N, M, M))
    ref = O.selsa_attention(q, k, v, heads)
    out = ops.selsa_attention(q.to('cuda'), k.to('cuda'), v.to('cuda'), heads, impl=ops.IMPL_TC)
    print('selsa tf32', (N, M, heads), 'rel err %.3e' % rel_err(out, ref))
    assert rel_err(out, ref) < SELSA_TF32_TOL, rel_err(out, ref)
    # V handed over pre-transposed (what SelsaAggregator does), padded leading dimension
    ld = (M + 7) // 8 * 8
    vt = torch.zeros(D, ld)
    vt[:, :M] = v.t()
    out_t = ops.selsa_attention(q.to('cuda'), k.to('cuda'), vt.to('cuda'), heads, v_transposed=True, impl=ops.IMPL_TC)
    assert rel_err(out_t, ref) < SELSA_TF32_TOL
    # SIMT kernel cross-check on the same inputs (exact fp32)
    out_s = ops.selsa_attention(q.to('cuda'), k.to('cuda'), v.to('cuda'), heads, impl=ops.IMPL_SIMT)
    assert rel_err(out_s, ref) < 2e-5
    # bf16 operands
    out_b = ops.selsa_attention(q.to('cuda').bfloat16(), k.to('cuda').bfloat16(), v.to('cuda').bfloat16(), heads, impl=ops.IMPL_TC)
    ref_b = O.selsa_attention(q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float(), heads)
    assert rel_err(out_b, ref_b) < SELSA_BF16_TOL, rel_err(out_b, ref_b)


def test_selsa_attention_tc_large_logits():
    """rows whose maximum moves between chunks exercise the running-max correction."""
    g = torch.Generator().manual_seed(3)
    N, M, heads = 200, 2000, 2
    q = torch.randn(N, 128, generator=g) * 3
    k = torch.randn(M, 128, generator=g) * 3
    k[M // 2:] *= 2.0
    v = torch.randn(M, 128, generator=g)
    ref = O.selsa_attention(q.double(), k.double(), v.double(), heads).float()
    out = ops.selsa_attention(q.to('cuda'), k.to('cuda'), v.to('cuda'), heads, impl=ops.IMPL_TC)
    assert rel_err(out, ref) < 2e-2   # tf32 logits of magnitude ~100: softmax amplifies operand rounding
    out_s = ops.selsa_attention(q.to('cuda'), k.to('cuda'), v.to('cuda'), heads, impl=ops.IMPL_SIMT)
    assert rel_err(out_s, ref) < 1e-3


def test_selsa_aggregator_golden_tc(golden):
    m = vod.build_aggregator(dict(type='SelsaAggregator', in_channels=128, num_attention_blocks=2))
    m.load_state_dict(params(golden, 'selsa64_p.'))
    m = m.to('cuda')
    torch.backends.cuda.matmul.allow_tf32 = False
    out = m(golden['selsa64_x'].to('cuda'), golden['selsa64_ref_x'].to('cuda'))   # auto -> tcgen05 (d = 64)
    assert rel_err(out, golden['selsa64_out']) < SELSA_TF32_TOL


# ------------------------------------------------------------------------------------------ (4) most-similar on tcgen05
def _msra_compare(N, C, T, H, W, seed, k=2, zero_rows=(), zero_pixels=(), cluster=0, impl=None):
    g = torch.Generator().manual_seed(seed)
    roi = torch.relu(torch.randn(N, C, 7, 7, generator=g))
    ref = torch.relu(torch.randn(T, C, H, W, generator=g))
    if cluster:
        # a run of `cluster` neighbouring pixels per frame that all look like the RoI vectors (cosines within ~1e-4 of each
        # other, far above every other pixel): the candidate pass must keep the true top-k of such a run
        base = torch.relu(torch.randn(C, generator=g)) + 0.1
        roi = base.view(1, C, 1, 1) * (1 + 0.05 * torch.randn(N, C, 7, 7, generator=g))
        for t in range(T):
            y, x0 = (3 * t + 2) % H, (5 * t + 1) % (W - cluster)
            ref[t, :, y, x0:x0 + cluster] = base.view(C, 1) * (1 + 0.05 * torch.randn(C, cluster, generator=g))
    for n, p in zero_rows:                  # an all-zero RoI vector: the reference divides by its zero norm -> NaN row
        roi[n, :, p // 7, p % 7] = 0
    for t, y, x in zero_pixels:             # an all-zero reference pixel: NaN similarity for every RoI row of frame t
        ref[t, :, y, x] = 0
    out0, idx0, sim0 = O.most_similar_roi_align(roi, ref, k, return_indices=True)
    ref_nhwc, norm, unit = ops.to_nhwc(ref.to('cuda'), want_norm=True, want_unit_bf16=True)
    rows = roi.permute(0, 2, 3, 1).reshape(N * 49, C).to('cuda')
    out1, idx1, val1 = ops.msra_topk_sample(rows, ref_nhwc, k, ref_norm=norm, ref_unit=unit,
                                            impl=ops.IMPL_TC if impl is None else impl, return_indices=True)
    idx1 = idx1.cpu().long()
    got = out1.view(T, N, 7, 7, C).permute(0, 1, 4, 2, 3)
    nan_rows = torch.isnan(out0).any(dim=2).permute(1, 2, 3, 0).reshape(N * 49, T)      # [NP, T]
    same = (idx1.sort(dim=2).values == idx0.sort(dim=2).values).all(dim=2) | nan_rows
    for r, t in (~same).nonzero().tolist():
        # tie tolerance: a differing location must have the same fp32 similarity up to 1e-6
        v_ours = sim0[r, t, idx1[r, t]].sort().values
        v_ref = sim0[r, t, idx0[r, t]].sort().values
        assert (v_ours - v_ref).abs().max() <= 1e-6, (r, t, v_ours, v_ref)
    # sampled features where the location sets agree (an fp32 tie that flips picks another pixel); also asserts that
    # the NaN patterns are identical
    ok = same.t().reshape(T, N, 1, 7, 7)
    assert rel_err(torch.where(ok, got.cpu(), out0), out0) < 1e-3
    return float(same.float().mean())


# C = 64 -> generic re-score (C % 128 != 0); C = 128/256/384/512 -> the four instantiations of the lean re-score kernel;
# 294, 980, 147, 245, 196 rows -> odd and even numbers of 128-row tiles for the CTA pairs
@pytest.mark.parametrize('N,C,T,H,W', [(6, 64, 3, 12, 20), (20, 512, 3, 38, 63), (3, 128, 2, 9, 15), (5, 256, 2, 10, 17),
                                       (4, 384, 3, 11, 13)])
def test_msra_tc_vs_oracle(N, C, T, H, W):
    frac = _msra_compare(N, C, T, H, W, N + C)
    assert frac > 0.999


@pytest.mark.parametrize('k,C', [(1, 512), (1, 128), (3, 512), (4, 256), (4, 64)])
def test_msra_other_k(k, C):
    """num_most_similar_points != 2: k = 1 runs the tensor-core path, k > 2 the exact scan (the candidate lists are sized
    for k <= 2; asking for the tensor-core path explicitly is refused)."""
    frac = _msra_compare(4, C, 3, 13, 21, 100 * k + C, k=k, impl=ops.IMPL_TC if k == 1 else ops.IMPL_AUTO)
    assert frac > 0.999
    if k > 2:
        with pytest.raises(vod.VodError):
            _msra_compare(4, C, 3, 13, 21, 1, k=k, impl=ops.IMPL_TC)


@pytest.mark.parametrize('cluster', [6, 12, 16])
def test_msra_tc_neighbouring_near_ties(cluster):
    """Smooth feature maps put several near-tied maxima next to each other.  The four top-4 lists of the tensor-core
    pass take every 4th location each, so a run of up to 16 neighbours is kept completely (a list per 32 CONSECUTIVE
    locations would drop all but 4 of them and pick the top-2 among those by bf16 noise)."""
    frac = _msra_compare(3, 512, 3, 20, 40, 900 + cluster, cluster=cluster)
    # (every differing pick was checked to be an fp32 tie, <= 1e-6 in similarity, inside _msra_compare)
    assert frac > 0.99


@pytest.mark.parametrize('C', [512, 64])
def test_msra_tc_zero_norm_vectors(C):
    """The reference has no epsilon: a zero RoI vector gives a NaN row in every frame, a zero reference pixel makes the
    similarity NaN for every RoI row of that frame (torch.topk ranks NaN first).  Same NaN pattern, same finite rows."""
    frac = _msra_compare(3, C, 3, 12, 19, 5 + C, zero_rows=((0, 0), (2, 30)), zero_pixels=((1, 4, 7),))
    assert frac > 0.999


def test_temporal_roi_align_golden_tc(golden):
    torch.backends.cudnn.allow_tf32 = False
    m = vod.build_roi_extractor(dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                     roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=64, featmap_strides=[16]))
    m.load_state_dict(params(golden, 'troi_p.'))
    m = m.to('cuda')
    out = m((golden['troi_feat'].to('cuda'),), golden['troi_rois'].to('cuda'), ref_feats=(golden['troi_ref'].to('cuda'),))
    assert rel_err(out, golden['troi_out']) < 1e-4


def test_msra_candidate_recall_full_scale():
    """cfg-3 scale (N=300 RoIs x 49 bins, T=15 frames of 38x63): the true fp32 top-2 of every (row, frame) must be
    among the 16 candidates the bf16 tensor-core pass keeps, and the end result must pick exactly them."""
    g = torch.Generator().manual_seed(77)
    N, C, T, H, W = 300, 512, 15, 38, 63
    ref = torch.relu(torch.randn(T, C, H, W, generator=g))
    roi = torch.relu(torch.randn(N, C, 7, 7, generator=g)) + 0.3 * ref[T - 1, :, :7, :7]   # correlated with the key frame
    out0, idx0, sim0 = O.most_similar_roi_align(roi, ref, 2, return_indices=True)
    ref_nhwc, norm, unit = ops.to_nhwc(ref.to('cuda'), want_norm=True, want_unit_bf16=True)
    rows = roi.permute(0, 2, 3, 1).reshape(N * 49, C).to('cuda')
    roi_unit = torch.nn.functional.normalize(rows, dim=1).bfloat16()
    cand = ops.msra_gemm_candidates(roi_unit, unit, T).cpu()
    locs = (cand & 0xFFF).long()                                  # [NP, T, 16]
    valid = cand != 0
    hit = ((locs.unsqueeze(-1) == idx0.unsqueeze(2)) & valid.unsqueeze(-1)).any(dim=2)   # [NP, T, 2]
    assert bool(hit.all()), 'recall %.6f' % hit.float().mean()
    out1, idx1, _ = ops.msra_topk_sample(rows, ref_nhwc, 2, ref_norm=norm, ref_unit=unit, impl=ops.IMPL_TC, return_indices=True)
    idx1 = idx1.cpu().long()
    same = (idx1.sort(dim=2).values == idx0.sort(dim=2).values).all(dim=2)
    for r, t in (~same).nonzero().tolist():
        v_ours = sim0[r, t, idx1[r, t]].sort().values
        v_ref = sim0[r, t, idx0[r, t]].sort().values
        assert (v_ours - v_ref).abs().max() <= 1e-6, (r, t)
    assert same.float().mean() > 0.9999
    # sampled features: compared where the location sets agree (an fp32 tie that flips picks a different pixel)
    got = out1.view(T, N * 49, C).cpu()
    want = out0.permute(0, 1, 3, 4, 2).reshape(T, N * 49, C)
    ok = same.t().unsqueeze(-1)                                  # [T, NP, 1]
    assert rel_err(torch.where(ok, got, want), want) < 1e-4


def test_msra_unit_operand_without_padding_is_copied():
    """HW % 4 != 0: the tensor-core pass may touch 3 rows past the unit-norm copy (include/vodagg.h).  ``to_nhwc`` allocates
    them; a caller-made tensor without slack (clone) is copied into a padded buffer by the Python layer."""
    g = torch.Generator().manual_seed(5)
    C, T, H, W = 128, 2, 9, 15                       # HW = 135
    ref = torch.relu(torch.randn(T, C, H, W, generator=g)).to('cuda')
    rows = torch.relu(torch.randn(3 * 49, C, generator=g)).to('cuda')
    nh, norm, unit = ops.to_nhwc(ref, want_norm=True, want_unit_bf16=True)
    assert unit.untyped_storage().nbytes() >= (T * H * W + 3) * C * 2
    tight = unit.clone()
    assert ops._padded_unit(unit, T * H * W).data_ptr() == unit.data_ptr()
    assert ops._padded_unit(tight, T * H * W).data_ptr() != tight.data_ptr()
    a = ops.msra_topk_sample(rows, nh, 2, ref_norm=norm, ref_unit=unit, impl=ops.IMPL_TC, return_indices=True)
    b = ops.msra_topk_sample(rows, nh, 2, ref_norm=norm, ref_unit=tight, impl=ops.IMPL_TC, return_indices=True)
    c = ops.msra_topk_sample(rows, nh, 2, impl=ops.IMPL_TC, return_indices=True)     # norms / unit copy made by the library
    for x, y in ((a, b), (a, c)):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
