"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/vodagg.h
declares, the registries / module signatures / parameter names mirror the reference, the reference's own
assertion behaviour is kept, the product path refuses to run without CUDA (no fallback) and never imports
the oracle."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'vodagg.h')).read()
    declared = set(re.findall(r'\b(vod_[a-z0-9_]+)\s*\(', header))
    declared.discard('vod_stream_t')
    assert len(declared) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), 'libvodagg.so does not export %s' % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.load().vod_version() >= 100
    assert _lib.load().vod_device_is_sm100() in (0, 1)   # no compute call without a GPU


def test_selftest_kernels_live_in_their_own_library():
    """The tcgen05 building-block GEMM (include/vodagg_selftest.h) is test code: exported by libvodagg_selftest.so and absent
    from the product library."""
    header = open(os.path.join(ROOT, 'include', 'vodagg_selftest.h')).read()
    declared = set(re.findall(r'\b(vod_[a-z0-9_]+)\s*\(', header))
    declared.discard('vod_last_error')               # mentioned in the header comment; shared helper, exported by both libraries
    assert declared == {'vod_test_gemm_nt'} and declared <= set(_lib.SELFTEST_SIGNATURES)
    st = _lib.load_selftest()
    product = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(st, name) and not hasattr(product, name)


def test_sass_is_blackwell_native():
    """The shipped cubin carries tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and TMA loads (UTMALDG)."""
    r = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in r.stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG'):
        assert mnemonic in r.stdout, mnemonic


def test_registries_and_signatures():
    assert set(vod.AGGREGATORS.module_dict) >= {'SelsaAggregator', 'EmbedAggregator'}
    assert set(vod.ROI_EXTRACTORS.module_dict) >= {'SingleRoIExtractor', 'TemporalRoIAlign'}
    m = vod.build_aggregator(dict(type='SelsaAggregator', in_channels=32, num_attention_blocks=4))
    assert list(m.state_dict()) == ['fc_embed.weight', 'fc_embed.bias', 'ref_fc_embed.weight', 'ref_fc_embed.bias',
                                    'fc.weight', 'fc.bias', 'ref_fc.weight', 'ref_fc.bias']
    assert list(inspect.signature(vod.SelsaAggregator.__init__).parameters) == ['self', 'in_channels', 'num_attention_blocks']
    e = vod.build_aggregator(dict(type='EmbedAggregator', num_convs=2, channels=8, kernel_size=3))
    assert list(e.state_dict()) == ['embed_convs.0.conv.weight', 'embed_convs.0.conv.bias',
                                    'embed_convs.1.conv.weight', 'embed_convs.1.conv.bias']
    assert list(inspect.signature(vod.EmbedAggregator.__init__).parameters) == \
        ['self', 'num_convs', 'channels', 'kernel_size', 'norm_cfg', 'act_cfg']
    t = vod.build_roi_extractor(dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                     roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=16, featmap_strides=[16]))
    assert list(t.state_dict()) == ['embed_network.conv.weight', 'embed_network.conv.bias']
    assert t.roi_layers[0].output_size == (7, 7) and abs(t.roi_layers[0].spatial_scale - 1 / 16) < 1e-12
    assert t.num_inputs == 1
    assert list(inspect.signature(vod.TemporalRoIAlign.forward).parameters) == ['self', 'feats', 'rois', 'roi_scale_factor', 'ref_feats']
    assert list(inspect.signature(vod.flow_warp_feats).parameters) == ['x', 'flow']
    assert list(inspect.signature(vod.batched_nms).parameters) == ['boxes', 'scores', 'idxs', 'nms_cfg', 'class_agnostic']
    assert list(inspect.signature(vod.RoIAlign.__init__).parameters) == \
        ['self', 'output_size', 'spatial_scale', 'sampling_ratio', 'pool_mode', 'aligned', 'use_torchvision']
    with pytest.raises(KeyError):
        vod.build_aggregator(dict(type='NoSuchAggregator'))
    with pytest.raises(KeyError):
        vod.AGGREGATORS.register_module()(vod.SelsaAggregator)     # duplicate without force
    vod.AGGREGATORS.register_module(force=True)(vod.SelsaAggregator)


def test_reference_assertions_are_kept():
    """mmtracking/tests/test_models/test_aggregators.py:9-20 and tests/test_core/test_motion_utils.py:13-29."""
    with pytest.raises(AssertionError):
        vod.EmbedAggregator(num_convs=0, channels=32, kernel_size=3)
    model = vod.EmbedAggregator(num_convs=3, channels=32, kernel_size=3)
    with pytest.raises(AssertionError):
        model(torch.randn(2, 32, 8, 8), torch.randn(4, 32, 8, 8))
    with pytest.raises(AssertionError):
        vod.flow_warp_feats(torch.randn(2, 8, 32, 32, 32), torch.randn(2, 2, 10, 10))
    with pytest.raises(AssertionError):
        vod.flow_warp_feats(torch.randn(2, 8, 32, 32), torch.randn(2, 2, 10, 10, 10))
    with pytest.raises(AssertionError):
        vod.flow_warp_feats(torch.randn(2, 8, 32, 32), torch.randn(2, 3, 10, 10))


def test_no_cpu_fallback():
    """The product path fails loudly on CPU tensors instead of silently computing somewhere else."""
    with pytest.raises(_lib.VodError):
        vod.flow_warp_feats(torch.randn(2, 8, 32, 32), torch.randn(2, 2, 10, 10))
    with pytest.raises(_lib.VodError):
        vod.SelsaAggregator(16, 4)(torch.randn(2, 16), torch.randn(4, 16))
    with pytest.raises(_lib.VodError):
        vod.batched_nms(torch.rand(4, 4), torch.rand(4), torch.zeros(4, dtype=torch.long), dict(type='nms', iou_threshold=0.5))
    with pytest.raises(_lib.VodError):
        vod.RoIAlign(7, 1 / 16, 2)(torch.randn(1, 4, 8, 8), torch.tensor([[0, 0., 0., 32., 32.]]))


def test_missing_library_is_loud(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'libvodagg.so'))
    with pytest.raises(_lib.VodError, match='no CPU / PyTorch fallback'):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'lowlightenvironmentvideoobjectdetection_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
                assert 'vod_oracle' not in src and 'ref_shim' not in src, f
    code = 'import sys; import lowlightenvironmentvideoobjectdetection_b200; assert not any(m.startswith("oracle") for m in sys.modules)'
    subprocess.run([sys.executable, '-c', code], check=True, cwd=ROOT)


def test_host_glue_matches_oracle_on_cpu():
    """delta2bbox / bbox2roi are plain torch (callers of the hot path): check them against the oracle's restatement."""
    from oracle import vod_oracle as O
    g = torch.Generator().manual_seed(1)
    rois = torch.rand(50, 4, generator=g) * 300
    rois[:, 2:] += rois[:, :2]
    deltas = torch.randn(50, 120, generator=g)
    a = vod.delta2bbox(rois, deltas, (0., 0., 0., 0.), (0.2, 0.2, 0.2, 0.2), max_shape=(600, 1000, 3))
    b = O.delta2bbox(rois, deltas, (0., 0., 0., 0.), (0.2, 0.2, 0.2, 0.2), max_shape=(600, 1000, 3))
    assert torch.equal(a, b)
    lst = [torch.rand(7, 4, generator=g) for _ in range(3)]
    r = vod.bbox2roi(lst)
    assert r.shape == (21, 5) and torch.equal(r[:, 0], torch.arange(3.).repeat_interleave(7))
    assert torch.equal(r[:, 1:], torch.cat(lst, 0))
    r2 = vod.bbox2roi([torch.rand(3, 4), torch.zeros(0, 4), torch.rand(2, 4)])
    assert r2.shape == (5, 5) and r2[:, 0].tolist() == [0, 0, 0, 2, 2]


def test_keyproj_weight_layout_and_identity_cpu(golden):
    """Host side of the key-projected TAFA logits (roi_extractors.TemporalRoIAlign._keyproj_weight): with the weight laid out as
    the C ABI documents (G [heads, N*P, C/cc, 9, cc]), contracting the raw RoI features with G = ek_head . W_head reproduces
    the reference's temporal_attentional_feature_aggregation (temporal_roi_align.py:44-97) on the golden TemporalRoIAlign
    parameters -- the identity the CUDA kernel relies on, checked here in plain torch (no kernel call)."""
    import torch
    import lowlightenvironmentvideoobjectdetection_b200 as vod
    from conftest import params, rel_err
    from oracle import vod_oracle as O
    m = vod.build_roi_extractor(dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                     roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=64, featmap_strides=[16]))
    m.load_state_dict(params(golden, 'troi_p.'))
    conv = m.embed_network.conv
    H, C, cc, P = 4, 64, 32, 49
    wr = m._keyproj_weight(conv, H, cc)
    assert wr.shape == (H, C // H, 9 * C)
    g = torch.Generator().manual_seed(5)
    T1, N = 5, 3
    x = torch.randn(1, N, C, 7, 7, generator=g)
    ref_x = torch.randn(T1 - 1, N, C, 7, 7, generator=g)
    want = O.tafa(x, ref_x, conv.weight.detach(), conv.bias.detach(), H)
    x_all = torch.cat((x, ref_x), 0)
    ek = torch.nn.functional.conv2d(x_all[0], conv.weight, conv.bias, padding=1).detach()
    ek_rows = ek.permute(0, 2, 3, 1).reshape(N * P, H, C // H)
    G = torch.bmm(ek_rows.transpose(0, 1), wr).view(H, N, P, C // cc, 9, cc)
    xr = x_all.permute(0, 1, 3, 4, 2)                                   # [T1, N, 7, 7, C]
    logits = torch.zeros(N, P, H, T1)
    for p in range(P):
        py, px = divmod(p, 7)
        for tap in range(9):
            qy, qx = py + tap // 3 - 1, px + tap % 3 - 1
            if 0 <= qy < 7 and 0 <= qx < 7:
                xv = xr[:, :, qy, qx, :].reshape(T1, N, C // cc, cc)
                logits[:, p] += torch.einsum('tnkc,hnkc->nht', xv, G[:, :, p, :, tap, :])
    w = torch.softmax(logits / (C // H) ** 0.5, dim=-1).repeat_interleave(C // H, dim=2)   # [N, P, C, T1]
    out = torch.einsum('tnpc,npct->npc', xr.reshape(T1, N, P, C), w).view(N, 7, 7, C).permute(0, 3, 1, 2)
    assert rel_err(out, want) < 1e-5
    # path selection: few stacked frames / unsupported convs take the full-embedding path
    assert m._keyproj_chunk(4, P, C) == 0                # below keyproj_min_frames
    m.keyproj = False
    assert m._keyproj_chunk(16, P, C) == 0


def test_keyproj_shape_gate_is_host_only():
    """vod_tafa_keyproj_chunk is pure host logic (no device call): which shapes the key-projected logits kernel accepts.
    Everything else must take the full-embedding path (vod_tafa_weighted_sum), never a silent partial result."""
    from lowlightenvironmentvideoobjectdetection_b200 import ops
    assert ops.tafa_keyproj_chunk(16, 49, 512, 4) == 32          # cfg 3
    assert ops.tafa_keyproj_chunk(32, 49, 512, 4) == 32          # sweep maximum (frames are tiled, any count works)
    assert ops.tafa_keyproj_chunk(1, 49, 64, 4) == 32
    assert ops.tafa_keyproj_chunk(16, 49, 512, 8) == 0           # heads != 4
    assert ops.tafa_keyproj_chunk(16, 49, 48, 4) == 0            # channels not a multiple of the 32-channel chunk
    assert ops.tafa_keyproj_chunk(16, 16 * 16, 512, 4) == 0      # 16x16 bins: the frame tile exceeds shared memory
    assert ops.tafa_keyproj_chunk(0, 49, 512, 4) == 0


# ------------------------------------------------------------------------------------------ round 2
def test_register_into_openmmlab_reports_what_failed(monkeypatch):
    """registration failures must not be swallowed.  Absent packages are reported as such; a registry that raises
    is reported with the exception text (and raised under strict=True)."""
    import types
    import warnings
    touched, failed = vod.register_into_openmmlab()
    assert touched == [] and failed == {'mmtrack.AGGREGATORS': 'not installed', 'mmdet.ROI_EXTRACTORS': 'not installed'}

    class BadRegistry:
        def register_module(self, name=None, force=False, module=None):
            raise TypeError('boom')

    class GoodRegistry:
        def __init__(self):
            self.got = {}

        def register_module(self, name=None, force=False, module=None):
            assert force
            self.got[name] = module

    good = GoodRegistry()
    for name in ('mmtrack', 'mmtrack.models'):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, 'mmtrack.models.builder', types.SimpleNamespace(AGGREGATORS=BadRegistry()))
    for name in ('mmdet', 'mmdet.models'):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, 'mmdet.models.builder', types.SimpleNamespace(ROI_EXTRACTORS=good))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        touched, failed = vod.register_into_openmmlab()
    assert touched == ['mmdet.ROI_EXTRACTORS'] and set(good.got) == {'SingleRoIExtractor', 'TemporalRoIAlign'}
    assert failed == {'mmtrack.AGGREGATORS': 'TypeError: boom'} and any('boom' in str(x.message) for x in w)
    with pytest.raises(TypeError):
        vod.register_into_openmmlab(strict=True)


def test_drop_ins_refuse_to_cut_a_training_graph():
    """The kernels have no backward: a forward the reference would differentiate through (training mode, autograd on, an input
    that requires grad) raises instead of silently returning a graph-less tensor; eval / no_grad / plain inputs go through
    (here: down to the 'no CPU fallback' error)."""
    agg = vod.SelsaAggregator(16, 4)
    x = torch.randn(2, 16, requires_grad=True)
    with pytest.raises(RuntimeError, match='inference-only'):
        agg(x, torch.randn(4, 16))
    with pytest.raises(_lib.VodError):
        agg.eval()(x, torch.randn(4, 16))
    agg.train()
    with torch.no_grad(), pytest.raises(_lib.VodError):
        agg(x, torch.randn(4, 16))
    t = vod.build_roi_extractor(dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=4, featmap_strides=[16]))
    with pytest.raises(RuntimeError, match='inference-only'):
        t((torch.randn(1, 4, 8, 8, requires_grad=True),), torch.tensor([[0, 0., 0., 32., 32.]]))


def test_selsa_aggregator_without_reference_proposals():
    """ref_x with zero rows: the reference's softmax over an empty axis and bmm with an empty operand give zeros, i.e. fc(0)
    (selsa_aggregator.py:61-72); the drop-in returns the same instead of failing on a null pointer."""
    agg = vod.SelsaAggregator(16, 4)
    out = agg(torch.randn(3, 16), torch.zeros(0, 16))
    assert out.shape == (3, 16) and torch.allclose(out, agg.fc.bias.detach().expand(3, 16))


def test_production_library_ignores_probe_environment():
    """the probe hooks of the key-projected logits kernel (VOD_KP_DBG ...) exist only in a -DVOD_PROBES build."""
    src = open(os.path.join(ROOT, 'lowlightenvironmentvideoobjectdetection_b200', 'csrc', 'tafa_keyproj.cu')).read()
    body = re.sub(r'#ifdef VOD_PROBES.*?#endif', '', src, flags=re.S)
    assert 'getenv' not in body
    from lowlightenvironmentvideoobjectdetection_b200 import build
    assert not any('VOD_PROBES' in f for f in build.NVCC_FLAGS)
    blob = open(_lib.LIB_PATH, 'rb').read()
    assert b'VOD_KP_DBG' not in blob


def test_ref_frame_cache_bookkeeping_cpu():
    """Host logic of the reference-frame cache (heads.RefFrameCache): a FIFO advance is found as the shift that preserves the
    most frames, the buffers move with it (in reference order, so reductions keep their order), and only frames the cache does
    not hold are left to recompute.  No kernel runs here."""
    from lowlightenvironmentvideoobjectdetection_b200.heads import RefFrameCache, _frame_key, _slot_runs
    head = vod.SelsaRoIHead(
        bbox_roi_extractor=dict(type='TemporalRoIAlign', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                out_channels=64, featmap_strides=[16]),
        bbox_head=dict(type='SelsaBBoxHead', num_shared_fcs=2, in_channels=64, fc_out_channels=128, num_classes=3,
                       aggregator=dict(type='SelsaAggregator', in_channels=128, num_attention_blocks=2)))
    T, N = 5, 4
    c = head.new_ref_cache(T, N, (64, 3, 4), 'cpu')
    assert c.maps.shape == (T, 3, 4, 64) and c.unit.shape == (T * 12, 64) and c.v_transposed and c.V[0].shape == (128, T * N)
    assert not c.filled()
    for t in range(T):                       # recognisable contents: slot t holds the value t
        c.maps[t] = t; c.norm[t * 12:(t + 1) * 12] = t; c.unit[t * 12:(t + 1) * 12] = t
        for i in range(2):
            c.K[i][t * N:(t + 1) * N] = t; c.V[i][:, t * N:(t + 1) * N] = t
    keys = [_frame_key(dict(video_id=1, frame_id=f, img_shape=(48, 64, 3))) for f in range(10)]
    c.mark(list(range(T)), keys[0:T])
    assert c.filled()
    c.align(keys[0:T])                       # same reference set: nothing to do
    assert c.filled() and c.keys == keys[0:T]
    c.align(keys[2:T + 2])                   # the set advanced by two frames
    assert c.keys[:3] == keys[2:5] and c._filled == [True, True, True, False, False]
    for t in range(3):
        assert float(c.maps[t].min()) == float(c.maps[t].max()) == t + 2
        assert float(c.norm[t * 12]) == t + 2 and float(c.unit[t * 12, 0]) == t + 2
        assert float(c.K[1][t * N, 0]) == t + 2 and float(c.V[0][5, t * N + 1]) == t + 2
    c.mark([3, 4], keys[5:7])
    want = list(keys[2:T + 2]); want[1] = keys[9]        # one slot replaced in place (the key frame's slot)
    c.align(want)
    assert c._filled == [True, False, True, True, True]
    c.align(keys[7:10] + keys[0:2])          # an unrelated set: everything is recomputed
    assert not any(c._filled)
    assert _slot_runs([3, 4, 0, 1, 2, 9]) == [(0, 3, 2), (2, 0, 3), (5, 9, 1)]
    # a cache is tied to the weights it was computed with
    assert c.compatible(head, T, N, (64, 3, 4), 'cpu')
    with torch.no_grad():
        head.bbox_head.shared_fcs[0].weight.mul_(2.0)
    assert not c.compatible(head, T, N, (64, 3, 4), 'cpu')


def test_selsa_fused_tail_algebra_cpu():
    """The identities behind vod_selsa_residual_relu (host side, CPU): projecting V^T without ref_fc's bias and adding
    ``SelsaAggregator.out_bias`` = fc.bias + fc.weight . ref_fc.bias after the last linear gives the reference's
    fc(softmax(.) (V + b_v)) (selsa_aggregator.py:64-72: softmax rows sum to one), and the layer tail
    relu(x + y + bias) / relu(ref_x) is selsa_bbox_head.py:56-58.  The bias is cached per weight version."""
    from oracle import vod_oracle as O
    torch.manual_seed(3)
    m = vod.SelsaAggregator(in_channels=128, num_attention_blocks=2)       # d = 64: the V^T (bias-free) projection layout
    for prm in m.parameters():
        torch.nn.init.normal_(prm, 0, 0.05)
    x, ref_x = torch.randn(7, 128), torch.randn(19, 128)
    p = {k: v.detach() for k, v in m.state_dict().items()}
    want = O.selsa_aggregate(x, ref_x, p, 2)
    with torch.no_grad():
        q, k = m.fc_embed(x), m.ref_fc_embed(ref_x)
        v0 = ref_x @ m.ref_fc.weight.t()                                    # V without its bias
        w = torch.softmax((q.view(7, 2, 64).transpose(0, 1) @ k.view(19, 2, 64).permute(1, 2, 0)) / 8.0, dim=-1)
        o = (w @ v0.view(19, 2, 64).transpose(0, 1)).transpose(0, 1).reshape(7, 128)
        y = torch.nn.functional.linear(o, m.fc.weight)                      # what attend(..., with_bias=False) returns
        b = m.out_bias(True)
        assert torch.allclose(y + b, want, atol=1e-5)
        assert m.out_bias(True) is b                                        # cached
        assert m.out_bias(False) is m.fc.bias                               # V projected with its bias: nothing to fold
        m.ref_fc.bias.add_(1.0)                                             # a weight change invalidates the cached bias
        assert not torch.equal(m.out_bias(True), b)
