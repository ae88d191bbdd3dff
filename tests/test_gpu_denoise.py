"""SURVEY 8f row N4 (last item): the fork's Denoising2Aggergator -- the two hand-written kernels (modulated deformable im2col,
temporal softmax fusion) against the oracle, and the module mirrors against the reference's own file (oracle/_ref, loaded
unmodified under the mmcv stand-ins) with the reference's state_dict loaded strictly."""
import pytest
import torch

import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops
from oracle import ref_shim, vod_oracle as O

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _dcn_inputs(g, B, C, H, W, G, stride=1, k=3, scale=2.5):
    Ho, Wo = (H + 2 - (k - 1) - 1) // stride + 1, (W + 2 - (k - 1) - 1) // stride + 1
    x = torch.randn(B, C, H, W, generator=g)
    logits = torch.randn(B, 3 * G * k * k, Ho, Wo, generator=g)
    logits[:, :2 * G * k * k] *= scale                      # offsets of a few pixels: many samples leave the map
    return x, logits


@pytest.mark.parametrize('B,C,H,W,G,Cout,stride', [(3, 64, 19, 23, 8, 64, 1), (2, 32, 12, 17, 2, 24, 2), (1, 512, 9, 11, 8, 128, 1), (2, 12, 8, 9, 4, 7, 1),
                                                   (2, 128, 21, 16, 8, 32, 1), (1, 256, 10, 9, 8, 16, 1), (1, 64, 16, 24, 8, 8, 2)])
def test_mdcn_im2col_gemm_vs_oracle(B, C, H, W, G, Cout, stride):
    """vod_mdcn_im2col + one GEMM == mmcv's modulated_deform_conv2d as restated in the oracle (pinned to torchvision there):
    raw conv_offset logits in (chunk / cat / sigmoid in the kernel), samples outside the map, stride 2.  64 / 128 / 256 / 512
    channels with 8 groups at stride 1 take the shared-memory tile kernel (offsets of +-2.5 sigma pixels: many corners fall
    outside the staged halo and come from global memory), the others the flat kernel (128-bit or scalar)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(B * 100 + C)
    x, logits = _dcn_inputs(g, B, C, H, W, G, stride)
    w = torch.randn(Cout, C, 3, 3, generator=g) * 0.1
    bias = torch.randn(Cout, generator=g)
    K = 9
    want = O.modulated_deform_conv2d(x, logits[:, :2 * G * K], torch.sigmoid(logits[:, 2 * G * K:]), w, bias, stride, 1, 1, 1, G)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    p = logits.permute(0, 2, 3, 1).contiguous().to(DEV)
    col = ops.mdcn_im2col(x_nhwc, p, None, G, 3, stride, 1, 1)
    got = torch.addmm(bias.to(DEV), col, w.to(DEV).permute(0, 2, 3, 1).reshape(Cout, -1).t())
    got = got.view(B, want.shape[2], want.shape[3], Cout).permute(0, 3, 1, 2)
    assert rel_err(got, want) < 1e-5
    # the logits as a per-image map plus one shared map (how TemporalAttentionFusion passes P[t] + Q[i])
    q = torch.randn(1, *p.shape[1:], generator=g).to(DEV)
    col2 = ops.mdcn_im2col(x_nhwc, p - q, q, G, 3, stride, 1, 1)
    assert rel_err(col2, col) < 1e-5


def test_mdcn_bad_arguments_and_empty_batch():
    x = torch.zeros(1, 4, 4, 12, device=DEV)
    with pytest.raises(vod.VodError):
        ops.mdcn_im2col(x, torch.zeros(1, 4, 4, 3 * 5 * 9, device=DEV), None, 5, 3, 1, 1, 1)      # 12 channels, 5 groups
    assert ops.mdcn_im2col(x[:0], torch.zeros(0, 4, 4, 27, device=DEV), None, 1).shape == (0, 108)


@pytest.mark.parametrize('I,T,shape', [(5, 5, (7, 9, 16)), (1, 9, (38, 63, 64)), (3, 1, (4, 4, 8))])
def test_temporal_softmax_fuse_vs_torch(I, T, shape):
    g = torch.Generator().manual_seed(I * 10 + T)
    cor = torch.randn(I, T, *shape, generator=g) * 4
    cor[0, 0] += 60.0                                           # one frame dominates: exp() of the others underflows
    x = torch.randn(T, *shape, generator=g)
    want = (torch.softmax(cor.double(), 1) * x.double()[None]).sum(1).float()
    got = ops.temporal_softmax_fuse(cor.to(DEV), x.to(DEV))
    assert got.shape == want.shape and rel_err(got, want) < 2e-6


def test_modulated_dcn_pack_forward_vs_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    pack = vod.ModulatedDCNPack(16, 24, 3, padding=1, deform_groups=4)
    assert pack.conv_offset.weight.abs().max() == 0 and pack.conv_offset.bias.abs().max() == 0     # init_offset (:70-71)
    torch.nn.init.normal_(pack.conv_offset.weight, 0, 0.1)
    torch.nn.init.normal_(pack.conv_offset.bias, 0, 1.0)
    x, extra = torch.randn(2, 16, 11, 14, generator=g), torch.randn(2, 16, 11, 14, generator=g)
    o = torch.nn.functional.conv2d(extra, pack.conv_offset.weight, pack.conv_offset.bias, padding=1)
    o1, o2, m = torch.chunk(o, 3, 1)
    want = O.modulated_deform_conv2d(x, torch.cat((o1, o2), 1), torch.sigmoid(m), pack.weight, pack.bias, 1, 1, 1, 1, 4)
    got = pack.to(DEV).eval()(x.to(DEV), extra.to(DEV))
    assert got.shape == want.shape and rel_err(got, want.detach()) < 1e-5


@pytest.mark.parametrize('T', [1, 4])
def test_temporal_attention_fusion_vs_oracle(T):
    """The mirror (4 T linear convs for the offsets instead of 2 T^2, deformable im2col + GEMM, fused softmax-sum) against the
    oracle's restatement of TemporalAttentionFusion.forward, same parameters."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    g = torch.Generator().manual_seed(6)
    taf = vod.TemporalAttentionFusion(24, 16, emb_nums=3)
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.weight, 0, 0.05)
    torch.nn.init.normal_(taf.dcn_pack.conv_offset.bias, 0, 0.7)
    x = torch.randn(T, 24, 13, 17, generator=g)
    want = O.temporal_attention_fusion(x, {k: v.detach() for k, v in taf.state_dict().items()})
    got = taf.to(DEV).eval()(x.to(DEV))
    assert got.shape == want.shape == x.shape
    assert rel_err(got, want) < 2e-5


@pytest.mark.skipif(not ref_shim.available(), reason='reference files not staged (oracle/make_ref.py)')
def test_denoising2_aggregator_vs_reference_file():
    """Whole module: the reference's own denoising2_aggregator.py on the CPU vs the mirror on the GPU, reference state_dict
    loaded with strict=True (same parameter names and shapes), two stages with downsampling, RDBs and fusion."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    R = ref_shim.load()
    assert R.Denoising2Aggergator is not None
    cfg = dict(in_channel=[16, 24], mid_channel=[8, 16], out_channel=[24, 12], layer_name=['layer1', 'layer2'], rdb_blocks=[1, 2],
               rdb_channel_growth=[8, 8], taf_embs=[2, 3], downsample=[True, False], with_rdb=[True, True], with_taf=[True, True])
    torch.manual_seed(2)
    ref = R.Denoising2Aggergator(**cfg).eval()
    for name, mod in ref.named_modules():
        if name.endswith('conv_offset'):
            torch.nn.init.normal_(mod.weight, 0, 0.05)
            torch.nn.init.normal_(mod.bias, 0, 0.5)
    ours = vod.build_aggregator(dict(type='Denoising2Aggergator', **cfg))
    ours.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(7)
    T = 3
    x_noise = [torch.randn(T, 16, 16, 20, generator=g), torch.randn(T, 24, 8, 10, generator=g)]
    all_x = [torch.randn(T, 12, 8, 10, generator=g)]
    with torch.no_grad():
        want_noise, want_all = ref([t.clone() for t in x_noise], [t.clone() for t in all_x])
    got_noise, got_all = ours.to(DEV).eval()([t.to(DEV) for t in x_noise], [t.to(DEV) for t in all_x])
    assert isinstance(got_noise, tuple) and isinstance(got_all, tuple) and len(got_noise) == 2 and len(got_all) == 1
    for a, b in zip(got_noise + got_all, want_noise + want_all):
        assert a.shape == b.shape and rel_err(a, b) < 5e-5
    # without fusion / dense blocks the module is plain convolutions: still the same outputs
    cfg2 = dict(cfg, with_rdb=[False, True], with_taf=[True, False])
    ref2 = R.Denoising2Aggergator(**cfg2).eval()
    ours2 = vod.Denoising2Aggergator(**cfg2)
    ours2.load_state_dict(ref2.state_dict(), strict=True)
    with torch.no_grad():
        w2 = ref2([t.clone() for t in x_noise], [t.clone() for t in all_x])
    g2 = ours2.to(DEV).eval()([t.to(DEV) for t in x_noise], [t.to(DEV) for t in all_x])
    for a, b in zip(g2[0] + g2[1], w2[0] + w2[1]):
        assert rel_err(a, b) < 5e-5


def test_denoising2_aggregator_vs_golden():
    """The same check against the COMMITTED fixture (tests/golden/denoise_golden.npz, generated from the reference's own file
    by tests/golden/make_denoise_golden.py): needs neither /root/reference nor the staged copy."""
    import os
    import numpy as np
    from tests_golden_cfg import AGG_CFG
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'denoise_golden.npz'))
    gd = {k: torch.from_numpy(z[k]) for k in z.files}
    taf = vod.TemporalAttentionFusion(16, 8, emb_nums=3)
    taf.load_state_dict({k[len('taf_p.'):]: v for k, v in gd.items() if k.startswith('taf_p.')}, strict=True)
    got = taf.to(DEV).eval()(gd['taf_x'].to(DEV))
    assert rel_err(got, gd['taf_out']) < 2e-5
    agg = vod.Denoising2Aggergator(**AGG_CFG)
    agg.load_state_dict({k[len('agg_p.'):]: v for k, v in gd.items() if k.startswith('agg_p.')}, strict=True)
    noise_out, all_out = agg.to(DEV).eval()([gd['agg_x_noise.0'].to(DEV), gd['agg_x_noise.1'].to(DEV)], [gd['agg_all_x.0'].to(DEV)])
    for i, t in enumerate(noise_out):
        assert rel_err(t, gd['agg_noise_out.%d' % i]) < 5e-5
    assert rel_err(all_out[0], gd['agg_all_out.0']) < 5e-5


def test_denoising_modules_are_inference_only():
    taf = vod.TemporalAttentionFusion(8, 8, emb_nums=1).to(DEV).train()
    x = torch.randn(2, 8, 6, 6, device=DEV, requires_grad=True)
    with pytest.raises(RuntimeError, match='inference-only'):
        taf(x)
