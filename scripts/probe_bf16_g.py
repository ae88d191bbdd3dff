"""Probe: how much precision does a bf16 G operand (key embedding x conv weight, TAFA key-projected logits) cost?
Emulated in torch (G rounded to bf16, then the shipped fp32 kernel): bbox_feats against the fp32-library-math result."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = types.SimpleNamespace(steps=2, warmup=1, no_overlap=True, clip_len=40)
ctx = bench.Ctx(args)
cfg = bench.CONFIGS['cfg3']
run = bench.SelsaRunner(ctx, cfg, n_sets=2, pinned=False)
ext = run.head.bbox_roi_extractor
ref_x, props = run.dev_sets[0]
rois, _ = bench.step_rois(cfg, props)
T = cfg['T']


def feats():
    return ext((ref_x[T - 1:T],), rois, ref_feats=(ref_x,)).float().clone()


def rel(a, b):
    per = (a - b).abs().flatten(1).amax(1) / b.abs().max()
    return float(per.max()), float(per.median())


with torch.no_grad():
    with bench.library_math(False):
        base = feats()
    with bench.library_math(True):
        tf32 = feats()
    orig = ext._keyproj_prepare

    def variant(mode):
        def prep(x_key, N, rh, rw, C, T1):
            heads = ext.num_temporal_attention_blocks
            conv = ext.embed_network.conv
            P = rh * rw
            cc = ext._keyproj_chunk(T1, P, C)
            key_patches = x_key.view(N, rh, rw, C).permute(0, 3, 1, 2)
            ek = torch.nn.functional.conv2d(key_patches, ext._conv_weight_cl(conv), conv.bias, 1, 1)
            ek = ek.permute(0, 2, 3, 1).contiguous().view(N * P, heads, C // heads)
            W = ext._keyproj_weight(conv, heads, cc)
            if mode == 'bf16_in_out':
                G = torch.bmm(ek.transpose(0, 1).bfloat16(), W.bfloat16()).float()
            else:   # tf32 GEMM, bf16 output rounding only
                G = torch.bmm(ek.transpose(0, 1), W).bfloat16().float()
            return G, cc
        return prep
    for mode in ('bf16_in_out', 'bf16_out'):
        ext._keyproj_prepare = variant(mode)
        with bench.library_math(True):
            v = feats()
        print(mode, 'vs fp32: max %.3e median %.3e' % rel(v, base))
    ext._keyproj_prepare = orig
    print('tf32 (shipped) vs fp32: max %.3e median %.3e' % rel(tf32, base))
