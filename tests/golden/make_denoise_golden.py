"""Generates tests/golden/denoise_golden.npz by running the reference's own UNMODIFIED
mmtrack/models/aggregators/denoising2_aggregator.py (loaded under the stand-ins of oracle/ref_shim.py: mmcv's
modulated_deform_conv2d = torchvision.ops.deform_conv2d) on fixed-seed inputs.  Run in the build container only:

    python tests/golden/make_denoise_golden.py

Holds inputs, weights and the reference's outputs of (i) one TemporalAttentionFusion call and (ii) one two-stage
Denoising2Aggergator call, so that the parity tests need neither /root/reference nor oracle/_ref.
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

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tests_golden_cfg import AGG_CFG  # noqa: E402


def trained_like_offsets(module):
    """conv_offset is zero-initialised (init_offset): give it values so that samples fall off-grid and partly outside the map."""
    for name, mod in module.named_modules():
        if name.endswith('conv_offset'):
            torch.nn.init.normal_(mod.weight, 0, 0.05)
            torch.nn.init.normal_(mod.bias, 0, 0.6)


def main():
    R = ref_shim.load()
    out = {}
    g = torch.Generator().manual_seed(20261019)
    with torch.no_grad():
        torch.manual_seed(11)
        taf = R.TemporalAttentionFusion(16, 8, emb_nums=3).eval()
        trained_like_offsets(taf)
        x = torch.randn(4, 16, 10, 13, generator=g)
        out['taf_x'] = x
        for k, v in taf.state_dict().items():
            out['taf_p.' + k] = v
        out['taf_out'] = taf(x.clone())

        torch.manual_seed(12)
        agg = R.Denoising2Aggergator(**AGG_CFG).eval()
        trained_like_offsets(agg)
        x_noise = [torch.randn(3, 16, 16, 20, generator=g), torch.randn(3, 24, 8, 10, generator=g)]
        all_x = [torch.randn(3, 12, 8, 10, generator=g)]
        for i, t in enumerate(x_noise):
            out['agg_x_noise.%d' % i] = t
        out['agg_all_x.0'] = all_x[0]
        for k, v in agg.state_dict().items():
            out['agg_p.' + k] = v
        noise_out, all_out = agg([t.clone() for t in x_noise], [t.clone() for t in all_x])
        for i, t in enumerate(noise_out):
            out['agg_noise_out.%d' % i] = t
        out['agg_all_out.0'] = all_out[0]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'denoise_golden.npz')
    np.savez_compressed(path, **{k: v.detach().cpu().numpy() for k, v in out.items()})
    print('wrote', path, '%d arrays, %d bytes' % (len(out), os.path.getsize(path)))


if __name__ == '__main__':
    main()
