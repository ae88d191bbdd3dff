"""Probe: the per-kernel table of bench.py with an experiment build of the library.
    python scripts/probe_kernels.py <path to .so> [substring filter ...]"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lowlightenvironmentvideoobjectdetection_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import bench  # noqa: E402

filt = sys.argv[2:]
with torch.no_grad(), bench.library_math(True):
    kr = bench.kernel_rooflines(bench.CONFIGS['cfg3'], torch.device('cuda', 0), bench.load_peaks())
print(os.path.basename(sys.argv[1]))
for k, v in kr.items():
    if not filt or any(f in k for f in filt):
        print('  %-32s %8.1f us %9.1f %s frac %.3f' % (k, v['seconds'] * 1e6, v['achieved'], v['unit'], v['frac']))
