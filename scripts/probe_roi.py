"""Probe: RoIAlign over the 4500 reference RoIs of cfg 3 (channels-last output) with an experiment build of the library.
    python scripts/probe_roi.py <path to .so> [more .so ...]
Prints the time per launch (CUDA events, L2 flushed between launches) and the largest difference from the FIRST library's output."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lowlightenvironmentvideoobjectdetection_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import bench  # noqa: E402
from lowlightenvironmentvideoobjectdetection_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
cfg = bench.CONFIGS['cfg3']
ref_x, props = bench.make_inputs(cfg, 0)
T, N = cfg['T'], cfg['N']
nhwc = ref_x.to(dev).permute(0, 2, 3, 1).contiguous()
rois = torch.zeros(T * N, 5, device=dev)
rois[:, 0] = torch.arange(T, device=dev, dtype=torch.float32).repeat_interleave(N)
rois[:, 1:] = props[:T].reshape(T * N, 4).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
golden = {}
for path in sys.argv[1:]:
    _lib.LIB_PATH, _lib._lib = os.path.abspath(path), None      # each library in turn (ctypes keeps them all mapped)
    for nhwc_out in (True, False):
        out = ops.roi_align_nhwc(nhwc, rois, 7, 1 / 16., 2, True, out_nhwc=nhwc_out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.roi_align_nhwc(nhwc, rois, 7, 1 / 16., 2, True, out_nhwc=nhwc_out, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        msg = '%-28s %s  median %.1f us  min %.1f us' % (os.path.basename(path), 'NHWC-out' if nhwc_out else 'NCHW-out', ts[len(ts) // 2], ts[0])
        if nhwc_out in golden:
            msg += '  max|d| vs first %.3g' % (out - golden[nhwc_out]).abs().max().item()
        else:
            golden[nhwc_out] = out.clone()
        print(msg, flush=True)
