"""Probe: the SELSA attention core (tcgen05 kernel + merge) at cfg-3 and sweep size with experiment builds of the library.
    python scripts/probe_selsa.py <lib.so> [more .so ...]      (results compared with the first library's)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lowlightenvironmentvideoobjectdetection_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from lowlightenvironmentvideoobjectdetection_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device='cuda').manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cases = {}
for name, (N, M) in (('cfg3 300x4500', (300, 4500)), ('cfg1 300x900', (300, 900)), ('sweep 1000x31000', (1000, 31000))):
    q = torch.randn(N, 1024, device=dev, generator=g) * 0.5
    k = torch.randn(M, 1024, device=dev, generator=g) * 0.5
    vt = torch.randn(1024, (M + 3) // 4 * 4, device=dev, generator=g)
    cases[name] = (q, k, vt, M)
golden = {}
for path in sys.argv[1:]:
    _lib.LIB_PATH, _lib._lib = os.path.abspath(path), None
    ops._ws.__init__()
    for name, (q, k, vt, M) in cases.items():
        fn = lambda: ops.selsa_attention(q, k, vt, 16, v_transposed=True)
        for _ in range(3):
            out = fn()
        ts = []
        for _ in range(11):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        msg = '%-26s %-18s median %7.1f us  min %7.1f us' % (os.path.basename(path), name, ts[len(ts) // 2], ts[0])
        if name in golden:
            msg += '  max|d| vs first %.2e (max|out| %.2f)' % ((out - golden[name]).abs().max().item(), golden[name].abs().max().item())
        else:
            golden[name] = out.clone()
        print(msg, flush=True)
