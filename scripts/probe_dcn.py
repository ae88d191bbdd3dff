import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowlightenvironmentvideoobjectdetection_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from lowlightenvironmentvideoobjectdetection_b200 import ops
dev = 'cuda'
for (T, h, w, mid) in ((9, 152, 252, 64), (9, 76, 126, 128), (9, 38, 63, 256), (9, 38, 63, 512)):
    y = torch.randn(T, h, w, mid, device=dev)
    pq = torch.randn(T + 1, h, w, 216, device=dev) * 0.7
    col = torch.empty(T * h * w, 9 * mid, device=dev)
    f = lambda: ops.mdcn_im2col(y, pq[:T], pq[T:], 8, 3, 1, 1, 1, out=col)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print('  mdcn_im2col T=%d %dx%d C=%d: %.1f us' % (T, h, w, mid, e0.elapsed_time(e1) * 100))
