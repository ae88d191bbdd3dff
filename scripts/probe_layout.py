import torch, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowlightenvironmentvideoobjectdetection_b200 import _lib
if len(sys.argv) > 1:
    import os
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from lowlightenvironmentvideoobjectdetection_b200 import ops
x = torch.relu(torch.randn(15, 512, 38, 63, device='cuda'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run():
    ops._nhwc_memo.clear()
    return ops._to_nhwc(x, want_norm=True, want_unit_bf16=True)
a = run(); torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort(); print('to_nhwc 15 maps: median %.1f us (min %.1f)' % (ts[len(ts) // 2], ts[0]))
ref = x.permute(0, 2, 3, 1).contiguous()
print('equal', torch.equal(a[0].reshape(-1), ref.reshape(-1)), float((a[1].reshape(-1) - ref.norm(dim=3).reshape(-1)).abs().max()))
