"""Times flow_warp_feats at the FGFA shape (31 maps of [512,38,63], flows at 608x1008), L2 flushed between launches."""
import sys, torch
sys.path.insert(0, '.')
from lowlightenvironmentvideoobjectdetection_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
T, C, H, W = 31, 512, 38, 63
x = torch.randn(T, C, H, W, device='cuda', generator=g)
flow = torch.randn(T, 2, H * 16, W * 16, device='cuda', generator=g) * 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): ops.flow_warp(x, flow)
tot = 0.0
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.flow_warp(x, flow); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
t = tot / 10 * 1e3
b = 2 * T * C * H * W * 4 + T * 2 * H * W * 4 * 4
print('flow_warp T=31: %.1f us, %.0f GB/s algorithmic' % (t, b / t / 1e3))
