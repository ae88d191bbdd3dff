"""Per-kernel time of one Denoising2Aggergator step (bench.py --config denoise shapes) with the torch profiler (CUPTI).
    python scripts/denoise_breakdown.py"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lowlightenvironmentvideoobjectdetection_b200 as vod  # noqa: E402

dev = torch.device('cuda', 0)
T = bench.CONFIGS['denoise']['T']
torch.manual_seed(0)
agg = vod.build_aggregator(dict(type='Denoising2Aggergator', **bench.DENOISE_SPEC)).eval().to(dev)
st = bench.denoise_inputs(0, T, dev)
with torch.no_grad(), bench.library_math(True):
    for _ in range(2):
        agg(st[:4], st[4:])
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU], record_shapes=True) as prof:
        agg(st[:4], st[4:])
        torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name[:90]][0] += e.device_time
        tot[e.name[:90]][1] += 1
total = sum(v[0] for v in tot.values())
print('kernel time of one step: %.2f ms' % (total / 1e3))
for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:18]:
    print('  %8.2f ms %5.1f%% x%-4d %s' % (t / 1e3, 100 * t / total, n, name))

print('by operator (device time of the kernels each aten op launched):')
rows = [(e.key, e.self_device_time_total, e.count, str(e.input_shapes)[:90]) for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0]
for key, t, n, shp in sorted(rows, key=lambda r: -r[1])[:22]:
    print('  %8.2f ms x%-4d %-32s %s' % (t / 1e3, n, key[:32], shp))
