"""Where the end-to-end loop's time goes: runs bench.py's uncached e2e loop (prefetch -> static buffers -> graph replay -> D2H)
for a few steps under the torch profiler and prints, per step, the GPU-side intervals on every stream (kernels grouped).
    python scripts/e2e_trace.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = types.SimpleNamespace(steps=6, warmup=1, no_overlap=False, clip_len=40)
ctx = bench.Ctx(args)
cfg = bench.CONFIGS['cfg3']
run = bench.SelsaRunner(ctx, cfg, n_sets=4, pinned=True)
sink = bench.DetectionSink(ctx, 8)
out_host = torch.empty(100, 6).pin_memory()
cnt_host = torch.empty(1, dtype=torch.int32).pin_memory()
with torch.no_grad(), bench.library_math(True):
    run.capture('tf32', tf32=True)
    pf = bench.Prefetcher(run.dev_sets[0])

    def e2e_loop(steps):
        pf.begin()
        pf.prefetch(0, run.host_sets[0])
        for i in range(steps):
            if i + 1 < steps:
                pf.prefetch(i + 1, run.host_sets[(i + 1) % 4])
            run.step('tf32', i, pf.wait(i), sink)
            pf.done(i)
            out_host.copy_(sink.buf[i % sink.frames], non_blocking=True)
            cnt_host.copy_(sink.cnt[i % sink.frames:i % sink.frames + 1], non_blocking=True)
    e2e_loop(3)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        e2e_loop(6)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
# merge consecutive events of the same kind on the same stream into blocks
blocks = []
for e in ev:
    kind = 'memcpy ' + e.name if 'Memcpy' in e.name or 'Memset' in e.name else 'kernels'
    s, t = e.time_range.start - t0, e.time_range.end - t0
    if blocks and blocks[-1][0] == kind and s - blocks[-1][2] < 30:
        blocks[-1][2] = t
        blocks[-1][3] += 1
    else:
        blocks.append([kind, s, t, 1])
for kind, s, t, n in blocks:
    if t - s > 20 or 'memcpy' in kind:
        print('%9.0f us .. %9.0f us  (%7.0f us, %4d events)  %s' % (s, t, t - s, n, kind))
