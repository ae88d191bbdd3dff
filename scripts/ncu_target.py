"""Short ncu target: one replay of the uncached cfg-3 graph step, one cache fill and one step through the reference-frame cache
(after capture and one warm replay each).  NVTX-free: the three phases are separated by a marker kernel count printed to stdout.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/ncu_target.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

args = types.SimpleNamespace(steps=2, warmup=1, no_overlap='--no-overlap' in sys.argv, clip_len=40)
ctx = bench.Ctx(args)
cfg = bench.CONFIGS['cfg3']
run = bench.SelsaRunner(ctx, cfg, n_sets=2, pinned=False)
sink = bench.DetectionSink(ctx, 2)
lib = run.lib
with torch.no_grad():
    run.capture('tf32', tf32=True)
    run.capture_cached(tf32=True)
    torch.cuda.synchronize()
    marker = torch.zeros(1, device=ctx.device)
    for phase, fn in (('uncached_step', lambda: run.step('tf32', 0, run.dev_sets[1], sink)),
                      ('cache_fill', lambda: (run.load_memo(*run.dev_sets[1]), run.graphs['fill'][0].replay())),
                      ('cached_step', lambda: (run.load_key(*run.dev_sets[1]), run.graphs['cached'][0].replay()))):
        marker.add_(1.0)                      # a recognisable 1-element kernel between the phases
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        print('phase', phase, 'done')
