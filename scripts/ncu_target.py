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
EAGER = '--eager' in sys.argv      # ncu does not list the kernels of a replayed torch CUDA graph: run the same three functions eagerly
with torch.no_grad(), bench.library_math(True):
    if EAGER:                      # (and no capture at all in the process: ncu stops listing launches once a stream capture has begun)
        run.setup_cached()
        run.load_inputs(*run.dev_sets[0]); run.load_memo(*run.dev_sets[0]); run.load_key(*run.dev_sets[0])
    else:
        run.capture('tf32', tf32=True)
        run.capture_cached(tf32=True)
    torch.cuda.synchronize()
    marker = torch.zeros(1, device=ctx.device)
    T, N = cfg['T'], cfg['N']
    head, cache = run.head, run.cache
    if EAGER:
        phases = (('uncached_step', lambda: (run.load_inputs(*run.dev_sets[1]), head.simple_test_device(
                      (run.st_ref[T - 1:T],), (run.st_ref,), run.st_rois, run.st_ref_rois, bench.IMG_SHAPE, (1., 1., 1., 1.)))),
                  ('cache_fill', lambda: (run.load_memo(*run.dev_sets[1]),
                                          head.update_ref_cache(cache, list(range(T - 1)), run.st_memo, run.st_memo_rois))),
                  ('cached_step', lambda: (run.load_key(*run.dev_sets[1]), head.simple_test_cached_device(
                      (run.st_key,), run.st_key_rois, run.st_key_ref_rois, cache, T - 1, bench.IMG_SHAPE, (1., 1., 1., 1.)))))
    else:
        phases = (('uncached_step', lambda: run.step('tf32', 0, run.dev_sets[1], sink)),
                  ('cache_fill', lambda: (run.load_memo(*run.dev_sets[1]), run.graphs['fill'][0].replay())),
                  ('cached_step', lambda: (run.load_key(*run.dev_sets[1]), run.graphs['cached'][0].replay())))
    if EAGER:
        for _, fn in phases:      # warm-up: workspaces, module loading, cuBLAS / cuDNN plans
            fn(); fn()
    for phase, fn in phases:
        marker.add_(1.0)                      # a recognisable 1-element kernel between the phases
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        print('phase', phase, 'done')
