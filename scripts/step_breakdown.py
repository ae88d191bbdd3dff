"""Per-kernel breakdown of the graph-replayed key-frame step (uncached and through the reference-frame cache) from the torch
profiler (CUPTI): kernel name, launches, total microseconds, share of the step.  Writes CSVs next to the printed tables.

    python scripts/step_breakdown.py [out_dir]
"""
import csv
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def profile(fn, name, out_dir):
    from torch.profiler import ProfilerActivity, profile as prof
    fn(); torch.cuda.synchronize()
    with prof(activities=[ProfilerActivity.CUDA]) as p:
        fn(); torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total) for e in p.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    total = sum(r[2] for r in rows)
    print('==== %s: %d kernels, %.1f us of kernel time' % (name, sum(r[1] for r in rows), total))
    for k, n, t in rows[:40]:
        print('  %-90s x%-3d %8.1f us  %5.1f %%' % (k[:90], n, t, 100 * t / total))
    if out_dir:
        with open(os.path.join(out_dir, 'r02_launches_%s.csv' % name), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel', 'launches', 'total_us', 'share'])
            for k, n, t in rows:
                w.writerow([k, n, '%.2f' % t, '%.4f' % (t / total)])


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else None
    args = types.SimpleNamespace(steps=4, warmup=2, no_overlap='--no-overlap' in sys.argv, clip_len=40)
    ctx = bench.Ctx(args)
    cfg = bench.CONFIGS['cfg3']
    run = bench.SelsaRunner(ctx, cfg)
    sink = bench.DetectionSink(ctx, 4)
    with torch.no_grad():
        run.capture('tf32', tf32=True)
        run.capture_cached(tf32=True)
        run.load_memo(*run.dev_sets[0]); run.graphs['fill'][0].replay()
        profile(lambda: run.step('tf32', 0, run.dev_sets[1], sink), 'uncached_step', out_dir)
        profile(lambda: (run.load_key(*run.dev_sets[1]), run.graphs['cached'][0].replay()), 'cached_step', out_dir)
        profile(lambda: (run.load_memo(*run.dev_sets[1]), run.graphs['fill'][0].replay()), 'cache_fill', out_dir)


if __name__ == '__main__':
    main()
