"""Probe: which aten ops of the eager cached step launch the big elementwise kernels (torch profiler, shapes recorded)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

args = types.SimpleNamespace(steps=4, warmup=2, no_overlap=True, clip_len=40)
ctx = bench.Ctx(args)
cfg = bench.CONFIGS['cfg3']
run = bench.SelsaRunner(ctx, cfg)
T, N = cfg['T'], cfg['N']
with torch.no_grad(), bench.library_math(True):
    head = run.head
    cache = head.new_ref_cache(T, N, (bench.C, bench.H, bench.W), ctx.device)
    ref_x, props = run.dev_sets[0]
    rois, ref_rois = bench.step_rois(cfg, props)
    memo_rois = ref_rois[:(T - 1) * N]
    head.update_ref_cache(cache, list(range(T - 1)), ref_x[:T - 1], memo_rois)
    fn = lambda: head.simple_test_cached_device((ref_x[T - 1:T],), rois, ref_rois[(T - 1) * N:], cache, T - 1, bench.IMG_SHAPE, (1., 1., 1., 1.))
    fn(); fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as p:
        fn(); torch.cuda.synchronize()
    print(p.key_averages(group_by_input_shape=True).table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=50, max_shapes_column_width=90))
