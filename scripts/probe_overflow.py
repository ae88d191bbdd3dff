"""Probe (GPU box): cost of the exact-top-k overflow fix on the benchmark's iid-noise input vs a coherent input.
Prints the flagged-pair count, the per-kernel times of one vod_msra_topk_sample call (torch profiler) and event timings."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lowlightenvironmentvideoobjectdetection_b200 import _lib, ops  # noqa: E402

if '--probes' in sys.argv:      # experiment build (python -m ...build --probes): kernel variants selected by VOD_* variables
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libvodagg_probes%s.so' % os.environ.get('VOD_PROBES_SUFFIX', ''))
    print('using', _lib.LIB_PATH, 'VOD_RF_VARIANT =', os.environ.get('VOD_RF_VARIANT'))


def run(name, ref_x, rois):
    dev = 'cuda'
    T, N = ref_x.shape[0], rois.shape[0]
    ref_nhwc, norm, unit = ops._to_nhwc(ref_x, want_norm=True, want_unit_bf16=True)
    key_rows = ops.roi_align_nhwc(ref_nhwc[T - 1:T].contiguous(), rois, 7, 1 / 16, 2, True, out_nhwc=True).view(N * 49, 512)
    fn = lambda: ops.msra_topk_sample(key_rows, ref_nhwc, 2, ref_norm=norm, ref_unit=unit)
    fn(); torch.cuda.synchronize()
    flagged = ops.msra_overflow_count(N * 49, 512, T, 38 * 63, key_rows.device)
    lib = _lib.load()
    ws = ops._ws.get(lib.vod_msra_workspace_bytes(N * 49, 512, T, 38 * 63, 2), key_rows.device)
    off = int(lib.vod_msra_overflow_counter_offset(N * 49, 512, T, 38 * 63))
    ctrl = ws[off:off + 4 * (1 + 4 * T)].view(torch.int32).cpu()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print('%s: flagged pairs %d of %d (%.2f%%), (pair,group) items %d, bins min/max %d/%d, op time %.0f us' % (
        name, flagged, N * 49 * T, 100.0 * flagged / (N * 49 * T), int(ctrl[1:].sum()), int(ctrl[1:].min()), int(ctrl[1:].max()), min(ts)))
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
            print('    %-60s %8.1f us' % (e.key[:60], e.device_time_total))
    except Exception as ex:
        print('    profiler unavailable:', ex)


def main():
    cfg = bench.CONFIGS['cfg3']
    ref_x, props = bench.make_inputs(cfg, 0)
    rois, _ = bench.step_rois(cfg, props, 'cuda')
    run('bench noise input', ref_x.cuda(), rois)
    # coherent clip: every frame is the key frame plus small noise
    g = torch.Generator().manual_seed(1)
    base = torch.relu(torch.randn(1, 512, 38, 63, generator=g))
    coh = torch.relu(base + 0.1 * torch.randn(15, 512, 38, 63, generator=g))
    run('coherent input', coh.cuda(), rois)


if __name__ == '__main__':
    main()
