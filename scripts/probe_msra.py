"""Times the msra similarity GEMM alone and the whole msra_topk_sample op (GEMM + 2 norms + re-score) at cfg-3 size."""
import os, sys, torch
sys.path.insert(0, '.')
from lowlightenvironmentvideoobjectdetection_b200 import _lib, ops
N, T, C, H, W = 300, 15, 512, 38, 63
HW = H * W
g = torch.Generator(device='cuda').manual_seed(0)
ref = torch.relu(torch.randn(T, C, H, W, device='cuda', generator=g))
rows = torch.relu(torch.randn(N * 49, C, device='cuda', generator=g))
ref_nhwc, norm, unit = ops.to_nhwc(ref, want_norm=True, want_unit_bf16=True)
ru = torch.nn.functional.normalize(rows, dim=1).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, iters=10):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
out = torch.empty(T, N * 49, C, device='cuda')
for path in (sys.argv[1:] or [_lib.LIB_PATH]):       # optional: experiment builds of the library, each in turn
    _lib.LIB_PATH, _lib._lib = os.path.abspath(path), None
    tg = timeit(lambda: ops.msra_gemm_candidates(ru, unit, T))
    tf = timeit(lambda: ops.msra_topk_sample(rows, ref_nhwc, 2, ref_norm=norm, ref_unit=unit, out=out))
    print('%-28s gemm %.1f us (%.0f TFLOP/s)   msra_topk_sample %.1f us   norms+rescore(+overflow) %.1f us' % (
        os.path.basename(path), tg, 2.0 * N * 49 * T * HW * C / tg / 1e6, tf, tf - tg), flush=True)
