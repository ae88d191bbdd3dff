import os, sys, torch
sys.path.insert(0, '.')
from lowlightenvironmentvideoobjectdetection_b200 import ops
N, T, C, HW = 300, 15, 512, 38 * 63
g = torch.Generator(device='cuda').manual_seed(0)
ru = torch.nn.functional.normalize(torch.rand(N * 49, C, device='cuda', generator=g), dim=1).bfloat16()
unit = torch.nn.functional.normalize(torch.rand(T * HW, C, device='cuda', generator=g), dim=1).bfloat16()
def t(probe):
    os.environ['VOD_MG_PROBE'] = str(probe & ~8)
    for _ in range(3): ops.msra_gemm_candidates(ru, unit, T)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ops.msra_gemm_candidates(ru, unit, T)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print('probe', probe, '%.1f us  %.0f TFLOP/s  (%.0f clk @1.965GHz)' % (us, 2.0 * N * 49 * T * HW * C / us / 1e6, us * 1965), flush=True)
    if probe & 8:
        os.environ['VOD_MG_PROBE'] = str(probe)
        ops.msra_gemm_candidates(ru, unit, T); torch.cuda.synchronize()
for pr in sys.argv[1:] or ['0', '1']:
    t(int(pr))
