"""ncu target: the SELSA attention core at sweep size (1000 x 31000, 16 heads, V^T layout), three launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowlightenvironmentvideoobjectdetection_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device='cuda').manual_seed(0)
N, M = (1000, 31000) if '--cfg3' not in sys.argv else (300, 4500)
q = torch.randn(N, 1024, device=dev, generator=g) * 0.5
k = torch.randn(M, 1024, device=dev, generator=g) * 0.5
vt = torch.randn(1024, (M + 3) // 4 * 4, device=dev, generator=g)
for _ in range(3):
    out = ops.selsa_attention(q, k, vt, 16, v_transposed=True)
torch.cuda.synchronize()
print('ok', float(out.abs().max()))
