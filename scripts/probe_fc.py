"""Times the shared fc_0 of SelsaBBoxHead (K = 25088 -> 1024, tf32 library math) for the key rows (300), the reference
rows (4500) and both together, and a K-split of the small product (more CTAs than a 300 x 1024 output has tiles)."""
import sys, torch
sys.path.insert(0, '.')
torch.backends.cuda.matmul.allow_tf32 = True
g = torch.Generator(device='cuda').manual_seed(0)
K, D = 25088, 1024
w = torch.randn(D, K, device='cuda', generator=g) * 0.01
b = torch.zeros(D, device='cuda')
xk = torch.randn(300, K, device='cuda', generator=g)
xr = torch.randn(4500, K, device='cuda', generator=g)
xa = torch.cat([xk, xr])
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
F = torch.nn.functional
print('fc_0 key rows (300)   %.1f us' % timeit(lambda: F.linear(xk, w, b)))
print('fc_0 ref rows (4500)  %.1f us' % timeit(lambda: F.linear(xr, w, b)))
print('fc_0 all rows (4800)  %.1f us' % timeit(lambda: F.linear(xa, w, b)))
for S in (4, 8, 16):
    wk = w.view(D, S, K // S).permute(1, 2, 0).contiguous()          # [S, K/S, D]
    def split():
        p = torch.bmm(xk.view(300, S, K // S).transpose(0, 1), wk)   # [S, 300, D]
        return p.sum(0) + b
    print('fc_0 key rows, K split %2d ways: %.1f us  (max diff %.2e)' % (S, timeit(split), float((split() - F.linear(xk, w, b)).abs().max())))
for M in (4500, 300):
    x = torch.randn(M, D, device='cuda', generator=g); w1 = torch.randn(D, D, device='cuda', generator=g)
    print('1024x1024 linear, %d rows: %.1f us' % (M, timeit(lambda: F.linear(x, w1, b))))
