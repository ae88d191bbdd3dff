"""Times TemporalRoIAlign's attention stage both ways at cfg-3 size (N=300, 16 stacked frames, C=512):
full-embedding path (cuDNN conv over 4800 patches + vod_tafa_weighted_sum) against the key-projected path
(conv on the 300 key patches, G GEMM, vod_tafa_keyproj_logits, vod_tafa_weighted_sum_logits), piece by piece.
Library math in tf32, as in bench.py."""
import sys, torch
sys.path.insert(0, '.')
import lowlightenvironmentvideoobjectdetection_b200 as vod
from lowlightenvironmentvideoobjectdetection_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
T1 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
C, P, H = 512, 49, 4
torch.manual_seed(0)
m = vod.build_roi_extractor(dict(type='TemporalRoIAlign', num_most_similar_points=2, num_temporal_attention_blocks=4,
                                 roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                 out_channels=512, featmap_strides=[16])).cuda()
m.roi_layers[0].channels_last_out = True
g = torch.Generator(device='cuda').manual_seed(0)
x_all = torch.relu(torch.randn(T1, N, P, C, device='cuda', generator=g))
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
conv = m.embed_network.conv
m.keyproj = False
t_old = timeit(lambda: m._tafa(x_all, 7, 7))
a = m._tafa(x_all, 7, 7)
m.keyproj = True
t_new = timeit(lambda: m._tafa(x_all, 7, 7))
b = m._tafa(x_all, 7, 7)
print('T1 %d N %d: full-embedding path %.1f us, key-projected path %.1f us, rel diff (tf32 library math) %.2e'
      % (T1, N, t_old, t_new, float((a - b).abs().max() / a.abs().max())))
cc = ops.tafa_keyproj_chunk(T1, P, C, H)
wcl = m._conv_weight_cl(conv); wr = m._keyproj_weight(conv, H, cc)
kp = x_all[0].view(N, 7, 7, C).permute(0, 3, 1, 2)
t_conv = timeit(lambda: torch.nn.functional.conv2d(kp, wcl, conv.bias, 1, 1))
ek = torch.nn.functional.conv2d(kp, wcl, conv.bias, 1, 1).permute(0, 2, 3, 1).contiguous().view(N * P, H, C // H)
t_bmm = timeit(lambda: torch.bmm(ek.transpose(0, 1), wr))
G = torch.bmm(ek.transpose(0, 1), wr)
t_log = timeit(lambda: ops.tafa_keyproj_logits(x_all, G, 7, H, cc))
parts = ops.tafa_keyproj_logits(x_all, G, 7, H, cc)
t_app = timeit(lambda: ops.tafa_weighted_sum_logits(x_all, parts, H, out_nhwc=True))
by = (G.numel() + x_all.numel() + parts.numel()) * 4
print('  key conv %.1f us | G GEMM %.1f us (%.0f TFLOP/s, %.0f GB/s written) | logits %.1f us (%.0f GB/s, %.1f TFMA/s) | apply %.1f us (%.0f GB/s)'
      % (t_conv, t_bmm, 2.0 * N * P * C * 9 * C / t_bmm / 1e6, G.numel() * 4 / t_bmm / 1e3, t_log, by / t_log / 1e3,
         T1 * N * 361 * H * C / t_log / 1e6, t_app, (x_all.numel() + x_all.numel() // T1) * 4 / t_app / 1e3))
