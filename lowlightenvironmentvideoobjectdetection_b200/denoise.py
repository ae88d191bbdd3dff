"""The fork's own temporal denoising aggregator (last item): ``Denoising2Aggergator`` with its residual dense
blocks and the deformable temporal-attention fusion, mirroring
mmtracking/mmtrack/models/aggregators/denoising2_aggregator.py (registered name, constructor arguments, ``forward(x_noise,
all_x)`` contract and state_dict keys as there, so the fork's checkpoints load).

Convolutions stay library calls (channels-last).  What changes is the temporal-attention fusion (:135-152 of that file), which the
reference runs as a Python loop over the reference frame i with, per i, T convolutions of cat([x_t, x_i]) -> offset features
-> conv_offset -> chunk / cat / sigmoid -> mmcv's modulated deformable conv -> product -> embed convs -> softmax -> sum:

  * offset_conv and conv_offset are linear and applied back to back, so for the pair (i, t)
        conv_offset(offset_conv(cat[x_t, x_i])) = P[t] + Q[i],
    P = conv_offset_w(offset_conv_w[:, :mid](x)), Q = conv_offset(offset_conv_w[:, mid:](x) + b): 4 T convolutions per call
    instead of 2 T^2, and the T^2 offset / mask tensors (216 channels each) are never materialised;
  * ``vod_mdcn_im2col`` takes P[t] + Q[i] directly, applies the chunk / cat / sigmoid semantics in registers and writes the
    modulated columns channels-last for ONE library GEMM per group of pairs (mmcv's op, absent here, does im2col + GEMM per image);
  * ``vod_temporal_softmax_fuse`` does softmax over frames + weighted sum in one pass;
  * the embed convs (:130-131) are a purely linear chain whose result only enters a softmax over the frames: their biases add
    a map that is the same for every frame t (and every i), which that softmax cancels, so the chain runs without biases
    (3 T^2 full-size bias passes less); everything runs channels-last so the library never re-lays tensors out between convs.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .registry import AGGREGATORS, inference_only


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _conv_relu(conv, x):
    """relu(conv(x) + bias) with the bias and the ReLU in the library convolution's epilogue (cuDNN fused op) on the GPU."""
    if x.is_cuda and x.dtype == torch.float32:
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(conv(x))


def _weights_channels_last(module):
    """Keeps every conv weight of ``module`` in channels-last memory (values, shapes and state_dict unchanged): with
    channels-last activations the library otherwise re-lays the weight out on EVERY convolution call."""
    if getattr(module, '_vod_weights_cl', None) is not None and all(p.data_ptr() == q for p, q in module._vod_weights_cl):
        return
    seen = []
    for prm in module.parameters():
        if prm.dim() == 4:
            prm.data = prm.data.contiguous(memory_format=torch.channels_last)
            seen.append((prm, prm.data_ptr()))
    module._vod_weights_cl = seen


def _nhwc(t):
    """[B,C,H,W] (made channels-last if needed) -> its [B,H,W,C] view."""
    return _cl(t).permute(0, 2, 3, 1)


class DenseLayer(nn.Module):
    """3x3 conv + ReLU whose output is appended to its input (denoising2_aggregator.py:10-35)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)

    def forward(self, x):
        return torch.cat((x, _conv_relu(self.conv, x)), 1)


class RDB(nn.Module):
    """Residual dense block: ``num_layers`` dense layers, 1x1 local feature fusion, local residual (:38-70)."""

    def __init__(self, in_channels, channel_growth, num_layers):
        super().__init__()
        self.layers = nn.Sequential(*[DenseLayer(in_channels + channel_growth * i, channel_growth) for i in range(num_layers)])
        self.lff = nn.Conv2d(in_channels + channel_growth * num_layers, in_channels, kernel_size=1)

    def forward(self, x):
        return x + self.lff(self.layers(x))


class ModulatedDCNPack(nn.Module):
    """Modulated deformable conv whose offsets / masks come from ANOTHER feature map (:73-112).  Parameters as mmcv's
    ModulatedDeformConv2d (``weight`` [out, in, kh, kw] uniform(+-1/sqrt(in*kh*kw)), ``bias`` zero) plus ``conv_offset``
    (zero-initialised: the op starts as a plain convolution with mask 0.5)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deform_groups=1, bias=True):
        super().__init__()
        assert groups == 1, 'ModulatedDCNPack: grouped weights are not used by the reference configs'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = ops._pair(kernel_size)
        self.stride, self.padding, self.dilation = int(stride), int(padding), int(dilation)
        self.groups, self.deform_groups = groups, deform_groups
        kh, kw = self.kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        bound = 1.0 / (in_channels * kh * kw) ** 0.5
        nn.init.uniform_(self.weight, -bound, bound)
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 3 * kh * kw, kernel_size=self.kernel_size,
                                     stride=self.stride, padding=self.padding, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def gemm_weight(self):
        """[Cout, K*Cin] with column k*Cin + c: the layout ``vod_mdcn_im2col`` writes its columns in."""
        return self.weight.permute(0, 2, 3, 1).reshape(self.out_channels, -1)

    def from_logits(self, x_nhwc, p, q=None, out=None):
        """x_nhwc [B,H,W,Cin]; p (+ q) raw conv_offset outputs channels-last [B or 1, Ho, Wo, 3*G*K] -> [B*Ho*Wo, Cout]."""
        col = ops.mdcn_im2col(x_nhwc, p, q, self.deform_groups, self.kernel_size, self.stride, self.padding, self.dilation)
        w = self.gemm_weight().t()
        if self.bias is None:
            return torch.mm(col, w, out=out) if out is not None else torch.mm(col, w)
        return torch.addmm(self.bias, col, w, out=out) if out is not None else torch.addmm(self.bias, col, w)

    @inference_only
    def forward(self, x, extra_feat):
        """As the reference's pack: offsets and masks from ``extra_feat``, applied to ``x`` (both [B,C,H,W]) -> [B,Cout,Ho,Wo]."""
        p = _nhwc(self.conv_offset(extra_feat.float()))
        y = self.from_logits(_nhwc(x.float()).contiguous(), p.contiguous())
        return y.view(x.shape[0], p.shape[1], p.shape[2], self.out_channels).permute(0, 3, 1, 2).contiguous().to(x.dtype)


class TemporalAttentionFusion(nn.Module):
    """Every frame re-aligned to every other by a deformable conv, correlated, and fused by a per-element softmax over
    frames (:115-152)."""

    def __init__(self, channels, mid_channels, emb_nums=3):
        super().__init__()
        self.channels, self.mid_channels, self.emb_nums = channels, mid_channels, emb_nums
        self.conv1 = nn.Conv2d(channels, mid_channels, kernel_size=3, padding=1)
        self.offset_conv = nn.Conv2d(mid_channels * 2, mid_channels, kernel_size=3, padding=1)
        self.dcn_pack = ModulatedDCNPack(mid_channels, mid_channels, 3, padding=1, deform_groups=8)
        self.emb_conv = nn.Sequential(*[nn.Conv2d(mid_channels, mid_channels, kernel_size=3, padding=1) for _ in range(emb_nums)])
        self.conv2 = nn.Conv2d(mid_channels, channels, kernel_size=3, padding=1)

    @inference_only
    def forward(self, x):
        in_dtype = x.dtype
        T, _, H, W = x.shape
        mid = self.mid_channels
        _weights_channels_last(self)
        y = _cl(_conv_relu(self.conv1, _cl(x.float())))                                     # :136
        # offsets / mask logits of the pair (i, t) = P[t] + Q[i]   (:141-142 and :75 of the pack, both linear)
        w_off, co = self.offset_conv.weight, self.dcn_pack.conv_offset
        a = F.conv2d(y, _cl(w_off[:, :mid]), None, padding=1)
        b = F.conv2d(y, _cl(w_off[:, mid:]), self.offset_conv.bias, padding=1)
        p = _nhwc(F.conv2d(a, co.weight, None, stride=co.stride, padding=co.padding)).contiguous()
        q = _nhwc(F.conv2d(b, co.weight, co.bias, stride=co.stride, padding=co.padding)).contiguous()
        y_nhwc = y.permute(0, 2, 3, 1)                                                      # contiguous view
        cor = torch.empty((T, T, H, W, mid), dtype=torch.float32, device=x.device)          # [i, t] correlation maps
        for i in range(T):              # one reference frame at a time: T pairs per im2col + GEMM (scratch T*HW*9*mid floats)
            d = self.dcn_pack.from_logits(y_nhwc, p, q[i:i + 1]).view(T, H, W, mid)         # :143, frame t aligned to frame i
            d.mul_(y_nhwc[i])                                                               # :144
            c = d.permute(0, 3, 1, 2)                                                       # channels-last [T, mid, H, W]
            for conv in self.emb_conv:      # no bias: constant over t, cancelled by the softmax over frames below
                c = F.conv2d(c, conv.weight, None, conv.stride, conv.padding)
            cor[i].copy_(c.permute(0, 2, 3, 1))
        fused = ops.temporal_softmax_fuse(cor, y_nhwc)                                      # :145-146 for every i
        out = _conv_relu(self.conv2, fused.permute(0, 3, 1, 2))                             # :149-151
        return out.to(in_dtype)                     # channels-last strides; Denoising2Aggergator hands NCHW-contiguous maps out


@AGGREGATORS.register_module()
class Denoising2Aggergator(nn.Module):
    """Multi-stage denoiser over the backbone's intermediate features (registered under the reference's spelling, :155-244).

    ``forward(x_noise, all_x)``: x_noise = the noisy clip's per-stage backbone features [T, in_channel[s], H_s, W_s];
    all_x = the detector's final feature maps.  Returns (per-stage denoised features, final features + the last stage's
    output), both tuples."""

    def __init__(self, in_channel=(256, 512, 1024, 2048), mid_channel=(64, 128, 256, 512), out_channel=(512, 1024, 2048, 512),
                 layer_name=('layer1', 'layer2', 'layer3', 'layer4'), rdb_blocks=(2, 2, 4, 2), rdb_channel_growth=(64, 64, 64, 64),
                 taf_embs=(3, 3, 3, 3), downsample=(True, True, False, False), with_rdb=(True, True, True, True),
                 with_taf=(True, True, True, True)):
        super().__init__()
        self.num_stage = len(in_channel)
        self.in_channel, self.mid_channel, self.out_channel = list(in_channel), list(mid_channel), list(out_channel)
        self.layer_name, self.rdb_blocks, self.taf_embs = list(layer_name), list(rdb_blocks), list(taf_embs)
        self.downsample, self.with_rdb, self.with_taf = list(downsample), list(with_rdb), list(with_taf)
        self.layers = nn.ModuleDict()
        for s, name in enumerate(self.layer_name):
            c_in = in_channel[s] + (out_channel[s - 1] if s else 0)
            self.layers[name + '_conv1'] = nn.Conv2d(c_in, in_channel[s], kernel_size=3, padding=1)
            if self.with_rdb[s]:
                self.layers[name + '_rdb'] = nn.Sequential(*[RDB(in_channel[s], rdb_channel_growth[s], 3) for _ in range(rdb_blocks[s])])
            if self.with_taf[s]:
                self.layers[name + '_taf'] = TemporalAttentionFusion(in_channel[s], mid_channel[s], emb_nums=taf_embs[s])
            self.layers[name + '_conv2'] = nn.Conv2d(in_channel[s], out_channel[s], kernel_size=3, padding=1,
                                                     stride=2 if downsample[s] else 1)

    @inference_only
    def forward(self, x_noise, all_x):
        denoised, carried = [], None
        last = self.num_stage - 1
        _weights_channels_last(self)
        x_noise = [_cl(t) for t in x_noise]             # channels-last once: every conv / cat below then stays in that layout
        for s, name in enumerate(self.layer_name):
            f = x_noise[s] if s == 0 else torch.cat((x_noise[s], carried), 1)               # :222-225
            x = self.layers[name + '_conv1'](f)
            if self.with_rdb[s]:
                x = self.layers[name + '_rdb'](x)
            if self.with_taf[s]:
                x = self.layers[name + '_taf'](x)
            res = x + x_noise[s]
            denoised.append(res.contiguous())                                               # :231
            carried = self.layers[name + '_conv2'](x if s == last else res)                 # :232-235
        fused = (all_x[-1] + carried).contiguous()                                          # :238-242: the same sum per level
        return tuple(denoised), tuple(fused for _ in all_x)
