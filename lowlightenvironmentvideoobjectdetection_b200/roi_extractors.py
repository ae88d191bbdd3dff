"""B200-native drop-ins for the RoI extractors on the hot path.

  BaseRoIExtractor    mmdetection/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:9-83
  SingleRoIExtractor  mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:9-108
                      + the mmtrack override that swallows ``ref_feats=`` (**kwargs),
                      mmtracking/mmtrack/models/roi_heads/roi_extractors/single_level_roi_extractor.py:7-19
  TemporalRoIAlign    mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:9-207

Same registry names, ctor kwargs, forward signatures and parameter names (``embed_network.conv.*``).
"""
import torch
import torch.nn as nn

from . import ops
from .registry import ROI_EXTRACTORS, ConvModule, force_fp32, inference_only


class BaseRoIExtractor(nn.Module):
    """Base class for RoI extractor.

    Args:
        roi_layer (dict): Specify RoI layer type and arguments.
        out_channels (int): Output channels of RoI layers.
        featmap_strides (List[int]): Strides of input feature maps.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides):
        super(BaseRoIExtractor, self).__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        """int: Number of input feature maps."""
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        """``getattr(mmcv.ops, type)(spatial_scale=1/s, **cfg)`` per level (base_roi_extractor.py:32-55)."""
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert hasattr(ops, layer_type)
        layer_cls = getattr(ops, layer_type)
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def roi_rescale(self, rois, scale_factor):
        """Scale RoI coordinates by scale factor (base_roi_extractor.py:57-79)."""
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        w = rois[:, 3] - rois[:, 1]
        h = rois[:, 4] - rois[:, 2]
        new_w = w * scale_factor
        new_h = h * scale_factor
        x1 = cx - new_w * 0.5
        x2 = cx + new_w * 0.5
        y1 = cy - new_h * 0.5
        y2 = cy + new_h * 0.5
        return torch.stack((rois[:, 0], x1, y1, x2, y2), dim=-1)


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(BaseRoIExtractor):
    """Extract RoI features from a single level feature map (FPN level mapping kept for multi-level
    inputs; the hot path's R-50-DC5 configs have one stride-16 level and take the single-level branch).

    Args:
        roi_layer (dict): Specify RoI layer type and arguments.
        out_channels (int): Output channels of RoI layers.
        featmap_strides (List[int]): Strides of input feature maps.
        finest_scale (int): Scale threshold of mapping to level 0. Default: 56.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super(SingleRoIExtractor, self).__init__(roi_layer, out_channels, featmap_strides)
        self.finest_scale = finest_scale

    def map_roi_levels(self, rois, num_levels):
        """single_level_roi_extractor.py:32-51."""
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def _extract(self, feats, rois, roi_scale_factor=None):
        out_size = self.roi_layers[0].output_size
        num_levels = len(feats)
        if num_levels == 1:
            if len(rois) == 0:
                return feats[0].new_zeros(0, self.out_channels, *out_size)
            return self.roi_layers[0](feats[0], rois)
        roi_feats = feats[0].new_zeros(rois.size(0), self.out_channels, *out_size)
        target_lvls = self.map_roi_levels(rois, num_levels)
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        for i in range(num_levels):
            inds = (target_lvls == i).nonzero(as_tuple=False).squeeze(1)
            if inds.numel() > 0:
                roi_feats[inds] = self.roi_layers[i](feats[i], rois[inds]).to(roi_feats.dtype)
        return roi_feats

    @force_fp32(apply_to=('feats', ), out_fp16=True)
    @inference_only
    def forward(self, feats, rois, roi_scale_factor=None, **kwargs):
        """Forward function (``**kwargs`` swallows ``ref_feats=`` as the mmtrack override does)."""
        return self._extract(feats, rois, roi_scale_factor)


@ROI_EXTRACTORS.register_module()
class TemporalRoIAlign(SingleRoIExtractor):
    """Temporal RoI Align module ("Temporal ROI Align for Video Object Recognition").

    Args:
        num_most_similar_points (int): Number of the most similar points in the Most Similar RoI Align.
            Defaults to 2.
        num_temporal_attention_blocks (int): Number of temporal attention blocks in the Temporal
            Attentional Feature Aggregation.  If not greater than 0, the RoI features are averaged with
            the Most Similar RoI features instead.  Defaults to 4.

    Device pipeline of ``forward(feats, rois, ref_feats=...)`` (everything stays NHWC, C contiguous):
      1. vod_nchw_to_nhwc   key / reference maps -> NHWC (+ per-pixel L2 norm, unit-norm bf16 copy)
      2. vod_roi_align_fwd  key RoI features straight into slot 0 of x_all [T+1, N, 49, C]
      3. vod_msra_topk_sample  bf16 tcgen05 GEMM with a register top-8 epilogue (the 2.1 GB similarity
                            tensor is never written) + exact fp32 re-score/top-k/softmax/gather into
                            slots 1..T of x_all
      4. embed_network      3x3 conv (cuDNN) over the KEY slot of x_all only; G = ek_head . W_head (one batched
                            cuBLAS GEMM); vod_tafa_keyproj_logits contracts x_all with G -- the reference-slot
                            embeddings (15/16 of the conv's 1.11 TFLOP at cfg 3) are never computed
      5. vod_tafa_weighted_sum_logits  softmax over frames, weighted sum, written as [N, C, 7, 7]
      (few stacked frames / unsupported shapes: conv over all (T+1)*N patches + vod_tafa_weighted_sum)
    """

    def __init__(self, num_most_similar_points=2, num_temporal_attention_blocks=4, *args, **kwargs):
        super(TemporalRoIAlign, self).__init__(*args, **kwargs)
        self.num_most_similar_points = num_most_similar_points
        self.num_temporal_attention_blocks = num_temporal_attention_blocks
        if self.num_temporal_attention_blocks > 0:
            self.embed_network = ConvModule(self.out_channels, self.out_channels, 3, padding=1, conv_cfg=None,
                                            norm_cfg=None, act_cfg=None)
        self.impl = ops.IMPL_AUTO  # test hook: force the exact SIMT scan or the tcgen05 candidate GEMM
        # key-projected attention logits (embed conv on the key slot only): None = when it pays (>= keyproj_min_frames
        # stacked frames), True / False = forced on (where the shape is supported) / off
        self.keyproj = None
        self.keyproj_min_frames = 8
        # the key-slot embed conv + G product depend on the key RoI features only: run them on a side stream next to the
        # most-similar search (False = everything on one stream)
        self.overlap = True
        # dtype of the key-projected G operand (a library GEMM's output, 70 % of the logits kernel's bytes): None = follow the
        # caller's library-math switch -- bf16 when torch.backends.cuda.matmul.allow_tf32 permits reduced-precision GEMMs
        # (features then agree with fp32 to ~5e-4 instead of ~1e-4, still inside the 1e-3 bar), fp32 otherwise;
        # torch.float32 / torch.bfloat16 force it
        self.keyproj_g_dtype = None

    def _stack_key_and_refs(self, feat, rois, ref_feat, return_indices=False, want_prepared=False):
        """RoIAlign(key) + most-similar RoI features, stacked as x_all [T+1, N, P, C] (NHWC rows)."""
        C = ref_feat.shape[1]
        key_nhwc, _, _ = ops.to_nhwc(feat)
        want_tc = ref_feat.is_cuda and C % 64 == 0 and C <= 512 and self.impl != ops.IMPL_SIMT
        ref_nhwc, ref_norm, ref_unit = ops.to_nhwc(ref_feat, want_norm=True, want_unit_bf16=want_tc)
        return self._stack_from_layout(key_nhwc.contiguous(), rois, ref_nhwc.contiguous(), ref_norm, ref_unit, return_indices,
                                       want_prepared)

    def _stack_from_layout(self, key_nhwc, rois, ref_nhwc, ref_norm, ref_unit, return_indices=False, want_prepared=False):
        """Same, from maps that are already laid out: key_nhwc [1,H,W,C], ref_nhwc [T,H,W,C] fp32 contiguous, ref_norm [T*H*W],
        ref_unit [T*H*W, C] bf16 (or None) -- the form a reference-frame cache holds (heads.RefFrameCache).

        ``want_prepared``: also returns what ``_tafa`` needs from the key slot alone (the key-projected G operand), computed
        on a side stream WHILE the most-similar search runs (it only depends on the key RoI features)."""
        layer = self.roi_layers[0]
        ph, pw = layer.output_size
        N = rois.shape[0]
        T, C = ref_nhwc.shape[0], ref_nhwc.shape[3]
        x_all = torch.empty((T + 1, N, ph * pw, C), dtype=torch.float32, device=key_nhwc.device)
        # temporal_roi_align.py:186 -- key RoI features, emitted as [N, 49, C] rows into slot 0
        ops.roi_align_nhwc(key_nhwc, rois, (ph, pw), layer.spatial_scale, layer.sampling_ratio,
                           layer.aligned, out_nhwc=True, out=x_all[0])
        prepared, branch = None, None
        if want_prepared and self.num_temporal_attention_blocks > 0 and self._keyproj_chunk(T + 1, ph * pw, C):
            if self.overlap and x_all.is_cuda:
                with ops.fork(x_all.device) as branch:
                    prepared = self._keyproj_prepare(x_all[0], N, ph, pw, C, T + 1)
            else:
                prepared = self._keyproj_prepare(x_all[0], N, ph, pw, C, T + 1)
        # temporal_roi_align.py:99-181
        res = ops.msra_topk_sample(x_all[0].view(N * ph * pw, C), ref_nhwc,
                                   k=self.num_most_similar_points, ref_norm=ref_norm, ref_unit=ref_unit,
                                   impl=self.impl, return_indices=return_indices, out=x_all[1:])
        if branch is not None:
            branch.join()
        if want_prepared:
            return x_all, prepared
        return (x_all, res[1], res[2]) if return_indices else x_all

    def forward_from_layout(self, key_nhwc, rois, ref_nhwc, ref_norm, ref_unit, out=None):
        """``forward(feats, rois, ref_feats=...)`` for maps already held in NHWC with their norms / unit-norm bf16 copy
        (reference-frame cache): skips the per-call layout pass over the T reference maps.  ``out``: optional preallocated
        destination of N*49*C floats (the layout follows ``roi_layers[0].channels_last_out``)."""
        out_size = self.roi_layers[0].output_size
        if len(rois) == 0:
            return key_nhwc.new_zeros(0, self.out_channels, *out_size)
        x_all, prepared = self._stack_from_layout(key_nhwc, rois, ref_nhwc, ref_norm, ref_unit, want_prepared=True)
        return self._tafa(x_all, out_size[0], out_size[1], out=out, prepared=prepared)

    @torch.no_grad()
    def most_similar_roi_align(self, roi_feats, ref_feats):
        """[roi_n, C, h, w], [img_n, C, H, W] -> [img_n, roi_n, C, h, w] (temporal_roi_align.py:99-181)."""
        roi_n, C, rh, rw = roi_feats.shape
        ref_nhwc, ref_norm, ref_unit = ops.to_nhwc(ref_feats, want_norm=True,
                                                   want_unit_bf16=(C % 64 == 0 and C <= 512 and self.impl != ops.IMPL_SIMT))
        rows = roi_feats.float().permute(0, 2, 3, 1).contiguous().view(roi_n * rh * rw, C)
        out = ops.msra_topk_sample(rows, ref_nhwc.contiguous(), k=self.num_most_similar_points, ref_norm=ref_norm,
                                   ref_unit=ref_unit, impl=self.impl)
        return out.view(ref_feats.shape[0], roi_n, rh, rw, C).permute(0, 1, 4, 2, 3)

    @torch.no_grad()
    def temporal_attentional_feature_aggregation(self, x, ref_x):
        """[1, roi_n, C, h, w], [img_n, roi_n, C, h, w] -> [roi_n, C, h, w] (temporal_roi_align.py:44-97)."""
        x_all = torch.cat((x, ref_x), dim=0).float()
        img_n, roi_n, C, rh, rw = x_all.shape
        rows = x_all.permute(0, 1, 3, 4, 2).contiguous().view(img_n, roi_n, rh * rw, C)
        return self._tafa(rows, rh, rw)

    def _conv_weight_cl(self, conv):
        # the weight in channels_last too (cached): otherwise cuDNN re-lays it out on every call (a 2.4 M element copy)
        key = (conv.weight._version, conv.weight.data_ptr())
        if getattr(self, '_w_cl', None) is None or self._w_cl[0] != key:
            self._w_cl = (key, conv.weight.detach().contiguous(memory_format=torch.channels_last))
        return self._w_cl[1]

    def _keyproj_weight(self, conv, heads, cc, dtype=torch.float32):
        """conv weight [C, C, 3, 3] -> [heads, C/heads, 9*C] with columns ordered (channel chunk, tap, channel in chunk):
        the right-hand side of the G GEMM, laid out so that each (RoI, chunk) CTA of the logits kernel reads contiguous rows."""
        key = (conv.weight._version, conv.weight.data_ptr(), heads, cc, dtype)
        if getattr(self, '_w_kp', None) is None or self._w_kp[0] != key:
            w = conv.weight.detach().float()
            C = w.shape[0]
            wr = w.view(heads, C // heads, C // cc, cc, 9).permute(0, 1, 2, 4, 3).reshape(heads, C // heads, 9 * C)
            self._w_kp = (key, wr.contiguous().to(dtype))
        return self._w_kp[1]

    def _g_dtype(self):
        if self.keyproj_g_dtype is not None:
            return self.keyproj_g_dtype
        return torch.bfloat16 if torch.backends.cuda.matmul.allow_tf32 else torch.float32

    def _keyproj_chunk(self, T1, P, C):
        """Channel-chunk width for the key-projected path, 0 when the full-embedding path must be taken."""
        heads = self.num_temporal_attention_blocks
        conv = self.embed_network.conv
        if self.keyproj is False or (self.keyproj is None and T1 < self.keyproj_min_frames):
            return 0
        if (tuple(conv.kernel_size), tuple(conv.padding), tuple(conv.stride), tuple(conv.dilation), conv.groups) != \
                ((3, 3), (1, 1), (1, 1), (1, 1), 1) or C % heads != 0:
            return 0
        return ops.tafa_keyproj_chunk(T1, P, C, heads)

    def _keyproj_prepare(self, x_key, N, rh, rw, C, T1):
        """temporal_roi_align.py:72-93 with the embed conv applied to the KEY slot only: the conv is linear, so
        <conv(x_t), ek>_head = <x_t, ek_head . W_head>; G = ek_head . W_head is one batched GEMM with the FLOPs of one slot's
        conv.  x_key [N, P, C] (slot 0 of x_all) -> (G [heads, N*P, 9*C], channel-chunk width)."""
        heads = self.num_temporal_attention_blocks
        conv = self.embed_network.conv
        P = rh * rw
        cc = self._keyproj_chunk(T1, P, C)
        key_patches = x_key.view(N, rh, rw, C).permute(0, 3, 1, 2)
        gd = self._g_dtype()
        # the conv runs without its bias; the bias add and the conversion to the GEMM's operand type are ONE pass
        # (torch.add computes in fp32 and rounds once into ``out``: the same values as add-then-convert, one launch fewer)
        ek = torch.nn.functional.conv2d(key_patches, self._conv_weight_cl(conv), None, 1, 1)
        ek = ek.permute(0, 2, 3, 1).contiguous().view(N * P, C)                   # no copy (channels_last)
        if conv.bias is not None:
            ek = torch.add(ek, conv.bias, out=torch.empty((N * P, C), dtype=gd, device=ek.device))
        else:
            ek = ek.to(gd)
        # [heads, N*P, 9*C]; bf16: the GEMM runs on bf16 operands and writes half the bytes (it is write-bound)
        G = torch.bmm(ek.view(N * P, heads, C // heads).transpose(0, 1), self._keyproj_weight(conv, heads, cc, gd))
        return G, cc

    def _tafa(self, x_all, rh, rw, out=None, prepared=None):
        T1, N, P, C = x_all.shape
        cl_out = self.roi_layers[0].channels_last_out
        heads = self.num_temporal_attention_blocks
        if heads > 0:
            conv = self.embed_network.conv
            cc = self._keyproj_chunk(T1, P, C)
            if cc:
                # the contraction of x_t with G runs in vod_tafa_keyproj_logits
                G, cc = prepared if prepared is not None else self._keyproj_prepare(x_all[0], N, rh, rw, C, T1)
                parts = ops.tafa_keyproj_logits(x_all, G, (rh, rw), heads, cc)
                out = ops.tafa_weighted_sum_logits(x_all, parts, heads, out_nhwc=cl_out, out=out)
            else:
                # embed conv on the channels_last view: logical [(T+1)*N, C, 7, 7], memory [(T+1)*N, 7, 7, C]
                patches = x_all.view(T1 * N, rh, rw, C).permute(0, 3, 1, 2)
                # temporal_roi_align.py:74 -- the conv runs without its bias; the bias is added on load inside
                # the weighting kernel (saves a full read+write pass over the [T+1,N,49,C] embedding)
                emb = torch.nn.functional.conv2d(patches, self._conv_weight_cl(conv), None, conv.stride, conv.padding,
                                                 conv.dilation, conv.groups)
                emb = emb.permute(0, 2, 3, 1).contiguous().view(T1, N, P, C)  # no copy: cuDNN keeps channels_last
                out = ops.tafa_weighted_sum(x_all, emb, heads, emb_bias=conv.bias, out_nhwc=cl_out, out=out)
        else:
            out = ops.tafa_weighted_sum(x_all, None, 0, out_nhwc=cl_out, out=out)   # plain mean, :203-206
        if cl_out:
            return out.view(N, rh, rw, C).permute(0, 3, 1, 2)              # logical [N,C,7,7], channels_last strides
        return out.view(N, C, rh, rw)

    @force_fp32(apply_to=('feats', 'ref_feats'), out_fp16=True)
    @inference_only
    def forward(self, feats, rois, roi_scale_factor=None, ref_feats=None):
        """Forward function."""
        if ref_feats is None:
            # RoI features of reference-frame proposals: plain RoIAlign (temporal_roi_align.py:188-191)
            return self._extract(feats, rois, roi_scale_factor)
        assert len(feats) == 1, 'TemporalRoIAlign hot path expects a single feature level (R-50-DC5 configs)'
        out_size = self.roi_layers[0].output_size
        if len(rois) == 0:
            return feats[0].new_zeros(0, self.out_channels, *out_size)
        # only the last level of the reference maps is used (:195-196)
        x_all, prepared = self._stack_key_and_refs(feats[0], rois, ref_feats[-1], want_prepared=True)
        return self._tafa(x_all, out_size[0], out_size[1], prepared=prepared).to(feats[0].dtype)
