"""B200-native drop-ins for mmtrack's AGGREGATORS (same registry names, ctor kwargs, forward
signatures and parameter names, so reference configs and checkpoints apply unchanged).

  SelsaAggregator  mmtracking/mmtrack/models/aggregators/selsa_aggregator.py:7-73
  EmbedAggregator  mmtracking/mmtrack/models/aggregators/embed_aggregator.py:8-81

The linear / conv projections stay library GEMMs (cuBLAS / cuDNN via nn.Linear / nn.Conv2d, as the
reference); the similarity-softmax-weighted-sum cores run in hand-written sm_100a kernels and never
materialise the [heads, N, M] weight tensor or the normalised embedding copies.  Inference path only
(the kernels have no backward).
"""
import torch
import torch.nn as nn

from . import ops
from .registry import AGGREGATORS, ConvModule, inference_only


@AGGREGATORS.register_module()
class SelsaAggregator(nn.Module):
    """Selsa aggregator module ("Sequence Level Semantics Aggregation for Video Object Detection").

    Args:
        in_channels (int): The number of channels of the features of proposal.
        num_attention_blocks (int): The number of attention blocks used in selsa aggregator module.
            Defaults to 16.
    """

    def __init__(self, in_channels, num_attention_blocks=16):
        super(SelsaAggregator, self).__init__()
        self.fc_embed = nn.Linear(in_channels, in_channels)
        self.ref_fc_embed = nn.Linear(in_channels, in_channels)
        self.fc = nn.Linear(in_channels, in_channels)
        self.ref_fc = nn.Linear(in_channels, in_channels)
        self.num_attention_blocks = num_attention_blocks
        self.impl = ops.IMPL_AUTO          # test hook: force the SIMT or tcgen05 kernel
        self.compute_dtype = torch.float32  # torch.bfloat16 selects bf16 tensor-core math (stated tolerance)

    @inference_only
    def forward(self, x, ref_x):
        """Aggregate the features `ref_x` of reference proposals.

        Args:
            x (Tensor): of shape [N, C]. N is the number of key frame proposals.
            ref_x (Tensor): of shape [M, C]. M is the number of reference frame proposals.

        Returns:
            Tensor: The aggregated features of key frame proposals with shape [N, C].
        """
        roi_n, ref_roi_n = x.shape[0], ref_x.shape[0]
        in_dtype = x.dtype
        x, ref_x = x.float(), ref_x.float()
        if ref_roi_n == 0:
            # no reference proposals: the reference's softmax over an empty axis and bmm with a [.., 0, d] operand give zeros,
            # so the result is fc(0) = the output bias in every row (selsa_aggregator.py:61-72)
            return self.fc(x.new_zeros(roi_n, x.shape[1])).to(in_dtype)
        k, v, use_vt = self.project_ref(ref_x)
        return self.attend(x, k, v, ref_roi_n, use_vt).to(in_dtype)

    def v_layout(self):
        """(V is kept transposed, row alignment in elements): the tensor-core kernel wants the reference axis contiguous for
        the P.V product, with a row stride that is a multiple of 16 bytes (TMA)."""
        d = self.fc_embed.out_features // self.num_attention_blocks
        return d == 64, (8 if self.compute_dtype == torch.bfloat16 else 4)

    def project_ref(self, ref_x, k_out=None, vt_out=None):
        """The reference-side projections of one aggregator layer: K = ref_fc_embed(ref_x) (selsa_aggregator.py:55) and
        V = ref_fc(ref_x) (:64), the latter emitted directly as V^T [C, M] WITHOUT its bias when d == 64 (softmax rows sum to
        one, so the bias passes through the attention unchanged, P (V0 + 1 b^T) = P V0 + b, and is added to the [N, C] result in
        ``attend`` instead of being broadcast into the [C, M] operand first).  They depend on the reference proposals only,
        which is what lets SelsaRoIHead cache them per reference frame.

        ``k_out`` [M, C] / ``vt_out`` [C, M(+pad)] (views into a cache) receive the results when given.
        Returns (k, v or V^T, v_is_transposed)."""
        ref_x = ref_x.float()
        M = ref_x.shape[0]
        use_vt, align = self.v_layout()
        use_vt = use_vt and M > 0
        if k_out is not None:
            torch.addmm(self.ref_fc_embed.bias, ref_x, self.ref_fc_embed.weight.t(), out=k_out)
            k = k_out
        else:
            k = self.ref_fc_embed(ref_x)                       # :55
        if not use_vt:
            assert vt_out is None
            return k, self.ref_fc(ref_x), False
        if vt_out is None:
            # row stride of V^T: a multiple of 16 bytes.  fp32: 4 elements -- M = 4500 then needs no padding and the GEMM
            # writes straight into a contiguous buffer (a strided `out=` costs a temporary + a copy pass per layer)
            ldv = (M + align - 1) // align * align
            vt = torch.empty((self.ref_fc.out_features, ldv), dtype=torch.float32, device=ref_x.device)
            if ldv != M:
                vt[:, M:].zero_()
            torch.mm(self.ref_fc.weight.float(), ref_x.t(), out=vt[:, :M])
            return k, vt, True
        torch.mm(self.ref_fc.weight.float(), ref_x.t(), out=vt_out)
        return k, vt_out, True

    def out_bias(self, v_transposed):
        """The bias of ``attend(..., with_bias=False)``'s result: fc.bias, plus fc.weight . ref_fc.bias when V^T was projected
        without its bias (``project_ref``): fc(o + b_v) = o W^T + (W b_v + b).  Cached per weight version."""
        if not v_transposed:
            return self.fc.bias
        key = tuple((t._version, t.data_ptr()) for t in (self.fc.weight, self.fc.bias, self.ref_fc.bias))
        if getattr(self, '_out_bias', None) is None or self._out_bias[0] != key:
            with torch.no_grad():
                b = torch.addmv(self.fc.bias.double(), self.fc.weight.double(), self.ref_fc.bias.double()).float()
            self._out_bias = (key, b.contiguous())
        return self._out_bias[1]

    def attend(self, x, k, v, M, v_transposed, q=None, with_bias=True):
        """fc(softmax(fc_embed(x) K^T / sqrt(d)) V) for the first ``M`` reference rows of k / columns of V^T
        (selsa_aggregator.py:50,57-72).  ``q``: fc_embed(x) when the caller already computed it.
        ``with_bias=False``: the result lacks ``out_bias(v_transposed)`` -- for callers that add it together with their
        residual and ReLU in one pass (``ops.selsa_residual_relu_``)."""
        roi_n = x.shape[0]
        if q is None:
            q = self.fc_embed(x.float())                       # :50
        k = k[:M]
        if not v_transposed:
            v = v[:M]
        if self.compute_dtype == torch.bfloat16:
            q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
        o = ops.selsa_attention(q, k, v, self.num_attention_blocks, v_transposed=v_transposed, impl=self.impl)  # :61-70
        if not with_bias:
            return torch.nn.functional.linear(o.view(roi_n, -1), self.fc.weight)
        if v_transposed:
            o += self.ref_fc.bias.float()
        return self.fc(o.view(roi_n, -1))                      # :72


@AGGREGATORS.register_module()
class EmbedAggregator(nn.Module):
    """Embedding convs to aggregate multi feature maps ("Flow-Guided Feature Aggregation").

    Args:
        num_convs (int): Number of embedding convs.
        channels (int): Channels of embedding convs. Defaults to 256.
        kernel_size (int): Kernel size of embedding convs, Defaults to 3.
        norm_cfg (dict): Configuration of normlization method after each conv. Defaults to None.
        act_cfg (dict): Configuration of activation method after each conv. Defaults to dict(type='ReLU').
    """

    def __init__(self, num_convs=1, channels=256, kernel_size=3, norm_cfg=None, act_cfg=dict(type='ReLU')):
        super(EmbedAggregator, self).__init__()
        assert num_convs > 0, 'The number of convs must be bigger than 1.'
        self.embed_convs = nn.ModuleList()
        for i in range(num_convs):
            if i == num_convs - 1:
                new_norm_cfg = None
                new_act_cfg = None
            else:
                new_norm_cfg = norm_cfg
                new_act_cfg = act_cfg
            self.embed_convs.append(
                ConvModule(in_channels=channels, out_channels=channels, kernel_size=kernel_size,
                           padding=(kernel_size - 1) // 2, norm_cfg=new_norm_cfg, act_cfg=new_act_cfg))

    def _embed(self, t):
        for embed_conv in self.embed_convs:
            t = embed_conv(t)
        return t

    @inference_only
    def forward(self, x, ref_x):
        """Aggregate reference feature maps `ref_x`.

        Args:
            x (Tensor): of shape [1, C, H, W]
            ref_x (Tensor): of shape [N, C, H, W]. N is the number of reference feature maps.

        Returns:
            Tensor: The aggregated feature map with shape [1, C, H, W].
        """
        assert len(x.shape) == 4 and len(x) == 1, "Only support 'batch_size == 1' for x"
        # the embed convs run in the module's own dtype, as in the reference (a bf16 / fp16 model keeps its convs there);
        # the weighting kernel up-converts the embeddings and computes in fp32
        x_embed = self._embed(x)                                # embed_aggregator.py:68-70
        ref_x_embed = self._embed(ref_x)                        # :73-75
        # :71-81 fused: cosine over C, softmax over frames, weighted sum of the raw ref features
        return ops.embed_weighted_sum(x_embed, ref_x_embed, ref_x).to(x.dtype)

    @torch.no_grad()
    def forward_fused_warp(self, x, raw_ref_x, flows, key_slot=-1):
        """FGFA step with the warp fused into the weighting (mmtracking/mmtrack/models/vid/fgfa.py:275-283):
        ``raw_ref_x`` are the un-warped memory features, ``flows`` the key->ref flows.  The warped maps are
        produced once for the embed convs; the weighted sum re-warps on the fly instead of re-reading them.
        Slot ``key_slot`` (the key frame's own position in the memory) is taken from ``x`` un-warped."""
        from .motion import flow_warp_feats
        assert len(x.shape) == 4 and len(x) == 1, "Only support 'batch_size == 1' for x"
        warped = flow_warp_feats(raw_ref_x, flows)
        if key_slot >= 0:
            warped[key_slot] = x[0]
        x_embed = self._embed(x)
        ref_x_embed = self._embed(warped)
        return ops.fgfa_warp_weighted_sum(x_embed, ref_x_embed, raw_ref_x, flows, key_x=x, key_slot=key_slot).to(x.dtype)
