"""Polyformer layer on the B200 kernels: the reference's code/networks/polyformer.py:8-106 — a squeezed
attention block WITHOUT the expansion FFN (both CrossAttFeatTrans run multi-mode, their value modes are soft-aggregated,
segtran_shared.py:452-457) applied to a 2x average-pooled CNN feature map and added back to it.

Same classes, constructor signatures and parameter names as the reference (`Polyformer(feat_dim, chan_axis=1, args=None)`,
`polyformer_layers.0.{in_ator_trans,ator_out_trans,attractors}`), so its checkpoints load.  The attention, value
projection, mode aggregation, LayerNorm, bilinear up-sampling and residual run on the library's kernels; the 2x2 average
pooling in front is the stock PyTorch op (it belongs to the CNN side of the block, like the backbone)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import Parameter

from .. import ops
from .segtran_shared import CrossAttFeatTrans, SegtranConfig, SegtranInitWeights


class PolyformerLayer(SegtranInitWeights):
    def __init__(self, name, config):
        super().__init__(config)
        self.name = name
        self.chan_axis = config.chan_axis
        self.feat_dim = config.feat_dim
        self.num_attractors = config.num_attractors
        self.qk_have_bias = config.qk_have_bias
        self.in_ator_trans = CrossAttFeatTrans(config, name + '-in-squeeze')
        self.ator_out_trans = CrossAttFeatTrans(config, name + '-squeeze-out')
        self.attractors = Parameter(torch.randn(1, self.num_attractors, self.feat_dim))
        self.infeat_norm_layer = nn.LayerNorm(self.feat_dim, eps=1e-12, elementwise_affine=False)
        self.poly_do_layernorm = config.poly_do_layernorm
        if self.poly_do_layernorm:
            raise NotImplementedError("segtran_b200: poly_do_layernorm (off in the reference: it costs 1-2 % accuracy, "
                                      "polyformer.py:42-45)")
        print("Polyformer layer {}: {} attractors, {} modes, {} channels, {} layernorm".format(
            name, self.num_attractors, config.num_modes, self.feat_dim, 'with' if self.poly_do_layernorm else 'no'))
        self.pool2x = nn.AvgPool2d(2)
        self.apply(self.init_weights)
        self.apply(self.tie_qk)                       # after the weight initialisation (reference :31-33)
        self.apply(self.add_identity_bias)

    def forward(self, in_feat):
        """in_feat [B,C,H,W] (chan_axis = 1) -> in_feat + upsample(squeezed attention over the 2x-pooled map)."""
        B = in_feat.shape[0]
        in_feat_half0 = self.pool2x(in_feat)
        in_feat_half = in_feat_half0.transpose(self.chan_axis, -1)            # reference :41 (swaps C with the LAST dim)
        vfeat = in_feat_half.reshape(B, -1, self.feat_dim)
        if not vfeat.is_contiguous():
            vfeat = vfeat.contiguous()
        att = self.in_ator_trans(self.attractors, vfeat)                      # attractors are batch-invariant: batch 1
        vfeat_out = self.ator_out_trans(vfeat, att)
        out_half = vfeat_out.view(in_feat_half.shape).transpose(self.chan_axis, -1)
        out_feat = ops.resize_linear(out_half.contiguous(), tuple(in_feat.shape[2:]))      # bilinear, align_corners=False
        return ops.add(in_feat, out_feat)


class Polyformer(nn.Module):
    def __init__(self, feat_dim, chan_axis=1, args=None):
        config = SegtranConfig()
        if args is None:
            config.num_attractors = 256
            config.num_modes = 4
            config.tie_qk_scheme = 'loose'
            config.qk_have_bias = True
            config.pos_code_type = 'lsinu'
        else:
            config.num_attractors = args.num_attractors
            config.num_modes = args.num_modes if args.num_modes != -1 else 4
            config.tie_qk_scheme = args.tie_qk_scheme
            config.qk_have_bias = args.qk_have_bias
            config.pos_code_type = args.pos_code_type
        config.num_layers = 1
        config.in_feat_dim = feat_dim
        config.feat_dim = feat_dim
        config.min_feat_dim = feat_dim
        config.v_has_bias = False
        config.has_FFN = False                        # aggregate the value modes only, no transformation (reference :84-87)
        config.ablate_multihead = False
        config.chan_axis = chan_axis
        config.poly_do_layernorm = False
        super().__init__()
        layers = []
        for i in range(config.num_layers):
            if i > 0:
                config.only_first_linear = False
            layers.append(PolyformerLayer(str(i), config))
        self.polyformer_layers = nn.Sequential(*layers)

    def forward(self, in_feat):
        return self.polyformer_layers(in_feat)
