"""Backbone-side modules of the hot path, named and parameterised like the reference
(/root/reference/libs/modeling/multimodal_backbones.py) so checkpoints load strictly:

* ``Alignment`` multiway transformer (:973-1235) with ``MultiWayTransformer`` / ``MultiHeadAttention`` / ``FFN``
* ``ConvTransformerBackbone`` (:626-841), registered as ``convTransformer``
* ``fusion_module`` (:367-619) with ``MaxSigmoidCSPLayerWithTwoConv`` / ``MaxSigmoidAttnBlock`` / ``downsample``
* ``Downsample_pyramid_levels`` (:22-48)

Every ``forward`` keeps the reference's signature and channels-first tensors and runs on the sm_100a
kernels (``.._fwd``).  The reference hard-codes the guide length 224 in ``fusion_module``
(guide_channels=224 x10, ``Conv1d(12, 224, 1)``); here it is the ``max_len`` argument
(default 224, identical parameters at the reference configuration).
"""
from __future__ import annotations

import torch
from torch import nn

from .blocks import LayerNorm, MaskedConv1D, MaskedMHCA, TransformerBlock, get_sinusoid_encoding
from .models import register_multimodal_backbone


class Downsample_pyramid_levels(nn.Module):
    """Depthwise k=3 stride-``scale_factor`` conv -> LayerNorm (:22-48)."""

    def __init__(self, n_embd, scale_factor):
        super().__init__()
        assert scale_factor == 1 or scale_factor % 2 == 0
        self.n_embd, self.x_stride = n_embd, scale_factor
        self.down_conv = MaskedConv1D(n_embd, n_embd, kernel_size=3, stride=scale_factor, padding=1,
                                      groups=n_embd, bias=False)
        self.down_norm = LayerNorm(n_embd)

    def forward(self, x, mask):
        from .. import _fwd
        return _fwd.pyramid_downsample(self, x, mask)


class MaxSigmoidAttnBlock(nn.Module):
    """Max-sigmoid guide attention (:118-197): per-head gate sigmoid(max_n <embed, guide_n>/sqrt(hc) + b)
    applied to a k=3 projection of the input."""

    def __init__(self, in_channels, out_channels, guide_channels, embed_channels, kernel_size=3, padding=1,
                 num_heads=1, use_depthwise=False, with_scale=False, conv_cfg=None, norm_cfg=None,
                 init_cfg=None, use_einsum=True):
        super().__init__()
        assert out_channels % num_heads == 0 and embed_channels % num_heads == 0
        assert embed_channels == in_channels, "embed_conv variant is not on the hot path"
        self.num_heads = num_heads
        self.head_channels = embed_channels // num_heads
        self.embed_conv = None
        self.guide_fc = nn.Linear(guide_channels, embed_channels)
        self.bias = nn.Parameter(torch.zeros(num_heads))
        assert not with_scale
        self.scale = 1.0
        self.project_conv = MaskedConv1D(in_channels, out_channels, kernel_size, stride=1, padding=padding)

    def forward(self, x, guide, mask):
        from .. import _fwd
        return _fwd.maxsig_attn_block(self, x, guide, mask)


class MaxSigmoidCSPLayerWithTwoConv(nn.Module):
    """CSP layer: 1x1 split conv, ``num_blocks`` chained MaskedMHCA(mid, 4 heads), MaxSigmoid block, 1x1 merge
    (:199-256; base class :51-116)."""

    def __init__(self, in_channels, out_channels, guide_channels, embed_channels, num_heads=1, expand_ratio=0.5,
                 num_blocks=1, with_scale=False, add_identity=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 init_cfg=None, use_einsum=True):
        super().__init__()
        self.mid_channels = int(out_channels * expand_ratio)
        self.main_conv = MaskedConv1D(in_channels, 2 * self.mid_channels, 1)
        self.final_conv = MaskedConv1D((3 + num_blocks) * self.mid_channels, out_channels, 1)
        self.blocks = nn.ModuleList(
            MaskedMHCA(self.mid_channels, n_head=4, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0, proj_pdrop=0)
            for _ in range(num_blocks))
        self.attn_block = MaxSigmoidAttnBlock(self.mid_channels, self.mid_channels, guide_channels=guide_channels,
                                              embed_channels=embed_channels, num_heads=num_heads,
                                              with_scale=with_scale)

    def forward(self, x, guide, mask):
        from .. import _fwd
        return _fwd.csp_layer(self, x, guide, mask)


class downsample(nn.Module):
    """Dense k=3 stride-2 conv (+bias) -> LayerNorm -> SiLU (:336-356)."""

    def __init__(self, n_embd, scale_factor=2):
        super().__init__()
        assert scale_factor == 1 or scale_factor % 2 == 0
        self.n_embd, self.scale_factor, self.x_stride = n_embd, scale_factor, scale_factor
        k = scale_factor + 1 if scale_factor > 1 else 3
        self.down_conv = MaskedConv1D(n_embd, n_embd, kernel_size=k, stride=scale_factor, padding=k // 2)
        self.down_norm = LayerNorm(n_embd)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x, mask):
        from .. import _fwd
        return _fwd.fusion_downsample(self, x, mask)


class fusion_module(nn.Module):
    """RepVL-PAN style top-down / bottom-up fusion of a 6-level pyramid, guided by the other modality
    (:367-619)."""

    TD_HEADS = (8, 4, 4, 4, 4)     # :420-467
    BU_HEADS = (8, 8, 8, 8, 8)     # :481-527

    def __init__(self, n_embd, max_len=224, n_levels=6):
        super().__init__()
        self.n_embd, self.scale_factor, self.max_len = n_embd, 2, max_len
        self.in_channels = [max_len >> l for l in range(n_levels)]
        self.text_enhancer = MaskedMHCA(n_embd, n_head=4, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0, proj_pdrop=0)
        shared_down = downsample(n_embd, scale_factor=2)
        self.downsample_layers = nn.ModuleList([shared_down] * (n_levels - 1))       # one module, 5 aliases (:400-408)

        def csp(heads):
            return MaxSigmoidCSPLayerWithTwoConv(in_channels=2 * n_embd, out_channels=n_embd, guide_channels=max_len,
                                                 embed_channels=n_embd // 2, num_heads=heads, expand_ratio=0.5,
                                                 num_blocks=3)

        self.top_down_layers = nn.ModuleList(csp(h) for h in self.TD_HEADS[:n_levels - 1])
        self.bottom_up_layers = nn.ModuleList(csp(h) for h in self.BU_HEADS[:n_levels - 1])
        self.projections = nn.ModuleList(MaskedConv1D(n_embd, n_embd, 1) for _ in range(3))   # unused (:537-540)
        self.pool_size, self.num_feats = 4, 3
        self.match_projection = nn.Conv1d(3 * self.pool_size, max_len, 1)

    def forward(self, img_feats, txt_feats, mask_img, mask_txt):
        from .. import _fwd
        return _fwd.fusion_forward(self, img_feats, txt_feats, mask_img, mask_txt)


@register_multimodal_backbone("convTransformer")
class ConvTransformerBackbone(nn.Module):
    """Conv embedding -> stem transformers -> shared depthwise pyramid -> two fusion passes (:626-841)."""

    def __init__(self, n_in_V, n_in_A, n_embd, n_head, n_embd_ks, max_len, arch=(2, 2, 5), scale_factor=2,
                 with_ln=False, attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0, use_abs_pe=False):
        super().__init__()
        assert len(arch) == 3
        self.arch, self.max_len, self.scale_factor, self.use_abs_pe = arch, max_len, scale_factor, use_abs_pe
        self.n_embd, self.n_head = n_embd, n_head
        self.relu = nn.GELU()
        if use_abs_pe:
            self.register_buffer("pos_embd", get_sinusoid_encoding(max_len, n_embd) / (n_embd ** 0.5),
                                 persistent=False)
        self.embd_V, self.embd_A = nn.ModuleList(), nn.ModuleList()
        self.embd_norm_V, self.embd_norm_A = nn.ModuleList(), nn.ModuleList()
        for idx in range(arch[0]):
            cv, ca = (n_in_V, n_in_A) if idx == 0 else (n_embd, n_embd)
            self.embd_V.append(MaskedConv1D(cv, n_embd, n_embd_ks, stride=1, padding=n_embd_ks // 2, bias=not with_ln))
            self.embd_A.append(MaskedConv1D(ca, n_embd, n_embd_ks, stride=1, padding=n_embd_ks // 2, bias=not with_ln))
            self.embd_norm_V.append(LayerNorm(n_embd) if with_ln else nn.Identity())
            self.embd_norm_A.append(LayerNorm(n_embd) if with_ln else nn.Identity())

        def tb(strides):
            return TransformerBlock(n_embd, n_head, n_ds_strides=strides, attn_pdrop=attn_pdrop,
                                    proj_pdrop=proj_pdrop, path_pdrop=path_pdrop)

        self.self_att_V = nn.ModuleList(tb((1, 1)) for _ in range(arch[1] - 1))
        self.self_att_A = nn.ModuleList(tb((1, 1)) for _ in range(arch[1] - 1))
        # constructed by the reference but never called in forward (:715-749); kept for strict loading
        self.ori_cross_att_Va = tb((1, 1))
        self.ori_cross_att_Av = tb((1, 1))
        self.cross_att_Va = nn.ModuleList(tb((scale_factor, scale_factor)) for _ in range(arch[2]))
        self.cross_att_Av = nn.ModuleList(tb((scale_factor, scale_factor)) for _ in range(arch[2]))
        self.downsample_list = nn.ModuleList(Downsample_pyramid_levels(n_embd, scale_factor) for _ in range(arch[2]))
        self.fusion_module = fusion_module(n_embd, max_len=max_len, n_levels=arch[2] + 1)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv1d)) and m.bias is not None:
                nn.init.constant_(m.bias, 0.0)

    def forward(self, x_V, x_A, mask):
        from .. import _fwd
        return _fwd.backbone_forward(self, x_V, x_A, mask)


class MultiHeadAttention(nn.Module):
    """q/k/v/m Linear projections around a masked softmax attention (:845-924)."""

    def __init__(self, dims, k_dims=None, v_dims=None, h_dims=None, o_dims=None, heads=8, p=0.1, bias=True):
        super().__init__()
        self._q_dims, self._k_dims, self._v_dims = dims, k_dims or dims, v_dims or dims
        self._h_dims, self._o_dims, self._heads = h_dims or dims, o_dims or dims, heads
        self._head_dims = self._h_dims // heads
        self.q = nn.Linear(self._q_dims, self._h_dims, bias=bias)
        self.k = nn.Linear(self._k_dims, self._h_dims, bias=bias)
        self.v = nn.Linear(self._v_dims, self._h_dims, bias=bias)
        self.m = nn.Linear(self._h_dims, self._o_dims, bias=bias)
        self.drop1, self.drop2 = nn.Dropout(p), nn.Dropout(p)
        for lin in (self.q, self.k, self.v, self.m):
            nn.init.xavier_normal_(lin.weight, gain=1.0)
            if lin.bias is not None:
                nn.init.constant_(lin.bias, 0)


class FFN(nn.Module):
    """Linear -> GELU -> Linear (:926-941)."""

    def __init__(self, num_input, p=0.1, ratio=4):
        super().__init__()
        self.fc1 = nn.Linear(num_input, num_input * ratio)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(p)
        self.fc2 = nn.Linear(num_input * ratio, num_input)
        self.drop2 = nn.Dropout(p)


class MultiWayTransformer(nn.Module):
    """Joint attention over [video; audio] tokens, then per-modality FFNs (:943-972)."""

    def __init__(self, num_hidden, dropout_attn=0.1):
        super().__init__()
        self.norm1_fused = nn.LayerNorm(num_hidden)
        self.attn_fusion = MultiHeadAttention(num_hidden, p=dropout_attn)
        self.norm2_video = nn.LayerNorm(num_hidden)
        self.ffn_video = FFN(num_hidden, p=dropout_attn, ratio=4)
        self.norm2_text = nn.LayerNorm(num_hidden)
        self.ffn_text = FFN(num_hidden, p=dropout_attn, ratio=4)


class Alignment(nn.Module):
    """Alignment multiway transformer in front of the backbone (:973-1235).  Inference part only: the
    score / cls heads are kept as parameters (checkpoint contract) but the loss-only tail (:1209-1233) is
    outside the hot path."""

    def __init__(self, video_dim, audio_dim, num_hidden=512, dropout_video=0.0, dropout_Audio=0.0, dropout_fc=0.0,
                 dropout_attn=0.0, num_layers=2, num_classes=100):
        super().__init__()
        self.num_classes, self.num_layers, self.num_hidden = num_classes, num_layers, num_hidden
        self.proj_fc_video = nn.Sequential(nn.Linear(video_dim, num_hidden, bias=True), nn.Dropout(dropout_video))
        self.proj_fc_text = nn.Sequential(nn.Linear(audio_dim, num_hidden, bias=True), nn.Dropout(dropout_Audio))
        self.pos_embed_video = nn.Parameter(torch.zeros(1, 5000, num_hidden))
        self.pos_embed_text = nn.Parameter(torch.zeros(1, 5000, num_hidden))
        self.type_video = nn.Parameter(torch.zeros(1, 1, num_hidden))
        self.type_text = nn.Parameter(torch.zeros(1, 1, num_hidden))
        self.cls_token_video = nn.Parameter(torch.zeros(1, 1, num_hidden))
        self.cls_token_text = nn.Parameter(torch.zeros(1, 1, num_hidden))
        # the SAME layer applied num_layers times (:1009)
        self.multiway_list = nn.ModuleList([MultiWayTransformer(num_hidden, dropout_attn=dropout_attn)] * num_layers)
        self.norm_video = nn.LayerNorm(num_hidden)
        self.norm_text = nn.LayerNorm(num_hidden)
        self.fc_video = nn.Sequential(nn.Linear(num_hidden, num_hidden), nn.ReLU(True), nn.Dropout(dropout_fc),
                                      nn.LayerNorm(num_hidden))
        self.fc_video_score = nn.Conv1d(num_hidden, 1, 1, bias=True)          # loss-only
        self.fc_video_cls = nn.Linear(num_hidden, num_classes)                # loss-only
        self.fc_text = nn.Sequential(nn.Linear(num_hidden, num_hidden), nn.ReLU(True), nn.Dropout(dropout_fc),
                                     nn.LayerNorm(num_hidden))
        self.fc_text_score = nn.Conv1d(num_hidden, 1, 1, bias=True)           # loss-only
        self.fc_text_cls = nn.Linear(num_hidden, num_classes)                 # loss-only
        for p in (self.pos_embed_video, self.pos_embed_text, self.type_video, self.type_text,
                  self.cls_token_video, self.cls_token_text):
            nn.init.trunc_normal_(p, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def forward(self, **kwargs):
        """Reference signature (:1127-1135): lists of one [B,C,T] tensor each; returns
        (new_video_list, new_text_list, contrastive_pairs={})."""
        from .. import _fwd
        return _fwd.alignment_forward(self, **kwargs)
