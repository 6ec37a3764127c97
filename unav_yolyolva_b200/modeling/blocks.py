"""Building blocks of the hot path with the reference's names, constructor signatures and parameter
names (/root/reference/libs/modeling/blocks.py), so that ``load_state_dict(strict=True)`` accepts
reference checkpoints.

Modules are parameter holders plus a ``forward`` that keeps the reference's calling convention
(channels-first ``[B, C, T]`` FP32 tensors and ``[B, 1, T]`` bool masks) but runs on the hand-written
sm_100a kernels through the C ABI (``.._fwd``).  There is no PyTorch-eager fallback: inputs must be CUDA
tensors on an sm_100 device.  ``PtTransformer.forward`` does not chain these module forwards — it runs
the fused token-major engine (``..engine``) over the same parameters.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn


class MaskedConv1D(nn.Module):
    """Conv1d (odd kernel, "same" zero padding) whose output is multiplied by the (strided) mask.

    Reference: blocks.py:8-61.  Parameters: ``conv.weight [Cout, Cin/groups, k]``, ``conv.bias [Cout]``.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode="zeros"):
        super().__init__()
        assert kernel_size % 2 == 1 and kernel_size // 2 == padding
        self.stride = stride
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                              padding_mode)
        if bias:
            nn.init.constant_(self.conv.bias, 0.0)

    def forward(self, x, mask):
        from .. import _fwd
        return _fwd.masked_conv1d(self, x, mask)


class LayerNorm(nn.Module):
    """LayerNorm over the channel axis of [B, C, T] (blocks.py:64-103); ``weight``/``bias`` are [1, C, 1]."""

    def __init__(self, num_channels, eps=1e-5, affine=True, device=None, dtype=None):
        super().__init__()
        kw = {"device": device, "dtype": dtype}
        self.num_channels, self.eps, self.affine = num_channels, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones([1, num_channels, 1], **kw))
            self.bias = nn.Parameter(torch.zeros([1, num_channels, 1], **kw))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        from .. import _fwd
        return _fwd.channel_layernorm(self, x)


def get_sinusoid_encoding(n_position, d_hid):
    """Sinusoid position table [1, d_hid, n_position] (blocks.py:106-117)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0).transpose(1, 2)


class MaskedMHCA(nn.Module):
    """Multi-head conv attention with mask (blocks.py:119-245): depthwise k=3 conv + LayerNorm + 1x1
    projection for each of q (from x2), k, v (from x1), masked softmax attention, output projection."""

    def __init__(self, n_embd, n_head, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_embd, self.n_head = n_embd, n_head
        self.n_channels = n_embd // n_head
        self.scale = 1.0 / math.sqrt(self.n_channels)
        assert n_qx_stride == 1 or n_qx_stride % 2 == 0
        assert n_kv_stride == 1 or n_kv_stride % 2 == 0
        self.n_qx_stride, self.n_kv_stride = n_qx_stride, n_kv_stride

        def dw(stride_for_kernel, stride):
            k = stride_for_kernel + 1 if stride_for_kernel > 1 else 3
            return MaskedConv1D(n_embd, n_embd, k, stride=stride, padding=k // 2, groups=n_embd, bias=False)

        # the reference strides the query conv by n_kv_stride as well (blocks.py:159-165)
        self.query_conv = dw(n_qx_stride, n_kv_stride)
        self.query_norm = LayerNorm(n_embd)
        self.key_conv = dw(n_kv_stride, n_kv_stride)
        self.key_norm = LayerNorm(n_embd)
        self.value_conv = dw(n_kv_stride, n_kv_stride)
        self.value_norm = LayerNorm(n_embd)
        self.key = nn.Conv1d(n_embd, n_embd, 1)
        self.query = nn.Conv1d(n_embd, n_embd, 1)
        self.value = nn.Conv1d(n_embd, n_embd, 1)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.proj_drop = nn.Dropout(proj_pdrop)
        self.proj = nn.Conv1d(n_embd, n_embd, 1)

    def forward(self, x1, x2, mask):
        from .. import _fwd
        return _fwd.masked_mhca(self, x1, x2, mask)


class TransformerBlock(nn.Module):
    """Pre-LN transformer block on [B, C, T] (blocks.py:247-323)."""

    def __init__(self, n_embd, n_head, n_ds_strides=(1, 1), n_out=None, n_hidden=None, act_layer=nn.GELU,
                 attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0):
        super().__init__()
        assert len(n_ds_strides) == 2
        self.ln11 = LayerNorm(n_embd)
        self.ln12 = LayerNorm(n_embd)
        self.ln2 = LayerNorm(n_embd)
        self.n_embd = n_embd
        self.attn = MaskedMHCA(n_embd, n_head, n_qx_stride=n_ds_strides[0], n_kv_stride=n_ds_strides[1],
                               attn_pdrop=attn_pdrop, proj_pdrop=proj_pdrop)
        if n_ds_strides[0] > 1:
            k, s = n_ds_strides[0] + 1, n_ds_strides[0]
            self.pool_skip = nn.MaxPool1d(k, stride=s, padding=(n_ds_strides[0] + 1) // 2)
        else:
            self.pool_skip = nn.Identity()
        n_hidden = 4 * n_embd if n_hidden is None else n_hidden
        n_out = n_embd if n_out is None else n_out
        self.mlp = nn.Sequential(
            nn.Conv1d(n_embd, n_hidden, 1), act_layer(), nn.Dropout(proj_pdrop, inplace=True),
            nn.Conv1d(n_hidden, n_out, 1), nn.Dropout(proj_pdrop, inplace=True))
        if path_pdrop > 0.0:
            self.drop_path_attn = AffineDropPath(n_embd, drop_prob=path_pdrop)
            self.drop_path_mlp = AffineDropPath(n_out, drop_prob=path_pdrop)
        else:
            self.drop_path_attn = nn.Identity()
            self.drop_path_mlp = nn.Identity()

    def forward(self, x1, x2, mask, pos_embd=None):
        from .. import _fwd
        return _fwd.transformer_block(self, x1, x2, mask, pos_embd)


class Scale(nn.Module):
    """Learnable scalar multiplier (blocks.py:326-344)."""

    def __init__(self, init_value=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(init_value, dtype=torch.float32), requires_grad=True)

    def forward(self, x):
        return x * self.scale


class AffineDropPath(nn.Module):
    """Per-channel scale (+ stochastic depth in training only) on the residual branch (blocks.py:375-391).
    Inference: ``scale * x``."""

    def __init__(self, num_dim, drop_prob=0.0, init_scale_value=1e-4):
        super().__init__()
        self.scale = nn.Parameter(init_scale_value * torch.ones((1, num_dim, 1)), requires_grad=True)
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob > 0.0:
            raise NotImplementedError("training (stochastic depth) is outside the inference hot path")
        return self.scale * x
