"""CPU oracle for the PtTransformer inference hot path — TEST INFRASTRUCTURE ONLY.

A functional (state_dict-driven) PyTorch-FP32 restatement of what the reference computes
between the collate dict and the decoded candidates.  Nothing in the product path imports
this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker or the timed CPU baseline.

Parity pin: ``tests/golden/make_golden.py`` runs the *real* reference (imported from
/root/reference in the build container) on seeded inputs/weights and commits its outputs;
``tests/test_oracle_golden.py`` checks this restatement against those fixtures.

Every function cites the reference lines it restates (paths relative to /root/reference).
Layout follows the reference: activations are channels-first ``[B, C, T]``, masks ``[B, 1, T]``.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# Optional operand-rounding hook used by the precision study (tests/precision_study.py):
# applied to both inputs of every dense GEMM/conv to emulate BF16/TF32 tensor-core operands
# with FP32 accumulation.  None = exact FP32 (the oracle proper).
OPERAND_ROUND: Optional[Callable[[torch.Tensor], torch.Tensor]] = None
# Same for the second (B / weight) operand only; None = use OPERAND_ROUND for both operands.
OPERAND_ROUND_B: Optional[Callable[[torch.Tensor], torch.Tensor]] = None


def _r(x):
    return x if OPERAND_ROUND is None else OPERAND_ROUND(x)


def _rb(x):
    if OPERAND_ROUND_B is not None:
        return OPERAND_ROUND_B(x)
    return _r(x)


def _linear(x, w, b=None):
    return F.linear(_r(x), _rb(w), b)


def _conv(x, w, b=None, stride=1, padding=0, groups=1):
    if groups == 1:
        return F.conv1d(_r(x), _rb(w), b, stride=stride, padding=padding)
    return F.conv1d(x, w, b, stride=stride, padding=padding, groups=groups)  # depthwise: FP32 CUDA cores


# ----------------------------------------------------------------------------- blocks.py
def masked_conv1d(sd: SD, p: str, x, mask, stride=1, groups=1):
    """libs/modeling/blocks.py:36-61 — conv, then multiply by the (nearest-downsampled) mask."""
    w = sd[p + ".conv.weight"]
    b = sd.get(p + ".conv.bias")
    k = w.shape[-1]
    out = _conv(x, w, b, stride=stride, padding=k // 2, groups=groups)
    if stride > 1:
        m = F.interpolate(mask.to(x.dtype), size=x.shape[-1] // stride, mode="nearest")
    else:
        m = mask.to(x.dtype)
    if m.shape[-1] != out.shape[-1]:       # blocks.py:56-57: a per-sequence mask [B] (Dependency_Block's co-occurrence branch)
        m = m[:, None, None]
    return out * m, m.bool()


def channel_ln(sd: SD, p: str, x, eps=1e-5):
    """libs/modeling/blocks.py:91-103 — LayerNorm over dim 1 of [B, C, T] (biased variance)."""
    mu = x.mean(dim=1, keepdim=True)
    r = x - mu
    var = (r * r).mean(dim=1, keepdim=True)
    return r / torch.sqrt(var + eps) * sd[p + ".weight"] + sd[p + ".bias"]


def masked_mhca(sd: SD, p: str, x1, x2, mask, n_head: int):
    """libs/modeling/blocks.py:198-245 — x1 -> k, v ; x2 -> q (stride 1 only on this path)."""
    B, C, T = x1.shape
    hs = C // n_head
    q, qm = masked_conv1d(sd, p + ".query_conv", x2, mask, groups=C)
    q = channel_ln(sd, p + ".query_norm", q)
    k, km = masked_conv1d(sd, p + ".key_conv", x1, mask, groups=C)
    k = channel_ln(sd, p + ".key_norm", k)
    v, _ = masked_conv1d(sd, p + ".value_conv", x1, mask, groups=C)
    v = channel_ln(sd, p + ".value_norm", v)
    q = _conv(q, sd[p + ".query.weight"], sd[p + ".query.bias"])
    k = _conv(k, sd[p + ".key.weight"], sd[p + ".key.bias"])
    v = _conv(v, sd[p + ".value.weight"], sd[p + ".value.bias"])
    k = k.view(B, n_head, hs, -1).transpose(2, 3)
    q = q.view(B, n_head, hs, -1).transpose(2, 3)
    v = v.view(B, n_head, hs, -1).transpose(2, 3)
    att = _r(q * (1.0 / math.sqrt(hs))) @ _rb(k).transpose(-2, -1)
    if T == mask.shape[-1]:                # blocks.py:229-236
        att = att.masked_fill(torch.logical_not(km[:, :, None, :]), float("-inf"))
        att = F.softmax(att, dim=-1)
    else:                                  # per-sequence mask: un-masked softmax, then the whole sequence is kept / zeroed
        att = F.softmax(att, dim=-1)
        att = att * km[:, :, :, None].to(att.dtype)
    out = _r(att) @ _rb(v * km[:, :, :, None].to(v.dtype))
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    out = _conv(out, sd[p + ".proj.weight"], sd[p + ".proj.bias"]) * qm.to(out.dtype)
    return out, qm


def transformer_block(sd: SD, p: str, x1, x2, mask, n_head: int):
    """libs/modeling/blocks.py:312-323 (eval: AffineDropPath = per-channel scale, :389-391)."""
    out, om = masked_mhca(sd, p + ".attn", channel_ln(sd, p + ".ln11", x1),
                          channel_ln(sd, p + ".ln12", x2), mask, n_head)
    mf = om.to(out.dtype)
    out = x1 * mf + sd[p + ".drop_path_attn.scale"] * out
    h = channel_ln(sd, p + ".ln2", out)
    h = _conv(h, sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"])
    h = F.gelu(h)
    h = _conv(h, sd[p + ".mlp.3.weight"], sd[p + ".mlp.3.bias"])
    out = out + sd[p + ".drop_path_mlp.scale"] * (h * mf)
    return out, om


def dependency_block(sd: SD, p: str, fpn_feats, fpn_masks, num_classes: int, n_head: int = 1):
    """libs/modeling/dependency_block.py:42-70 — per level: expand to H=128 channels per class, a temporal TransformerBlock
    over [B*C, H, T] and a co-occurrence TransformerBlock over [B*T, H, C] (per-sequence mask), sum, squeeze back."""
    outs = []
    for feats, mask in zip(fpn_feats, fpn_masks):
        fe, mask = masked_conv1d(sd, p + ".feature_expand", feats, mask)
        fe = F.relu(fe).view(feats.shape[0], num_classes, -1, feats.shape[-1]).contiguous()
        B, C, H, T = fe.shape
        tf = fe.view(-1, H, T)
        t_out, _ = transformer_block(sd, p + ".temporal_branch", tf, tf, mask.repeat(C, 1, 1), n_head)
        t_out = t_out.view(B, C, H, T).contiguous()
        cf = fe.transpose(1, 3).contiguous().view(-1, H, C)
        c_out, _ = transformer_block(sd, p + ".cooccur_branch", cf, cf, mask.flatten(), n_head)
        c_out = c_out.view(B, T, H, C).contiguous()
        out = t_out + c_out.transpose(1, 3).contiguous()
        out = out.view(out.shape[0], -1, out.shape[-1])
        out, mask = masked_conv1d(sd, p + ".feature_squeeze", out, mask)
        outs.append(out)
    return outs


# ------------------------------------------------------------- multimodal_backbones.py
def alignment(sd: SD, p: str, visual, audio, mask, num_layers=2, heads=8, return_cls=False):
    """libs/modeling/multimodal_backbones.py:1144-1207 (inference part only).

    visual [B,2048,T], audio [B,128,T], mask [B,1,T] bool -> two [B,512,T] maps.
    The loss-only tail (:1209-1233) is restated in ``oracle/losses_ref.py``; ``return_cls=True`` also returns the
    [CLSV] / [CLST] tokens after the multiway layers ([B,1,512] each, :1189-1190, :1201-1202) that it needs.
    """
    video = visual.transpose(1, 2)
    text = audio.transpose(1, 2)
    m = mask.transpose(1, 2).squeeze(2)
    B, T = m.shape
    video = _linear(video, sd[p + ".proj_fc_video.0.weight"], sd[p + ".proj_fc_video.0.bias"])
    text = _linear(text, sd[p + ".proj_fc_text.0.weight"], sd[p + ".proj_fc_text.0.bias"])
    res_v, res_t = video, text
    video = torch.cat([sd[p + ".cls_token_video"].expand(B, -1, -1), video], dim=1)
    text = torch.cat([sd[p + ".cls_token_text"].expand(B, -1, -1), text], dim=1)
    ones = torch.ones(B, 1, dtype=m.dtype)
    mv = torch.cat([ones, m], dim=1)
    mt = torch.cat([ones, m], dim=1)
    N = T + 1
    video = video + sd[p + ".pos_embed_video"][:, :N, :] + sd[p + ".type_video"]
    text = text + sd[p + ".pos_embed_text"][:, :N, :] + sd[p + ".type_text"]
    # fused mask (:1173-1183): same-modality valid keys, plus the time-aligned token of the
    # other modality (frame_sentence_ratio=1 -> identity band; CLS rows/cols have no cross links)
    mf = torch.zeros(B, 2 * N, 2 * N, dtype=torch.bool)
    mf[:, :N, :N] = mv[:, None, :].expand(B, N, N)
    mf[:, N:, N:] = mt[:, None, :].expand(B, N, N)
    eye = torch.eye(T, dtype=torch.bool)
    mf[:, 1:N, N + 1:2 * N] = eye
    mf[:, N + 1:2 * N, 1:N] = eye
    add_mask = torch.where(mf, 0.0, float("-inf")).repeat_interleave(heads, dim=0)

    fused = torch.cat([video, text], dim=1)
    q0 = p + ".multiway_list.0"          # the same module aliased num_layers times (:1009)
    hd = fused.shape[-1] // heads
    for _ in range(num_layers):
        residual = fused
        x = F.layer_norm(fused, (fused.shape[-1],), sd[q0 + ".norm1_fused.weight"], sd[q0 + ".norm1_fused.bias"])
        # MultiHeadAttention.forward (:891-924)
        q = _linear(x, sd[q0 + ".attn_fusion.q.weight"], sd[q0 + ".attn_fusion.q.bias"]).transpose(0, 1).contiguous()
        k = _linear(x, sd[q0 + ".attn_fusion.k.weight"], sd[q0 + ".attn_fusion.k.bias"]).transpose(0, 1).contiguous()
        v = _linear(x, sd[q0 + ".attn_fusion.v.weight"], sd[q0 + ".attn_fusion.v.bias"]).transpose(0, 1).contiguous()
        b = q.size(1) * heads
        q = q.view(-1, b, hd).transpose(0, 1)
        k = k.view(-1, b, hd).transpose(0, 1)
        v = v.view(-1, b, hd).transpose(0, 1)
        att = torch.bmm(_r(q), _rb(k).transpose(1, 2)) / hd ** 0.5
        att = att + add_mask
        att = att.softmax(-1)
        o = torch.bmm(_r(att), _rb(v)).transpose(0, 1).contiguous()
        o = o.view(o.size(0), -1, heads * hd).transpose(0, 1)
        o = _linear(o, sd[q0 + ".attn_fusion.m.weight"], sd[q0 + ".attn_fusion.m.bias"])
        residual = residual + o
        rv, rt = torch.split(residual, [N, N], dim=1)
        # per-modality FFN (:962-970)
        hv = F.layer_norm(rv, (rv.shape[-1],), sd[q0 + ".norm2_video.weight"], sd[q0 + ".norm2_video.bias"])
        hv = _linear(F.gelu(_linear(hv, sd[q0 + ".ffn_video.fc1.weight"], sd[q0 + ".ffn_video.fc1.bias"])),
                     sd[q0 + ".ffn_video.fc2.weight"], sd[q0 + ".ffn_video.fc2.bias"])
        rv = rv + hv
        ht = F.layer_norm(rt, (rt.shape[-1],), sd[q0 + ".norm2_text.weight"], sd[q0 + ".norm2_text.bias"])
        ht = _linear(F.gelu(_linear(ht, sd[q0 + ".ffn_text.fc1.weight"], sd[q0 + ".ffn_text.fc1.bias"])),
                     sd[q0 + ".ffn_text.fc2.weight"], sd[q0 + ".ffn_text.fc2.bias"])
        rt = rt + ht
        fused = torch.cat([rv, rt], dim=1)
    video, text = rv[:, 1:], rt[:, 1:]
    C = video.shape[-1]
    video = F.layer_norm(res_v + video, (C,), sd[p + ".norm_video.weight"], sd[p + ".norm_video.bias"])
    text = F.layer_norm(res_t + text, (C,), sd[p + ".norm_text.weight"], sd[p + ".norm_text.bias"])
    video = F.layer_norm(F.relu(_linear(video, sd[p + ".fc_video.0.weight"], sd[p + ".fc_video.0.bias"])),
                         (C,), sd[p + ".fc_video.3.weight"], sd[p + ".fc_video.3.bias"])
    text = F.layer_norm(F.relu(_linear(text, sd[p + ".fc_text.0.weight"], sd[p + ".fc_text.0.bias"])),
                        (C,), sd[p + ".fc_text.3.weight"], sd[p + ".fc_text.3.bias"])
    if return_cls:
        return video.transpose(1, 2), text.transpose(1, 2), rv[:, :1], rt[:, :1]
    return video.transpose(1, 2), text.transpose(1, 2)


def maxsig_attn_block(sd: SD, p: str, x, guide, mask, num_heads: int):
    """libs/modeling/multimodal_backbones.py:166-197 (embed_conv is None: embed == in channels)."""
    B, C, H = x.shape
    hc = C // num_heads
    g = _linear(guide, sd[p + ".guide_fc.weight"], sd[p + ".guide_fc.bias"])      # over the guide's time axis
    g = g.reshape(B, -1, num_heads, hc)
    embed = x.reshape(B, num_heads, hc, H)
    aw = torch.einsum("bmch,bnmc->bmhn", _r(embed), _rb(g))
    aw = aw.max(dim=-1)[0]
    aw = aw / (hc ** 0.5)
    aw = aw + sd[p + ".bias"][None, :, None]
    aw = aw.sigmoid()
    y, m = masked_conv1d(sd, p + ".project_conv", x, mask)
    y = y.reshape(B, num_heads, -1, H) * aw.unsqueeze(2)
    return y.reshape(B, -1, H), m


def csp_layer(sd: SD, p: str, x, guide, mask, num_heads: int, num_blocks=3):
    """libs/modeling/multimodal_backbones.py:243-256 (MaxSigmoidCSPLayerWithTwoConv.forward)."""
    xm, mask = masked_conv1d(sd, p + ".main_conv", x, mask)
    mid = xm.shape[1] // 2
    parts = list(xm.split((mid, mid), 1))
    for i in range(num_blocks):
        y, mask = masked_mhca(sd, f"{p}.blocks.{i}", parts[-1], parts[-1], mask, 4)
        parts.append(y)
    y, mask = maxsig_attn_block(sd, p + ".attn_block", parts[-1], guide, mask, num_heads)
    parts.append(y)
    return masked_conv1d(sd, p + ".final_conv", torch.cat(parts, 1), mask)


def fusion_module(sd: SD, p: str, feats: List[torch.Tensor], guide, masks, mask_guide,
                  pool_size=4, num_pool_feats=3):
    """libs/modeling/multimodal_backbones.py:552-619."""
    L = len(feats)
    td_heads = [sd[f"{p}.top_down_layers.{i}.attn_block.bias"].numel() for i in range(L - 1)]
    bu_heads = [sd[f"{p}.bottom_up_layers.{i}.attn_block.bias"].numel() for i in range(L - 1)]
    inner = [feats[-1]]
    for idx in range(L - 1, 0, -1):
        up = F.interpolate(inner[0], scale_factor=2, mode="nearest")
        m_up = masks[idx].repeat_interleave(2, dim=-1)      # the COARSE mask up-sampled (:568-570)
        x = torch.cat([up, feats[idx - 1]], 1)
        out, _ = csp_layer(sd, f"{p}.top_down_layers.{L - 1 - idx}", x, guide, m_up, td_heads[L - 1 - idx])
        inner.insert(0, out)
    pooled = [F.adaptive_avg_pool1d(inner[i], pool_size) for i in range(num_pool_feats)]   # un-masked (:358-365)
    pooled = torch.cat(pooled, dim=-1).transpose(1, 2)                                     # [B, 12, C]
    q = _conv(pooled, sd[p + ".match_projection.weight"], sd[p + ".match_projection.bias"]).transpose(1, 2)
    guide, _ = masked_mhca(sd, p + ".text_enhancer", guide, q, mask_guide, 4)               # (:600)
    outs = [inner[0]]
    for idx in range(L - 1):
        d, dm = masked_conv1d(sd, f"{p}.downsample_layers.{idx}.down_conv", outs[-1], masks[idx], stride=2)
        d = F.silu(channel_ln(sd, f"{p}.downsample_layers.{idx}.down_norm", d))
        out, _ = csp_layer(sd, f"{p}.bottom_up_layers.{idx}", torch.cat([d, inner[idx + 1]], 1), guide, dm, bu_heads[idx])
        outs.append(out)
    return outs


def backbone(sd: SD, p: str, x_v, x_a, mask, n_head=4, n_embd_layers=2, n_stem=2, n_levels=6):
    """libs/modeling/multimodal_backbones.py:771-841 (eval branch, use_abs_pe=True)."""
    T = x_v.shape[-1]
    mv = ma = mask
    for i in range(n_embd_layers):
        x_v, mv = masked_conv1d(sd, f"{p}.embd_V.{i}", x_v, mv)
        x_v = F.gelu(channel_ln(sd, f"{p}.embd_norm_V.{i}", x_v))
        x_a, ma = masked_conv1d(sd, f"{p}.embd_A.{i}", x_a, ma)
        x_a = F.gelu(channel_ln(sd, f"{p}.embd_norm_A.{i}", x_a))
    pe = sd[p + ".pos_embd"]
    if T >= pe.shape[-1]:
        pe = F.interpolate(pe, T, mode="linear", align_corners=False)
    x_v = x_v + pe[:, :, :T] * mv.to(x_v.dtype)
    x_a = x_a + pe[:, :, :T] * ma.to(x_a.dtype)
    for i in range(n_stem):
        x_v, mv = transformer_block(sd, f"{p}.self_att_V.{i}", x_v, x_v, mv, n_head)
        x_a, ma = transformer_block(sd, f"{p}.self_att_A.{i}", x_a, x_a, ma, n_head)

    def pyramid(x, m):
        xs, ms = [x], [m]
        for i in range(n_levels - 1):
            y, m2 = masked_conv1d(sd, f"{p}.downsample_list.{i}.down_conv", xs[-1], ms[-1], stride=2,
                                  groups=xs[-1].shape[1])
            xs.append(channel_ln(sd, f"{p}.downsample_list.{i}.down_norm", y))
            ms.append(m2)
        return xs, ms

    xs_v, ms_v = pyramid(x_v, mv)
    feats_v = fusion_module(sd, p + ".fusion_module", xs_v, x_a, ms_v, ma)
    xs_a, ms_a = pyramid(x_a, ma)
    feats_a = fusion_module(sd, p + ".fusion_module", xs_a, x_v, ms_a, mv)
    return feats_v, feats_a, ms_v


def sinusoid_pos_embd(n_position: int, d_hid: int) -> torch.Tensor:
    """libs/modeling/blocks.py:106-117 scaled by 1/sqrt(d) (multimodal_backbones.py:657): [1, C, T]."""
    import numpy as np
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0).transpose(1, 2) / (d_hid ** 0.5)


# ------------------------------------------------------------- multimodal_meta_archs.py
def heads(sd: SD, feats_av: List[torch.Tensor], masks, num_classes=100):
    """PtTransformerClsHead.forward (:166-178) and PtTransformerRegHead.forward (:245-259)."""
    logits, offsets = [], []
    for l, (x, m) in enumerate(zip(feats_av, masks)):
        c = r = x
        for i in range(2):
            c, _ = masked_conv1d(sd, f"cls_head.head.{i}", c, m)
            c = F.relu(channel_ln(sd, f"cls_head.norm.{i}", c))
            r, _ = masked_conv1d(sd, f"reg_head.head.{i}", r, m)
            r = F.relu(channel_ln(sd, f"reg_head.norm.{i}", r))
        c, _ = masked_conv1d(sd, "cls_head.cls_head", c, m)
        r, _ = masked_conv1d(sd, "reg_head.offset_head", r, m)
        r = F.relu(r * sd[f"reg_head.scale.{l}.scale"])
        logits.append(c.permute(0, 2, 1))                                              # [B, T_l, ncls]
        r = r.permute(0, 2, 1)
        offsets.append(r.reshape(r.shape[0], r.shape[1], num_classes, -1).contiguous())  # [B, T_l, ncls, 2]
    return logits, offsets


def forward_logits(sd: SD, visual, audio, mask, pos_embd=None, use_dependency=False):
    """PtTransformer.forward up to the permuted head outputs (:426-493), losses skipped.

    Returns (logits per level [B,T_l,ncls], offsets per level [B,T_l,ncls,2], masks per level [B,T_l]).
    """
    sd = dict(sd)
    if "backbone.pos_embd" not in sd:          # non-persistent buffer (multimodal_backbones.py:658)
        C = sd["backbone.embd_V.0.conv.weight"].shape[0]
        sd["backbone.pos_embd"] = pos_embd if pos_embd is not None else sinusoid_pos_embd(visual.shape[-1], C)
    v, a = alignment(sd, "alignment", visual, audio, mask)
    fv, fa, ms = backbone(sd, "backbone", v, a, mask)
    feats = [torch.cat((x, y), 1) for x, y in zip(fv, fa)]
    if use_dependency:                         # multimodal_meta_archs.py:474-475
        feats = dependency_block(sd, "dependency_block", feats, ms, sd["cls_head.cls_head.conv.bias"].shape[0])
    logits, offsets = heads(sd, feats, ms)
    return logits, offsets, [m.squeeze(1) for m in ms]


def make_points(T: int, n_levels=6, scale_factor=2,
                regression_range=((0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000))):
    """libs/datasets/loc_generators.py:61-79 — per level [T_l, 4] = (t, reg_lo, reg_hi, stride)."""
    pts = []
    for l in range(n_levels):
        s = scale_factor ** l
        t = torch.arange(0, T, s, dtype=torch.float32)[:, None]
        rr = torch.tensor(regression_range[l], dtype=torch.float32)[None].repeat(t.shape[0], 1)
        st = torch.full((t.shape[0], 1), float(s))
        pts.append(torch.cat((t, rr, st), dim=1))
    return pts


def decode_single_video(points, masks, logits, offsets, pre_nms_thresh=0.001, pre_nms_topk=2000,
                        duration_thresh=0.05, num_classes=100, stable=True):
    """PtTransformer.inference_single_video (:745-817) for one video.

    ``stable=True`` pins the reference's unspecified tie order (``sort`` without ``stable``) to the
    canonical one of SURVEY.md §8a: score descending, flat index ascending.
    Returns segs [N,2], scores [N], labels [N] (int64), flat ids [N] (level*2^20 + flat index).
    """
    segs_all, scores_all, cls_all, ids_all = [], [], [], []
    for l, (cls_i, off_i, pts_i, m_i) in enumerate(zip(logits, offsets, points, masks)):
        prob = (cls_i.sigmoid() * m_i.unsqueeze(-1)).flatten()
        keep1 = prob > pre_nms_thresh
        prob = prob[keep1]
        idx = keep1.nonzero(as_tuple=True)[0]
        k = min(pre_nms_topk, idx.size(0))
        prob, order = prob.sort(descending=True, stable=stable)
        prob = prob[:k].clone()
        idx = idx[order[:k]].clone()
        pt = torch.div(idx, num_classes, rounding_mode="floor")
        cls = torch.fmod(idx, num_classes)
        off = off_i.view(-1, off_i.shape[-1])[idx]
        pts = pts_i[pt]
        left = pts[:, 0] - off[:, 0] * pts[:, 3]
        right = pts[:, 0] + off[:, 1] * pts[:, 3]
        keep2 = (right - left) > duration_thresh
        segs_all.append(torch.stack((left, right), -1)[keep2])
        scores_all.append(prob[keep2])
        cls_all.append(cls[keep2])
        ids_all.append(idx[keep2] + (l << 20))
    return torch.cat(segs_all), torch.cat(scores_all), torch.cat(cls_all), torch.cat(ids_all)


def to_seconds(segs, feat_stride, feat_num_frames, fps, duration):
    """PtTransformer.postprocessing (:852-856)."""
    if segs.shape[0] > 0:
        segs = (segs * feat_stride + 0.5 * feat_num_frames) / fps
        segs[segs <= 0.0] *= 0.0
        segs[segs >= duration] = segs[segs >= duration] * 0.0 + duration
    return segs
