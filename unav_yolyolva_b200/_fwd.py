"""Module-level forwards with the reference's calling convention (channels-first ``[B, C, T]`` FP32 tensors,
``[B, 1, T]`` bool masks) on the sm_100a kernels.

These back ``MaskedConv1D.forward``, ``LayerNorm.forward``, ``MaskedMHCA.forward`` ... so that the sub-modules of
the hot path are usable (and testable) one by one like the reference's.  Arithmetic runs in the C-ABI kernels
(GEMM, dwconv+LN, LayerNorm, attention, MaxSigmoid gate, pool+match); torch is used only to allocate buffers and for
pure data movement of the convenience path (``cat`` / ``split`` / nearest up-sampling).  ``PtTransformer.forward`` does
NOT go through here: it runs the fused token-major engine (``engine.py``), which avoids all layout conversions.

Precision follows ``MODE`` ("bf16x3" tensor-core split by default, "fp32", "bf16").
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import kernels as K
from .engine import MODES

MODE = "bf16x3"


def _need_cuda(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise RuntimeError("the B200 hot path has no CPU fallback: expected a CUDA tensor")


def _op():
    return MODES[MODE]


def _packed(param: torch.Tensor, kind: str) -> torch.Tensor:
    """Packed weight operand, cached per parameter (storage, version, mode): lin = [N,K(,1)], conv3 = [N,Cin,3] -> [N,3Cin]."""
    op, _ = _op()
    # the cache lives on the parameter object itself (an address-keyed table would alias freed parameters)
    cache = param.__dict__.setdefault("_unav_packed", {})
    key = (param.data_ptr(), param._version, kind, op)
    w = cache.get(key)
    if w is None:
        cache.clear()
        t = param.detach().float()
        if kind == "conv3":
            t = t.permute(0, 2, 1).reshape(t.shape[0], -1).contiguous()
        else:
            t = t.reshape(t.shape[0], -1)
        w = K.pack_operand(t, op)
        cache[key] = w
    return w


def _vec(p):
    return None if p is None else p.detach().float().reshape(-1).contiguous()


# ------------------------------------------------------------------------------ layout helpers
def to_rows(x: torch.Tensor) -> torch.Tensor:
    """[B, C, T] -> token-major [B*T, C] FP32."""
    B, C, T = x.shape
    out = torch.empty(B * T, C, dtype=torch.float32, device=x.device)
    K.transpose_cast(x.contiguous().float(), T, out, B, C, T, K.F32)
    return out


def from_rows(y: torch.Tensor, B: int, T: int) -> torch.Tensor:
    """[B*T, C] -> [B, C, T]."""
    C = y.shape[1]
    out = torch.empty(B * C, T, dtype=torch.float32, device=y.device)
    K.transpose_cast(y, y.stride(0), out, B, T, C, K.F32)
    return out.view(B, C, T)


def mask_rows(mask: torch.Tensor, stride: int = 1) -> torch.Tensor:
    m = mask.reshape(mask.shape[0], -1)
    if stride > 1:
        m = m[:, ::stride]
    return m.to(torch.uint8).contiguous().reshape(-1)


def _operand(x_rows, B, T, C, ntaps=1, stride=1):
    """FP32 rows -> GEMM operand (cast, or k=3 im2col with per-video zero padding, optional stride)."""
    op, _ = _op()
    To = T // stride
    buf = K.new_operand(B * To, ntaps * C, op, x_rows.device)
    K.rowcopy([{"src": x_rows, "dst": buf, "nseg": B, "seg_len_in": T, "seg_len_out": To, "num": stride, "den": 1,
                "ntaps": ntaps, "tap_stride": C, "C": C}], op)
    return buf


def _gemm(A, W, M, N, Kd, **epi):
    op, backend = _op()
    out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    act = epi.pop("act", K.ACT_NONE)
    res_masked = epi.pop("res_masked", False)
    K.gemm([dict(A=A, W=W, out_f32=out, **epi)], M, N, Kd, op, act, res_masked, backend)
    return out


# ------------------------------------------------------------------------------ rows-space primitives
def conv_rows(conv: torch.nn.Conv1d, x_rows, B, T, mask_out_u8, stride=1, **epi):
    Cout, Cin_g, k = conv.weight.shape
    if conv.groups == 1:
        W = _packed(conv.weight, "conv3" if k == 3 else "lin")
        A = _operand(x_rows, B, T, Cin_g, ntaps=k, stride=stride)
        return _gemm(A, W, B * (T // stride), Cout, k * Cin_g, bias=_vec(conv.bias), rowmask=mask_out_u8, **epi)
    assert conv.groups == Cout == x_rows.shape[1] and k == 3 and conv.bias is None, "unsupported grouped conv"
    out = torch.empty(B * (T // stride), Cout, dtype=torch.float32, device=x_rows.device)
    K.dwconv_ln([{"x": x_rows, "mask_out": mask_out_u8,
                  "outs": [{"dw": _vec(conv.weight), "ln_w": None, "ln_b": None, "out_f32": out}]}], B, T, stride, Cout, K.F32)
    return out


def ln_rows(norm, x_rows, act=K.ACT_NONE):
    M, C = x_rows.shape
    out = torch.empty(M, C, dtype=torch.float32, device=x_rows.device)
    K.layernorm_rows([{"x": x_rows, "w": _vec(norm.weight), "b": _vec(norm.bias), "out_f32": out}], M, C, K.F32, act=act,
                     eps=norm.eps)
    return out


def mhca_rows(mod, x1_rows, x2_rows, m_u8, B, T, pre=None, key_mask=True, **proj_epi):
    """MaskedMHCA (blocks.py:198-245) on rows.  pre = ((ln11_w, ln11_b), (ln12_w, ln12_b)) fuses the TransformerBlock
    input LayerNorms (blocks.py:314).  key_mask=False: the per-sequence-mask branch (blocks.py:233-236) — every key of a
    kept sequence takes part in the softmax; masked sequences are zeroed by the row masks."""
    assert mod.n_qx_stride == 1 and mod.n_kv_stride == 1, "strided MaskedMHCA is not on the hot path"
    op, backend = _op()
    C, nh = mod.n_embd, mod.n_head
    dev = x1_rows.device
    M = B * T
    qi, ki, vi = (K.new_operand(M, C, op, dev) for _ in range(3))

    def out(name, buf, src):
        return {"dw": _vec(getattr(mod, name + "_conv").conv.weight), "ln_w": _vec(getattr(mod, name + "_norm").weight),
                "ln_b": _vec(getattr(mod, name + "_norm").bias), "src": src, "out_op": buf}

    same = x1_rows.data_ptr() == x2_rows.data_ptr()
    if same:
        g = {"x": x1_rows, "mask_out": m_u8, "outs": [out("query", qi, 1 if pre else -1), out("key", ki, 0 if pre else -1),
                                                       out("value", vi, 0 if pre else -1)]}
        if pre:
            g["pre"] = [pre[0], pre[1]]
        K.dwconv_ln([g], B, T, 1, C, op)
    else:
        gkv = {"x": x1_rows, "mask_out": m_u8, "outs": [out("key", ki, 0 if pre else -1), out("value", vi, 0 if pre else -1)]}
        gq = {"x": x2_rows, "mask_out": m_u8, "outs": [out("query", qi, 0 if pre else -1)]}
        if pre:
            gkv["pre"], gq["pre"] = [pre[0]], [pre[1]]
        K.dwconv_ln([gkv], B, T, 1, C, op)
        K.dwconv_ln([gq], B, T, 1, C, op)
    ao = K.new_operand(M, C, op, dev)
    hs = C // nh
    km = m_u8 if key_mask else None
    lins = (mod.query, mod.key, mod.value)
    if backend == K.GEMM_TCGEN05 and T <= 256 and hs in (64, 128):
        # tensor-core attention: q, k as operand rows, values written transposed per item by the GEMM epilogue
        qo, ko = K.new_operand(M, C, op, dev), K.new_operand(M, C, op, dev)
        vt = K.new_operand(B * C, T, op, dev)
        outs = ({"out_op": qo}, {"out_op": ko}, {"out_opT": vt, "t_seg": T})
        K.gemm([dict({"A": a, "W": _packed(lin.weight, "lin"), "bias": _vec(lin.bias)}, **o)
                for a, lin, o in zip((qi, ki, vi), lins, outs)], M, C, C, op, K.ACT_NONE, False, backend)
        K.attention_tc([{"q": qo, "k": ko, "vt": vt, "kmask": km, "out": ao}], B, T, T, nh, hs, 1.0 / math.sqrt(hs), op)
    else:
        qp, kp, vp = (torch.empty(M, C, dtype=torch.float32, device=dev) for _ in range(3))
        K.gemm([{"A": a, "W": _packed(lin.weight, "lin"), "bias": _vec(lin.bias), "out_f32": o}
                for a, lin, o in zip((qi, ki, vi), lins, (qp, kp, vp))], M, C, C, op, K.ACT_NONE, False, backend)
        K.attention([{"q": qp, "k": kp, "v": vp, "kmask": km, "out": ao}], B, T, T, nh, hs, 1.0 / math.sqrt(hs), op)
    return _gemm(ao, _packed(mod.proj.weight, "lin"), M, C, C, bias=_vec(mod.proj.bias), rowmask=m_u8, **proj_epi)


# ------------------------------------------------------------------------------ blocks.py
def masked_conv1d(mod, x, mask):
    """blocks.py:36-61."""
    _need_cuda(x)
    B, C, T = x.shape
    assert T % mod.stride == 0
    m_out = mask_rows(mask, mod.stride)
    y = conv_rows(mod.conv, to_rows(x), B, T, m_out, stride=mod.stride)
    return from_rows(y, B, T // mod.stride), m_out.view(B, 1, -1).bool()


def channel_layernorm(mod, x):
    """blocks.py:91-103."""
    _need_cuda(x)
    assert x.shape[1] == mod.num_channels and mod.affine
    B, C, T = x.shape
    return from_rows(ln_rows(mod, to_rows(x)), B, T)


def masked_mhca(mod, x1, x2, mask):
    """blocks.py:198-245."""
    _need_cuda(x1)
    B, C, T = x1.shape
    r1 = to_rows(x1)
    r2 = r1 if x2 is x1 else to_rows(x2)
    y = mhca_rows(mod, r1, r2, mask_rows(mask), B, T)
    return from_rows(y, B, T), mask.bool()


def transformer_block(mod, x1, x2, mask, pos_embd=None):
    """blocks.py:312-323 (stride-1 blocks; eval-mode AffineDropPath = per-channel scale)."""
    _need_cuda(x1)
    assert pos_embd is None and isinstance(mod.pool_skip, torch.nn.Identity)
    op, backend = _op()
    B, C, T = x1.shape
    M = B * T
    seq_mask = mask.shape[-1] != T            # a per-sequence mask [B] (blocks.py:56-57, :233-236): Dependency_Block
    m = (mask.reshape(B, 1).expand(B, T).to(torch.uint8).contiguous().reshape(-1) if seq_mask else mask_rows(mask))
    r1 = to_rows(x1)
    r2 = r1 if x2 is x1 else to_rows(x2)
    sa = _vec(mod.drop_path_attn.scale) if hasattr(mod.drop_path_attn, "scale") else None
    sm = _vec(mod.drop_path_mlp.scale) if hasattr(mod.drop_path_mlp, "scale") else None
    pre = ((_vec(mod.ln11.weight), _vec(mod.ln11.bias)), (_vec(mod.ln12.weight), _vec(mod.ln12.bias)))
    o1 = mhca_rows(mod.attn, r1, r2, m, B, T, pre=pre, key_mask=not seq_mask, res=r1, colscale=sa, res_masked=True)
    hn = K.new_operand(M, C, op, x1.device)
    K.layernorm_rows([{"x": o1, "w": _vec(mod.ln2.weight), "b": _vec(mod.ln2.bias), "out_op": hn}], M, C, op)
    fc1, fc2 = mod.mlp[0], mod.mlp[3]
    H = fc1.weight.shape[0]
    hm = K.new_operand(M, H, op, x1.device)
    K.gemm([{"A": hn, "W": _packed(fc1.weight, "lin"), "bias": _vec(fc1.bias), "out_op": hm}], M, H, C, op, K.ACT_GELU, False, backend)
    out = _gemm(hm, _packed(fc2.weight, "lin"), M, fc2.weight.shape[0], H, bias=_vec(fc2.bias), rowmask=m, res=o1, colscale=sm)
    return from_rows(out, B, T), (mask.reshape(B, 1, 1) if seq_mask else mask).bool()


def dependency_block_forward(mod, fpn_feats, fpn_masks):
    """dependency_block.py:42-70.  Per pyramid level: feature_expand (k=3 conv 2C -> 128*num_classes, ReLU fused in the
    GEMM epilogue: relu(conv*mask) = relu(conv)*mask), a temporal TransformerBlock over [B*classes, 128, T], a
    co-occurrence TransformerBlock over [B*T, 128, classes] with the per-sequence mask, sum, feature_squeeze.  The
    reshapes / transposes between the three views are the reference's own (torch views and copies); the mask of the
    temporal branch is `mask.repeat(C, 1, 1)` exactly as the reference builds it (:51)."""
    assert len(fpn_feats) == len(fpn_masks)
    outs = []
    for feats, mask in zip(fpn_feats, fpn_masks):
        _need_cuda(feats)
        B, Cin, T = feats.shape
        C = mod.num_classes
        m_u8 = mask_rows(mask)
        fe_rows = conv_rows(mod.feature_expand.conv, to_rows(feats), B, T, m_u8, act=K.ACT_RELU)     # [B*T, C*H]
        H = fe_rows.shape[1] // C
        fe = from_rows(fe_rows, B, T).view(B, C, H, T)
        tf = fe.view(B * C, H, T)
        t_out, _ = transformer_block(mod.temporal_branch, tf, tf, mask.repeat(C, 1, 1))
        cf = fe.transpose(1, 3).contiguous().view(B * T, H, C)
        c_out, _ = transformer_block(mod.cooccur_branch, cf, cf, mask.flatten())
        out = t_out.view(B, C, H, T) + c_out.view(B, T, H, C).transpose(1, 3)
        out = out.reshape(B, C * H, T)
        sq = conv_rows(mod.feature_squeeze.conv, to_rows(out), B, T, m_u8)
        outs.append(from_rows(sq, B, T))
    return tuple(outs), fpn_masks


# ------------------------------------------------------------------------------ multimodal_backbones.py
def pyramid_downsample(mod, x, mask):
    """multimodal_backbones.py:44-48: depthwise k=3 stride-s conv * mask -> LayerNorm (one fused kernel)."""
    _need_cuda(x)
    B, C, T = x.shape
    s = mod.x_stride
    m_out = mask_rows(mask, s)
    out = torch.empty(B * (T // s), C, dtype=torch.float32, device=x.device)
    K.dwconv_ln([{"x": to_rows(x), "mask_out": m_out,
                  "outs": [{"dw": _vec(mod.down_conv.conv.weight), "ln_w": _vec(mod.down_norm.weight),
                            "ln_b": _vec(mod.down_norm.bias), "out_f32": out}]}], B, T, s, C, K.F32)
    return from_rows(out, B, T // s), m_out.view(B, 1, -1).bool()


def fusion_downsample(mod, x, mask):
    """multimodal_backbones.py:352-356: dense k=3 stride-2 conv (+bias) * mask -> LayerNorm -> SiLU."""
    _need_cuda(x)
    B, C, T = x.shape
    s = mod.x_stride
    m_out = mask_rows(mask, s)
    y = conv_rows(mod.down_conv.conv, to_rows(x), B, T, m_out, stride=s)
    return from_rows(ln_rows(mod.down_norm, y, act=K.ACT_SILU), B, T // s), m_out.view(B, 1, -1).bool()


def maxsig_attn_block(mod, x, guide, mask):
    """multimodal_backbones.py:166-197; guide [B, n_words, guide_channels] (channels-first stem output: n_words = C)."""
    _need_cuda(x)
    op, backend = _op()
    B, C, T = x.shape
    H, hc = mod.num_heads, mod.head_channels
    nw, gc = guide.shape[1], guide.shape[2]
    assert gc % 4 == 0, "guide length must be a multiple of 4"
    g_op = K.new_operand(B * nw, gc, op, x.device)
    K.rowcopy([{"src": guide.reshape(B * nw, gc).contiguous().float(), "dst": g_op, "nseg": 1, "seg_len_in": B * nw,
                "seg_len_out": B * nw, "C": gc}], op)
    G = _gemm(g_op, _packed(mod.guide_fc.weight, "lin"), B * nw, mod.guide_fc.weight.shape[0], gc, bias=_vec(mod.guide_fc.bias))
    xr = to_rows(x)
    m = mask_rows(mask)
    gate = torch.empty(B * T, H, dtype=torch.float32, device=x.device)
    K.maxsig_gate(xr, G, _vec(mod.bias), gate, B, T, nw, H, hc)
    y = conv_rows(mod.project_conv.conv, xr, B, T, m, gate=gate, gate_groups=H, gate_width=mod.project_conv.conv.weight.shape[0] // H)
    return from_rows(y, B, T), mask.bool()


def csp_layer(mod, x, guide, mask):
    """multimodal_backbones.py:243-256."""
    xm, mask = mod.main_conv(x, mask)
    parts = list(xm.split((mod.mid_channels, mod.mid_channels), 1))
    for blk in mod.blocks:
        y, mask = blk(parts[-1].contiguous(), parts[-1].contiguous(), mask)
        parts.append(y)
    y, mask = mod.attn_block(parts[-1], guide, mask)
    parts.append(y)
    return mod.final_conv(torch.cat(parts, 1), mask)


def fusion_forward(mod, img_feats, txt_feats, mask_img, mask_txt):
    """multimodal_backbones.py:552-619."""
    L = len(mod.in_channels)
    assert len(img_feats) == L
    inner = [img_feats[-1]]
    for idx in range(L - 1, 0, -1):
        up = inner[0].repeat_interleave(2, dim=-1)                       # nearest x2 (:565-566)
        m_up = mask_img[idx].repeat_interleave(2, dim=-1)                # coarse mask up-sampled (:568-570)
        out, _ = mod.top_down_layers[L - 1 - idx](torch.cat([up, img_feats[idx - 1]], 1), txt_feats, m_up)
        inner.insert(0, out)
    B, C, T0 = inner[0].shape
    Tq = mod.match_projection.weight.shape[0]
    rows = [to_rows(inner[i]) for i in range(mod.num_feats)]
    q = torch.empty(B * Tq, C, dtype=torch.float32, device=rows[0].device)
    wm = mod.match_projection.weight.detach().float().reshape(Tq, -1).contiguous()
    K.pool_match(rows[0], rows[1], rows[2], inner[0].shape[-1], inner[1].shape[-1], inner[2].shape[-1], wm,
                 _vec(mod.match_projection.bias), q, B, C, Tq, mod.pool_size)
    txt_feats, mask_txt = mod.text_enhancer(txt_feats, from_rows(q, B, Tq), mask_txt)
    outs = [inner[0]]
    for idx in range(L - 1):
        d, dm = mod.downsample_layers[idx](outs[-1], mask_img[idx])
        out, _ = mod.bottom_up_layers[idx](torch.cat([d, inner[idx + 1]], 1), txt_feats, dm)
        outs.append(out)
    return tuple(outs), txt_feats, mask_img, mask_txt


def backbone_forward(mod, x_V, x_A, mask):
    """multimodal_backbones.py:771-841 (eval branch)."""
    _need_cuda(x_V)
    B, C, T = x_V.shape
    mV = mA = mask
    for i in range(len(mod.embd_V)):
        x_V, mV = mod.embd_V[i](x_V, mV)
        x_A, mA = mod.embd_A[i](x_A, mA)
        if isinstance(mod.embd_norm_V[i], torch.nn.Identity):
            raise NotImplementedError("embd_with_ln=False is not on the hot path")
        x_V = from_rows(ln_rows(mod.embd_norm_V[i], to_rows(x_V), act=K.ACT_GELU), B, T)
        x_A = from_rows(ln_rows(mod.embd_norm_A[i], to_rows(x_A), act=K.ACT_GELU), B, T)
    if mod.use_abs_pe:
        pe = mod.pos_embd
        if T > pe.shape[-1]:
            raise NotImplementedError("T > max_len (position-embedding interpolation) is not on the hot path")
        x_V = x_V + pe[:, :, :T] * mV.to(x_V.dtype)
        x_A = x_A + pe[:, :, :T] * mA.to(x_A.dtype)
    for i in range(len(mod.self_att_V)):
        x_V, mV = mod.self_att_V[i](x_V, x_V, mV)
        x_A, mA = mod.self_att_A[i](x_A, x_A, mA)

    def pyramid(x, m):
        xs, ms = [x], [m]
        for d in mod.downsample_list:
            y, m2 = d(xs[-1], ms[-1])
            xs.append(y); ms.append(m2)
        return xs, ms

    xs_v, ms_v = pyramid(x_V, mV)
    feats_v, _, masks_v, _ = mod.fusion_module(xs_v, x_A, ms_v, mA)
    xs_a, ms_a = pyramid(x_A, mA)
    feats_a, _, _, _ = mod.fusion_module(xs_a, x_V, ms_a, mV)
    return feats_v, feats_a, tuple(masks_v)


def alignment_forward(mod, **kwargs):
    """multimodal_backbones.py:1127-1207: inference part; returns (new_video_list, new_text_list, {})."""
    from .engine import HotPathEngine  # noqa: F401  (documentation pointer: the fused version lives there)
    video, text, mask = kwargs["video"][0], kwargs["text"][0], kwargs["mask_video"][0]
    _need_cuda(video)
    op, backend = _op()
    B, _, T = video.shape
    C, N1, dev = mod.num_hidden, T + 1, video.device
    half, hm = B * T, B * (T + 1)
    Xv, Xa = K.new_operand(half, video.shape[1], op, dev), K.new_operand(half, text.shape[1], op, dev)
    K.transpose_cast(video.contiguous().float(), T, Xv, B, video.shape[1], T, op)
    K.transpose_cast(text.contiguous().float(), T, Xa, B, text.shape[1], T, op)
    x0 = torch.empty(2 * half, C, dtype=torch.float32, device=dev)
    for X, lin, dst in ((Xv, mod.proj_fc_video[0], x0[:half]), (Xa, mod.proj_fc_text[0], x0[half:])):
        K.gemm([{"A": X, "W": _packed(lin.weight, "lin"), "bias": _vec(lin.bias), "out_f32": dst}], half, C, lin.weight.shape[1],
               op, K.ACT_NONE, False, backend)
    F = torch.empty(2 * hm, C, dtype=torch.float32, device=dev)
    K.align_embed(x0, _vec(mod.cls_token_video), _vec(mod.cls_token_text), mod.pos_embed_video[0, :N1].detach().float().contiguous(),
                  mod.pos_embed_text[0, :N1].detach().float().contiguous(), _vec(mod.type_video), _vec(mod.type_text), F, B, T, C)
    m_cls = torch.cat([torch.ones(B, 1, dtype=torch.uint8, device=dev), mask.reshape(B, T).to(torch.uint8)], 1).contiguous()
    lay = mod.multiway_list[0]
    att = lay.attn_fusion
    wqkv = K.pack_operand(torch.cat([att.q.weight, att.k.weight, att.v.weight], 0).detach().float(), op)
    bqkv = torch.cat([_vec(att.q.bias), _vec(att.k.bias), _vec(att.v.bias)])
    F1 = torch.empty_like(F)
    Fn, AO, Ha = K.new_operand(2 * hm, C, op, dev), K.new_operand(2 * hm, C, op, dev), K.new_operand(2 * hm, 4 * C, op, dev)
    QKV = torch.empty(2 * hm, 3 * C, dtype=torch.float32, device=dev)
    mods = (("video", lay.norm2_video, lay.ffn_video), ("text", lay.norm2_text, lay.ffn_text))
    nh = att._heads
    for _ in range(mod.num_layers):
        K.layernorm_rows([{"x": F, "w": _vec(lay.norm1_fused.weight), "b": _vec(lay.norm1_fused.bias), "out_op": Fn}], 2 * hm, C, op)
        K.gemm([{"A": Fn, "W": wqkv, "bias": bqkv, "out_f32": QKV}], 2 * hm, 3 * C, C, op, K.ACT_NONE, False, backend)
        groups = []
        for g in range(2):
            own, oth = QKV[g * hm:(g + 1) * hm], QKV[(1 - g) * hm:(2 - g) * hm]
            groups.append({"q": K.View(own, 0, C), "k": K.View(own, C, C), "v": K.View(own, 2 * C, C), "kmask": m_cls,
                           "xk": K.View(oth, C, C), "xv": K.View(oth, 2 * C, C), "x_first": 1, "out": AO[g * hm:(g + 1) * hm]})
        K.attention(groups, B, N1, N1, nh, C // nh, 1.0 / math.sqrt(C // nh), op)
        K.gemm([{"A": AO, "W": _packed(att.m.weight, "lin"), "bias": _vec(att.m.bias), "res": F, "out_f32": F1}], 2 * hm, C, C, op,
               K.ACT_NONE, False, backend)
        K.layernorm_rows([{"x": F1[g * hm:(g + 1) * hm], "w": _vec(n.weight), "b": _vec(n.bias), "out_op": Fn[g * hm:(g + 1) * hm]}
                          for g, (_, n, _f) in enumerate(mods)], hm, C, op)
        K.gemm([{"A": Fn[g * hm:(g + 1) * hm], "W": _packed(f.fc1.weight, "lin"), "bias": _vec(f.fc1.bias),
                 "out_op": Ha[g * hm:(g + 1) * hm]} for g, (_, _n, f) in enumerate(mods)], hm, 4 * C, C, op, K.ACT_GELU, False, backend)
        K.gemm([{"A": Ha[g * hm:(g + 1) * hm], "W": _packed(f.fc2.weight, "lin"), "bias": _vec(f.fc2.bias),
                 "res": F1[g * hm:(g + 1) * hm], "out_f32": F[g * hm:(g + 1) * hm]} for g, (_, _n, f) in enumerate(mods)],
               hm, C, 4 * C, op, K.ACT_NONE, False, backend)
    Z = K.new_operand(2 * half, C, op, dev)
    fin = ((mod.norm_video, mod.fc_video), (mod.norm_text, mod.fc_text))
    K.layernorm_rows([{"x": F[g * hm:(g + 1) * hm], "x_seg_rows": T, "x_seg_stride": N1, "x_row_off": 1,
                       "add": x0[g * half:(g + 1) * half], "w": _vec(n.weight), "b": _vec(n.bias),
                       "out_op": Z[g * half:(g + 1) * half]} for g, (n, _f) in enumerate(fin)], half, C, op)
    Y = torch.empty(2 * half, C, dtype=torch.float32, device=dev)
    K.gemm([{"A": Z[g * half:(g + 1) * half], "W": _packed(f[0].weight, "lin"), "bias": _vec(f[0].bias),
             "out_f32": Y[g * half:(g + 1) * half]} for g, (_n, f) in enumerate(fin)], half, C, C, op, K.ACT_RELU, False, backend)
    out = torch.empty(2 * half, C, dtype=torch.float32, device=dev)
    K.layernorm_rows([{"x": Y[g * half:(g + 1) * half], "w": _vec(f[3].weight), "b": _vec(f[3].bias),
                       "out_f32": out[g * half:(g + 1) * half]} for g, (_n, f) in enumerate(fin)], half, C, K.F32)
    return [from_rows(out[:half], B, T)], [from_rows(out[half:], B, T)], {}


# ------------------------------------------------------------------------------ multimodal_meta_archs.py
def _head_trunk(mod, x, mask):
    B, C, T = x.shape
    m = mask_rows(mask)
    r = to_rows(x)
    for conv, norm in zip(mod.head, mod.norm):
        r = ln_rows(norm, conv_rows(conv.conv, r, B, T, m), act=K.ACT_RELU)
    return r, m, B, T


def cls_head_forward(mod, fpn_feats, fpn_masks):
    """multimodal_meta_archs.py:166-178."""
    outs = tuple()
    for x, mask in zip(fpn_feats, fpn_masks):
        _need_cuda(x)
        r, m, B, T = _head_trunk(mod, x, mask)
        outs += (from_rows(conv_rows(mod.cls_head.conv, r, B, T, m), B, T),)
    return outs


def reg_head_forward(mod, fpn_feats, fpn_masks):
    """multimodal_meta_archs.py:245-259: relu(scale_l * (conv * mask))."""
    outs = tuple()
    for l, (x, mask) in enumerate(zip(fpn_feats, fpn_masks)):
        _need_cuda(x)
        r, m, B, T = _head_trunk(mod, x, mask)
        rs = mod.scale[l].scale.detach().float().reshape(1).expand(B * T).contiguous()
        outs += (from_rows(conv_rows(mod.offset_head.conv, r, B, T, m, rowscale=rs, act=K.ACT_RELU), B, T),)
    return outs


def inference_from_heads(model, video_list, fpn_masks, out_cls_logits, out_offsets):
    """multimodal_meta_archs.py:689-742: decode + soft-NMS + seconds from per-level [B,T_l,ncls] / [B,T_l,ncls,2]."""
    _need_cuda(out_cls_logits[0])
    dev = out_cls_logits[0].device
    B, ncls = out_cls_logits[0].shape[0], model.num_classes
    logits = torch.cat([x.float() for x in out_cls_logits], 1).contiguous()
    offsets = torch.cat([x.float().reshape(B, x.shape[1], -1) for x in out_offsets], 1).contiguous()
    masks = torch.cat([m.reshape(B, -1) for m in fpn_masks], 1).to(torch.uint8).contiguous()
    Tl = [x.shape[1] for x in out_cls_logits]
    off = [0]
    for t in Tl:
        off.append(off[-1] + t)
    pts = torch.cat([p[0].float() for p in video_list["points"]], 0).to(dev).contiguous()
    topk = model.test_pre_nms_topk
    cap = sum(min(topk, t * ncls) for t in Tl)
    cs = torch.empty(B, cap, 2, device=dev); csc = torch.empty(B, cap, device=dev)
    cl = torch.empty(B, cap, dtype=torch.int32, device=dev)
    K.decode(logits, offsets, masks, pts, off, B, ncls, model.class_aware, model.test_pre_nms_thresh, topk,
             model.test_duration_thresh, cs, csc, cl, cap)
    Kd = model.test_max_seg_num
    meta = torch.tensor([[float(video_list["feat_stride"][i]), float(video_list["feat_num_frames"][i]), float(video_list["fps"][i]),
                          float(video_list["duration"][i])] for i in range(B)], dtype=torch.float32, device=dev)
    o_s = torch.empty(B, Kd, 2, device=dev); o_sc = torch.empty(B, Kd, device=dev)
    o_l = torch.empty(B, Kd, dtype=torch.int64, device=dev); o_c = torch.empty(B, dtype=torch.int32, device=dev)
    ws = torch.empty(K.softnms_workspace_bytes(B, ncls, Kd), dtype=torch.uint8, device=dev)
    from .utils.nms import nms_method_code
    K.softnms_batched(cs, csc, cl, B, cap, ncls, model.test_iou_threshold, model.test_nms_sigma, model.test_min_score,
                      nms_method_code(model.test_nms_method, model.test_multiclass_nms), Kd, off[-1], meta, o_s, o_sc, o_l, o_c, ws)
    return model.collect_results({"out_segs": o_s, "out_scores": o_sc, "out_labels": o_l, "out_counts": o_c})
