"""Row kernels vs torch FP32 references of the same ops (LayerNorm family, depthwise conv + LN, copies,
transposes, token embedding, masks, pool + match projection, MaxSigmoid gate)."""
import pytest
import torch
import torch.nn.functional as F

from unav_yolyolva_b200 import kernels as K

pytestmark = pytest.mark.gpu


def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    r = x - mu
    return r / torch.sqrt((r * r).mean(-1, keepdim=True) + eps) * w + b


def _edges(nseg, T):
    e = torch.zeros(nseg, T, dtype=torch.uint8)
    e[:, 0] |= 1
    e[:, -1] |= 2
    return e.reshape(-1)


def _read_op(buf, Kc, op):
    if op == K.F32:
        return buf[:, :Kc].cpu()
    v = buf[:, :Kc].float().cpu()
    if op == K.BF16X2:
        h = buf.shape[1] // 2
        v = v + buf[:, h:h + Kc].float().cpu()
    return v


@pytest.mark.parametrize("C", [256, 512])
@pytest.mark.parametrize("op", [K.F32, K.BF16X2])
def test_layernorm_rows_all_outputs(cuda, C, op):
    g = torch.Generator().manual_seed(C)
    nseg, T = 3, 28
    M = nseg * T
    x = torch.randn(M, C, generator=g) * 2 + 0.3
    add = torch.randn(M, C, generator=g)
    w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    post = torch.randn(T, C, generator=g)
    mask = (torch.rand(M, generator=g) > 0.3).to(torch.uint8)
    of, oo, oi = torch.zeros(M, C, device=cuda), K.new_operand(M, C, op, cuda), K.new_operand(M, 3 * C, op, cuda)
    oi.fill_(7)
    K.layernorm_rows([{"x": x.to(cuda), "add": add.to(cuda), "w": w.to(cuda), "b": b.to(cuda), "post": post.to(cuda),
                       "post_rows": T, "rowmask": mask.to(cuda), "edge": _edges(nseg, T).to(cuda), "out_f32": of, "out_op": oo,
                       "out_im2col": oi}], M, C, op, act=K.ACT_GELU)
    ref = F.gelu(_ln(x + add, w, b)) + post.repeat(nseg, 1) * mask[:, None].float()
    tol = 2e-5 if op == K.F32 else 2e-4
    assert (of.cpu() - ref).abs().max() < 2e-5
    assert (_read_op(oo, C, op) - ref).abs().max() < tol
    r3 = ref.view(nseg, T, C)
    col = torch.cat([F.pad(r3, (0, 0, 1, 0))[:, :T], r3, F.pad(r3, (0, 0, 0, 1))[:, 1:]], -1).reshape(M, 3 * C)
    assert (_read_op(oi, 3 * C, op) - col).abs().max() < tol


@pytest.mark.parametrize("C", [256, 512, 1024, 384])
@pytest.mark.parametrize("act", [K.ACT_NONE, K.ACT_RELU, K.ACT_GELU, K.ACT_SILU])
def test_layernorm_rows_exact_width_kernels_same_bits(cuda, monkeypatch, C, act):
    """The exact-width / compile-time-activation instantiations (C = 256, 512, 1024) write the same bits as the generic
    kernel (UNAV_LN_GENERIC=1) on every output; C = 384 takes the generic kernel either way."""
    g = torch.Generator().manual_seed(C + act)
    nseg, T = 5, 23
    M = nseg * T
    x = (torch.randn(M, C, generator=g) * 2 + 0.3).to(cuda)
    add = torch.randn(M, C, generator=g).to(cuda)
    w, b = (torch.rand(C, generator=g) + 0.5).to(cuda), (torch.randn(C, generator=g) * 0.1).to(cuda)
    post = torch.randn(T, C, generator=g).to(cuda)
    mask = (torch.rand(M, generator=g) > 0.3).to(torch.uint8).to(cuda)
    edge = _edges(nseg, T).to(cuda)
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("UNAV_LN_GENERIC", "1")
        else:
            monkeypatch.delenv("UNAV_LN_GENERIC", raising=False)
        of, oo, oi = torch.zeros(M, C, device=cuda), K.new_operand(M, C, K.BF16X2, cuda), K.new_operand(M, 3 * C, K.BF16X2, cuda)
        K.layernorm_rows([{"x": x, "add": add, "w": w, "b": b, "post": post, "post_rows": T, "rowmask": mask, "edge": edge,
                           "out_f32": of, "out_op": oo, "out_im2col": oi}], M, C, K.BF16X2, act=act)
        o2 = torch.zeros(M, C, device=cuda)
        K.layernorm_rows([{"x": x, "w": w, "b": b, "out_f32": o2}], M, C, K.F32, act=act)
        outs.append((of, oo, oi, o2))
    torch.cuda.synchronize()
    for a, b_ in zip(*outs):
        assert torch.equal(a.view(torch.int16) if a.dtype == torch.bfloat16 else a.view(torch.int32),
                           b_.view(torch.int16) if b_.dtype == torch.bfloat16 else b_.view(torch.int32))


def test_layernorm_rows_seg_mapping_and_groups(cuda):
    g = torch.Generator().manual_seed(9)
    B, T, C = 3, 20, 512
    F_ = torch.randn(2, B, T + 1, C, generator=g)
    x0 = torch.randn(2, B, T, C, generator=g)
    ws = [(torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)) for _ in range(2)]
    Fd, x0d = F_.reshape(-1, C).to(cuda), x0.reshape(-1, C).to(cuda)
    out = torch.zeros(2 * B * T, C, device=cuda)
    hm, half = B * (T + 1), B * T
    K.layernorm_rows([{"x": Fd[m * hm:(m + 1) * hm], "x_seg_rows": T, "x_seg_stride": T + 1, "x_row_off": 1,
                       "add": x0d[m * half:(m + 1) * half], "w": ws[m][0].to(cuda), "b": ws[m][1].to(cuda),
                       "out_f32": out[m * half:(m + 1) * half]} for m in range(2)], half, C, K.F32)
    for m in range(2):
        ref = _ln(F_[m, :, 1:] + x0[m], *ws[m]).reshape(-1, C)
        assert (out[m * half:(m + 1) * half].cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("stride,C,n_pre", [(1, 512, 2), (1, 256, 0), (2, 512, 0)])
def test_dwconv_ln(cuda, stride, C, n_pre):
    g = torch.Generator().manual_seed(stride * 7 + C)
    nseg, T = 4, 56
    To = T // stride
    ld = C + 256                               # strided input view (column slice of a wider buffer)
    xb = torch.randn(nseg * T, ld, generator=g)
    x = xb[:, 128:128 + C]
    mask = (torch.rand(nseg * To, generator=g) > 0.2).to(torch.uint8)
    pre = [(torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2) for _ in range(n_pre)]
    outs, outs_d = [], []
    n_out = 3 if stride == 1 else 1
    for j in range(n_out):
        outs.append({"dw": torch.randn(C, 1, 3, generator=g), "w": torch.rand(C, generator=g) + 0.5,
                     "b": torch.randn(C, generator=g) * 0.1, "src": (j % 2 if n_pre else -1)})
    xbd = xb.to(cuda)
    for o in outs:
        buf = torch.zeros(nseg * To, C, device=cuda)
        outs_d.append({"dw": o["dw"].reshape(-1).to(cuda), "ln_w": o["w"].to(cuda), "ln_b": o["b"].to(cuda), "src": o["src"],
                       "out_f32": buf})
    K.dwconv_ln([{"x": K.View(xbd, 128, C), "mask_out": mask.to(cuda), "pre": [(a.to(cuda), b.to(cuda)) for a, b in pre],
                  "outs": outs_d}], nseg, T, stride, C, K.F32)
    for o, od in zip(outs, outs_d):
        xin = x if o["src"] < 0 else _ln(x, *pre[o["src"]])
        xc = xin.reshape(nseg, T, C).transpose(1, 2)
        y = F.conv1d(xc, o["dw"], None, stride=stride, padding=1, groups=C).transpose(1, 2).reshape(-1, C)
        ref = _ln(y * mask[:, None].float(), o["w"], o["b"])
        assert (od["out_f32"].cpu() - ref).abs().max() < 3e-5


@pytest.mark.parametrize("C,stride,n_pre,n_out,nseg,T", [(512, 1, 2, 3, 6, 224), (256, 1, 0, 3, 5, 56), (512, 2, 0, 1, 4, 112),
                                                      (256, 2, 0, 1, 3, 14), (512, 1, 0, 2, 3, 7), (256, 1, 2, 3, 2, 1)])
def test_dwconv_ln_streaming_equals_tiled(cuda, monkeypatch, C, stride, n_pre, n_out, nseg, T):
    """The streaming dwconv_ln kernel (weights staged once per CTA, 3-row register window per warp) keeps the tiled kernel's
    per-lane channel mapping and operation order: same bits on FP32 and split-operand outputs, for both strides, with and
    without pre-LayerNorms, plain depthwise conv (no LN) included, short segments and strips that do not divide them."""
    if T % stride:
        pytest.skip("odd length")
    g = torch.Generator().manual_seed(C + stride + T)
    To = T // stride
    xb = torch.randn(nseg * T, C + 128, generator=g).to(cuda)
    mask = (torch.rand(nseg * To, generator=g) > 0.2).to(torch.uint8).to(cuda)
    pre = [((torch.rand(C, generator=g) + 0.5).to(cuda), (torch.randn(C, generator=g) * 0.2).to(cuda)) for _ in range(n_pre)]
    wts = [{"dw": torch.randn(3 * C, generator=g).to(cuda), "ln_w": (torch.rand(C, generator=g) + 0.5).to(cuda),
            "ln_b": (torch.randn(C, generator=g) * 0.1).to(cuda), "src": (j % 2 if n_pre else -1)} for j in range(n_out)]
    if n_out == 2:
        wts[1]["ln_w"] = wts[1]["ln_b"] = None            # plain masked depthwise conv
    res = {}
    for mode, strip in (("tiled", None), ("stream", None), ("stream", "3")):
        monkeypatch.delenv("UNAV_DWCONV_TILED", raising=False)
        monkeypatch.delenv("UNAV_DWCONV_STRIP", raising=False)
        monkeypatch.setenv("UNAV_DWCONV_STREAM", "1")
        if mode == "tiled":
            monkeypatch.setenv("UNAV_DWCONV_TILED", "1")
        if strip:
            monkeypatch.setenv("UNAV_DWCONV_STRIP", strip)
        outs = []
        for w in wts:
            d = {k: v for k, v in w.items() if v is not None}
            d["out_f32"] = torch.zeros(nseg * To, C, device=cuda)
            d["out_op"] = K.new_operand(nseg * To, C, K.BF16X2, cuda)
            outs.append(d)
        K.dwconv_ln([{"x": K.View(xb, 64, C), "mask_out": mask, "pre": pre, "outs": outs}], nseg, T, stride, C, K.BF16X2)
        torch.cuda.synchronize()
        res[(mode, strip)] = [(o["out_f32"].clone(), o["out_op"].clone()) for o in outs]
    for key in (("stream", None), ("stream", "3")):
        for (a32, aop), (b32, bop) in zip(res[("tiled", None)], res[key]):
            assert torch.equal(a32, b32) and torch.equal(aop, bop), key


def test_rowcopy_upsample_im2col_stride2(cuda):
    g = torch.Generator().manual_seed(2)
    nseg, C = 3, 256
    x = torch.randn(nseg * 14, C, generator=g)
    xd = x.to(cuda)
    up = torch.zeros(nseg * 28, 2 * C, device=cuda)
    K.rowcopy([{"src": xd, "dst": K.View(up, C, C), "nseg": nseg, "seg_len_in": 14, "seg_len_out": 28, "num": 1, "den": 2,
                "C": C}], K.F32)
    ref = x.view(nseg, 14, C).repeat_interleave(2, dim=1).reshape(-1, C)
    assert torch.equal(up[:, C:].cpu(), ref) and up[:, :C].abs().max() == 0
    s2 = torch.zeros(nseg * 7, 3 * C, device=cuda)
    K.rowcopy([{"src": xd, "dst": s2, "nseg": nseg, "seg_len_in": 14, "seg_len_out": 7, "num": 2, "den": 1, "ntaps": 3,
                "tap_stride": C, "C": C}], K.F32)
    xp = F.pad(x.view(nseg, 14, C), (0, 0, 1, 1))
    ref = torch.cat([xp[:, 0:14:2], xp[:, 1:15:2], xp[:, 2:16:2]], -1).reshape(-1, 3 * C)
    assert torch.equal(s2.cpu(), ref)
    # video-major scatter of two levels with a BF16 split destination
    dst = K.new_operand(nseg * 21, 3 * C, K.BF16X2, cuda)
    y = torch.randn(nseg * 7, C, generator=g)
    K.rowcopy([{"src": xd, "dst": dst, "nseg": nseg, "seg_len_in": 14, "seg_len_out": 14, "dst_seg_stride": 21, "dst_row_off": 0,
                "ntaps": 3, "tap_stride": C, "C": C},
               {"src": y.to(cuda), "dst": dst, "nseg": nseg, "seg_len_in": 7, "seg_len_out": 7, "dst_seg_stride": 21,
                "dst_row_off": 14, "ntaps": 3, "tap_stride": C, "C": C}], K.BF16X2)
    got = (dst[:, :3 * C].float() + dst[:, 3 * C:].float()).cpu().view(nseg, 21, 3 * C)
    def col(t, T):
        p = F.pad(t.view(nseg, T, C), (0, 0, 1, 1))
        return torch.cat([p[:, 0:T], p[:, 1:T + 1], p[:, 2:T + 2]], -1)
    assert (got[:, :14] - col(x, 14)).abs().max() < 2e-4 and (got[:, 14:] - col(y, 7)).abs().max() < 2e-4


@pytest.mark.parametrize("op", [K.BF16X2, K.BF16])
def test_rowcopy_im2col_scatter_equals_gather(cuda, monkeypatch, op):
    """The scatter-form k = 3 im2col (one thread per source element, three stores) writes the same bytes as the gather form
    (UNAV_ROWCOPY_GATHER=1): segments of length 1, 7 and 224, column windows of wider source / destination buffers, a
    destination with another segment stride and row offset, edge taps zero-filled."""
    g = torch.Generator().manual_seed(21)
    nseg, C = 5, 256
    jobs_def = [(1, 0), (7, 1), (224, 8)]                 # (segment length, row offset inside the destination segment)
    stride_rows = 1 + 7 + 224 + 3
    srcs = [torch.randn(nseg * T, C + 64, generator=g).to(cuda) for T, _ in jobs_def]
    outs = []
    for gather in ("1", None):
        if gather:
            monkeypatch.setenv("UNAV_ROWCOPY_GATHER", gather)
        else:
            monkeypatch.delenv("UNAV_ROWCOPY_GATHER", raising=False)
        dst = K.new_operand(nseg * stride_rows, 2 * 3 * C, op, cuda)
        dst.fill_(3.0)
        jobs = [{"src": K.View(src, 32, C), "dst": K.View(dst, 3 * C, 3 * C), "nseg": nseg, "seg_len_in": T, "seg_len_out": T,
                 "dst_seg_stride": stride_rows, "dst_row_off": off, "ntaps": 3, "tap_stride": C, "C": C}
                for src, (T, off) in zip(srcs, jobs_def)]
        K.rowcopy(jobs, op)
        torch.cuda.synchronize()
        outs.append(dst.clone())
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))


def test_transpose_align_embed_masks(cuda):
    g = torch.Generator().manual_seed(4)
    nb, R, Cc = 3, 128, 224
    x = torch.randn(nb, R, Cc, generator=g)
    out = K.new_operand(nb * Cc, R, K.BF16, cuda)
    K.transpose_cast(x.to(cuda), Cc, out, nb, R, Cc, K.BF16)
    assert torch.equal(out[:, :R].cpu(), x.transpose(1, 2).reshape(-1, R).to(torch.bfloat16))
    B, T, C = 2, 24, 512
    x0 = torch.randn(2, B, T, C, generator=g)
    par = [torch.randn(C, generator=g) for _ in range(4)]
    pos = [torch.randn(T + 1, C, generator=g) for _ in range(2)]
    tok = torch.zeros(2 * B * (T + 1), C, device=cuda)
    K.align_embed(x0.to(cuda), par[0].to(cuda), par[1].to(cuda), pos[0].to(cuda), pos[1].to(cuda), par[2].to(cuda),
                  par[3].to(cuda), tok, B, T, C)
    for m in range(2):
        seq = torch.cat([par[m].expand(B, 1, C), x0[m]], 1) + pos[m] + par[2 + m]
        assert torch.equal(tok.cpu().view(2, B, T + 1, C)[m], seq)
    T, L = 32, 4
    mask = (torch.arange(T)[None] < torch.tensor([[13], [32], [1]])).to(torch.uint8)
    nbs, nbt = 3, 6
    Ttot = sum(T >> l for l in range(L))
    mt, mu = torch.zeros(nbt * Ttot, dtype=torch.uint8, device=cuda), torch.zeros(nbt * (Ttot - (T >> (L - 1))), dtype=torch.uint8, device=cuda)
    mc, mh = torch.zeros(nbs, T + 1, dtype=torch.uint8, device=cuda), torch.zeros(nbs, Ttot, dtype=torch.uint8, device=cuda)
    K.build_masks(mask.to(cuda), mt, mu, mc, mh, nbt, nbs, T, L)
    off = offu = 0
    m2 = mask.repeat(2, 1)
    lvl = 0
    for l in range(L):
        Tl = T >> l
        ref = m2[:, ::(1 << l)]
        assert torch.equal(mt[off:off + nbt * Tl].cpu().view(nbt, Tl), ref)
        assert torch.equal(mh[:, lvl:lvl + Tl].cpu(), ref[:nbs])
        if l + 1 < L:
            refu = m2[:, ::(2 << l)].repeat_interleave(2, dim=1)
            assert torch.equal(mu[offu:offu + nbt * Tl].cpu().view(nbt, Tl), refu)
            offu += nbt * Tl
        off += nbt * Tl
        lvl += Tl
    assert torch.equal(mc.cpu(), torch.cat([torch.ones(nbs, 1, dtype=torch.uint8), mask], 1))


def test_pool_match_and_maxsig(cuda):
    g = torch.Generator().manual_seed(6)
    nb, C, Tq = 3, 512, 224
    us = [torch.randn(nb * T, C, generator=g) for T in (224, 112, 56)]
    Wm, bm = torch.randn(Tq, 12, generator=g), torch.randn(Tq, generator=g)
    q = torch.zeros(nb * Tq, C, device=cuda)
    K.pool_match(us[0].to(cuda), us[1].to(cuda), us[2].to(cuda), 224, 112, 56, Wm.to(cuda), bm.to(cuda), q, nb, C, Tq, 4)
    pooled = torch.cat([F.adaptive_avg_pool1d(u.view(nb, -1, C).transpose(1, 2), 4) for u in us], -1)   # [nb, C, 12]
    ref = F.conv1d(pooled.transpose(1, 2), Wm[:, :, None], bm).transpose(1, 2)                          # [nb, C, Tq]
    assert (q.cpu().view(nb, Tq, C) - ref.transpose(1, 2)).abs().max() < 3e-5
    for H in (8, 4):
        T, Ce, nw = 56, 256, 512
        hc = Ce // H
        x = torch.randn(nb * T, Ce, generator=g)
        G = torch.randn(nb * nw, 1280, generator=g)
        hb = torch.randn(H, generator=g)
        gate = torch.zeros(nb * T, H, device=cuda)
        K.maxsig_gate(x.to(cuda), K.View(G.to(cuda), 512, Ce), hb.to(cuda), gate, nb, T, nw, H, hc)
        emb = x.view(nb, T, H, hc)
        gd = G[:, 512:512 + Ce].reshape(nb, nw, H, hc)
        aw = torch.einsum("bthc,bnhc->bhtn", emb, gd).max(-1)[0] / hc ** 0.5 + hb[None, :, None]
        ref = aw.sigmoid().permute(0, 2, 1).reshape(nb * T, H)
        assert (gate.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("op,tol", [(K.BF16X2, 3e-5), (K.BF16, 2e-2)])
def test_maxsig_gate_tensor_core(cuda, op, tol):
    g = torch.Generator().manual_seed(12)
    for nb, T, H in ((3, 224, 8), (2, 56, 4), (4, 7, 8), (2, 112, 4)):
        Ce, nw = 256, 512
        hc = Ce // H
        xw = torch.randn(nb * T, 1536, generator=g)              # x is the window [1024, 1280) of a concat buffer
        Gw = torch.randn(nb * nw, 1280, generator=g)             # G is the window [512, 768) of the 5-layer guide_fc output
        hb = torch.randn(H, generator=g)
        xo, Go = K.pack_operand(xw.to(cuda), op), K.pack_operand(Gw.to(cuda), op)
        gate = torch.zeros(nb * T, H, device=cuda)
        K.maxsig_gate_tc(xo, 1024, Go, 512, hb.to(cuda), gate, nb, T, nw, H, hc, op)
        torch.cuda.synchronize()
        def rt(x):
            hi = x.to(torch.bfloat16).float()
            return hi + ((x - hi).to(torch.bfloat16).float() if op == K.BF16X2 else 0)
        emb = rt(xw[:, 1024:1280]).view(nb, T, H, hc)
        gd = rt(Gw[:, 512:768]).reshape(nb, nw, H, hc)
        aw = torch.einsum("bthc,bnhc->bhtn", emb.double(), gd.double()).max(-1)[0] / hc ** 0.5 + hb[None, :, None]
        ref = aw.sigmoid().permute(0, 2, 1).reshape(nb * T, H).float()
        assert (gate.cpu() - ref).abs().max() < tol, (nb, T, H)
