"""Fused attention kernel vs torch FP32 reference (masked softmax attention, optional per-query extra key)."""
import math

import pytest
import torch

from unav_yolyolva_b200 import kernels as K

pytestmark = pytest.mark.gpu


def _ref(q, k, v, kmask, nh, hs, scale, xk=None, xv=None, x_first=0):
    nb, Tq, _ = q.shape
    Tk = k.shape[1]
    out = torch.zeros(nb, Tq, nh * hs, dtype=torch.float64)
    for b in range(nb):
        for h in range(nh):
            sl = slice(h * hs, (h + 1) * hs)
            s = (q[b, :, sl].double() @ k[b, :, sl].double().t()) * scale
            s = s.masked_fill(~kmask[b].bool()[None, :], float("-inf"))
            vv = v[b, :, sl].double()
            if xk is not None:
                sx = (q[b, :, sl].double() * xk[b, :, sl].double()).sum(-1) * scale
                sx[:x_first] = float("-inf")
                s = torch.cat([s, sx[:, None]], 1)
                p = s.softmax(-1)
                out[b, :, sl] = p[:, :Tk] @ vv + p[:, Tk:] * xv[b, :, sl].double()
            else:
                out[b, :, sl] = s.softmax(-1) @ vv
    return out.float()


@pytest.mark.parametrize("nb,T,nh,hs", [(3, 224, 4, 128), (2, 112, 4, 64), (5, 7, 4, 64), (2, 56, 4, 64), (1, 300, 2, 64)])
def test_mhca_attention(cuda, nb, T, nh, hs):
    g = torch.Generator().manual_seed(T)
    C = nh * hs
    q, k, v = (torch.randn(nb, T, C, generator=g) for _ in range(3))
    lens = torch.randint(1, T + 1, (nb,), generator=g)
    kmask = (torch.arange(T)[None] < lens[:, None]).to(torch.uint8)
    out = torch.zeros(nb * T, C, device=cuda)
    K.attention([{"q": q.reshape(-1, C).to(cuda), "k": k.reshape(-1, C).to(cuda), "v": v.reshape(-1, C).to(cuda),
                  "kmask": kmask.to(cuda), "out": out}], nb, T, T, nh, hs, 1 / math.sqrt(hs), K.F32)
    ref = _ref(q, k, v, kmask, nh, hs, 1 / math.sqrt(hs))
    assert (out.cpu().view(nb, T, C) - ref).abs().max() < 2e-5


def test_alignment_attention_with_cross_key(cuda):
    g = torch.Generator().manual_seed(3)
    nb, N, nh, hs = 2, 225, 8, 64
    C = nh * hs
    qkv = [torch.randn(2, nb, N, 3 * C, generator=g) for _ in range(1)][0]     # [modality, b, token, qkv]
    lens = torch.tensor([180, 61])
    kmask = torch.cat([torch.ones(nb, 1), (torch.arange(N - 1)[None] < lens[:, None]).float()], 1).to(torch.uint8)
    dq = qkv.reshape(2 * nb * N, 3 * C).to(cuda)
    out = torch.zeros(2 * nb * N, C, device=cuda)
    hm = nb * N
    groups = []
    for m in range(2):
        own, oth = dq[m * hm:(m + 1) * hm], dq[(1 - m) * hm:(2 - m) * hm]
        groups.append({"q": K.View(own, 0, C), "k": K.View(own, C, C), "v": K.View(own, 2 * C, C), "kmask": kmask.to(cuda),
                       "xk": K.View(oth, C, C), "xv": K.View(oth, 2 * C, C), "x_first": 1, "out": out[m * hm:(m + 1) * hm]})
    K.attention(groups, nb, N, N, nh, hs, 0.125, K.F32)
    for m in range(2):
        q, k, v = qkv[m, ..., :C], qkv[m, ..., C:2 * C], qkv[m, ..., 2 * C:]
        xk, xv = qkv[1 - m, ..., C:2 * C], qkv[1 - m, ..., 2 * C:]
        ref = _ref(q, k, v, kmask, nh, hs, 0.125, xk, xv, 1)
        assert (out[m * hm:(m + 1) * hm].cpu().view(nb, N, C) - ref).abs().max() < 2e-5


# ------------------------------------------------------------------------------------------------------------------
# tcgen05 / TMEM version
def _tc_inputs(cuda, q, k, v, op):
    """operand-dtype q, k rows + transposed values [C, nb*T]."""
    nb, T, C = q.shape
    qo = K.pack_operand(q.reshape(-1, C).to(cuda), op)
    ko = K.pack_operand(k.reshape(-1, C).to(cuda), op)
    vt = K.new_operand(nb * C, T, op, cuda)
    K.transpose_cast(v.reshape(-1, C).to(cuda).contiguous(), C, vt, nb, T, C, op)
    return qo, ko, vt


def _rt(x, op):
    dt = K.OP_TORCH_DTYPE[op]
    hi = x.to(dt).float()
    return hi + ((x - hi).to(dt).float() if op in K.SPLIT_DTYPES else 0)


def _read(buf, C, op):
    v = buf[:, :C].float()
    if op in K.SPLIT_DTYPES:
        v = v + buf[:, buf.shape[1] // 2: buf.shape[1] // 2 + C].float()
    return v.cpu()


@pytest.mark.parametrize("nb,T,nh,hs", [(3, 224, 4, 128), (2, 224, 4, 64), (2, 112, 4, 64), (3, 56, 4, 64), (2, 28, 4, 64),
                                         (5, 14, 4, 64), (5, 7, 4, 64), (2, 200, 2, 128)])
@pytest.mark.parametrize("op,tol", [(K.BF16X2, 4e-5), (K.BF16, 2e-2), (K.F16X2, 4e-5), (K.F16, 3e-3)])
def test_tc_attention(cuda, nb, T, nh, hs, op, tol):
    g = torch.Generator().manual_seed(T + hs)
    C = nh * hs
    q, k, v = (torch.randn(nb, T, C, generator=g) for _ in range(3))
    lens = torch.randint(1, T + 1, (nb,), generator=g)
    kmask = (torch.arange(T)[None] < lens[:, None]).to(torch.uint8)
    qo, ko, vt = _tc_inputs(cuda, q, k, v, op)
    out = K.new_operand(nb * T, C, op, cuda)
    K.attention_tc([{"q": qo, "k": ko, "vt": vt, "kmask": kmask.to(cuda), "out": out}], nb, T, T, nh, hs, 1 / math.sqrt(hs), op)
    torch.cuda.synchronize()
    ref = _ref(_rt(q, op), _rt(k, op), _rt(v, op), kmask, nh, hs, 1 / math.sqrt(hs))
    err = (_read(out, C, op).view(nb, T, C) - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("nb,T,nh,hs", [(4, 224, 4, 64), (3, 224, 4, 128), (3, 112, 4, 64), (4, 225, 8, 64)])
def test_tc_attention_skips_masked_work(cuda, nb, T, nh, hs):
    """Keys beyond an item's last valid key are neither loaded nor multiplied (cost follows the valid length), and with a
    query mask the 128-query tiles without a valid query are skipped: valid rows equal the reference (and the un-masked
    call to rounding), rows of skipped tiles are exactly zero (finite, so the masked projection that follows stays finite)."""
    op = K.BF16X2
    g = torch.Generator().manual_seed(11 * T + hs)
    C = nh * hs
    q, k, v = (torch.randn(nb, T, C, generator=g) for _ in range(3))
    lens = torch.tensor([1, 70, 128, 129, T][:nb] if nb > 3 else [17, 128, T - 5])
    kmask = (torch.arange(T)[None] < lens[:, None]).to(torch.uint8)
    qo, ko, vt = _tc_inputs(cuda, q, k, v, op)
    out_a = K.new_operand(nb * T, C, op, cuda); out_a.fill_(7.0)
    out_b = K.new_operand(nb * T, C, op, cuda)
    scale = 1 / math.sqrt(hs)
    K.attention_tc([{"q": qo, "k": ko, "vt": vt, "kmask": kmask.to(cuda), "qmask": kmask.to(cuda), "out": out_a}], nb, T, T, nh, hs, scale, op)
    K.attention_tc([{"q": qo, "k": ko, "vt": vt, "kmask": kmask.to(cuda), "out": out_b}], nb, T, T, nh, hs, scale, op)
    torch.cuda.synchronize()
    ref = _ref(_rt(q, op), _rt(k, op), _rt(v, op), kmask, nh, hs, scale)
    a, b_ = _read(out_a, C, op).view(nb, T, C), _read(out_b, C, op).view(nb, T, C)
    tol = 4e-5 * max(1.0, ref.abs().max().item())
    assert (b_ - ref).abs().max().item() < tol
    for i in range(nb):
        L = int(lens[i])
        assert (a[i, :L] - ref[i, :L]).abs().max().item() < tol                  # valid queries
        assert torch.equal(a[i, :L], b_[i, :L])                                   # same bits with and without the query mask
        first_skipped = ((L + 127) // 128) * 128                                  # tiles that hold no valid query
        if first_skipped < T:
            assert torch.count_nonzero(a[i, first_skipped:]) == 0
        assert torch.isfinite(a[i]).all()


@pytest.mark.parametrize("op,tol", [(K.BF16X2, 4e-5), (K.BF16, 2e-2)])
def test_tc_alignment_attention_with_cross_key(cuda, op, tol):
    g = torch.Generator().manual_seed(3)
    nb, N, nh, hs = 2, 225, 8, 64
    C = nh * hs
    qkv = torch.randn(2, nb, N, 3 * C, generator=g)
    lens = torch.tensor([180, 61])
    kmask = torch.cat([torch.ones(nb, 1), (torch.arange(N - 1)[None] < lens[:, None]).float()], 1).to(torch.uint8)
    d32 = qkv.reshape(2 * nb * N, 3 * C).to(cuda)
    dop = K.pack_operand(d32, op)
    hm = nb * N
    out = K.new_operand(2 * hm, C, op, cuda)
    vts = []
    for m in range(2):
        vt = K.new_operand(nb * C, N, op, cuda)
        K.transpose_cast(K.View(d32[m * hm:(m + 1) * hm], 2 * C, C), 3 * C, vt, nb, N, C, op)
        vts.append(vt)
    groups = []
    for m in range(2):
        own, oth = dop[m * hm:(m + 1) * hm], d32[(1 - m) * hm:(2 - m) * hm]
        groups.append({"q": K.View(own, 0, C), "k": K.View(own, C, C), "vt": vts[m], "kmask": kmask.to(cuda),
                       "q32": K.View(d32[m * hm:(m + 1) * hm], 0, C), "xk": K.View(oth, C, C), "xv": K.View(oth, 2 * C, C),
                       "x_first": 1, "out": out[m * hm:(m + 1) * hm]})
    K.attention_tc(groups, nb, N, N, nh, hs, 0.125, op)
    torch.cuda.synchronize()
    for m in range(2):
        q, k, v = qkv[m, ..., :C], qkv[m, ..., C:2 * C], qkv[m, ..., 2 * C:]
        xk, xv = qkv[1 - m, ..., C:2 * C], qkv[1 - m, ..., 2 * C:]
        ref = _ref(_rt(q, op), _rt(k, op), _rt(v, op), kmask, nh, hs, 0.125, xk, xv, 1)
        err = (_read(out[m * hm:(m + 1) * hm], C, op).view(nb, N, C) - ref).abs().max().item()
        assert err < tol * 4, err


@pytest.mark.parametrize("nb,T,nh,hs", [(2, 300, 4, 64), (2, 576, 4, 128), (1, 1152, 4, 64), (1, 2304, 4, 128), (3, 257, 2, 64)])
@pytest.mark.parametrize("op,tol", [(K.BF16X2, 6e-5), (K.BF16, 2e-2)])
def test_tc_attention_long_keys_chunked(cuda, nb, T, nh, hs, op, tol):
    """Key lengths above 256 (BASELINE.json config 4: T = 2304 and the pyramid levels 1152 / 576 / 288): the same tcgen05
    kernel on 256-key chunks + the merge kernel, vs the FP64 reference; valid lengths that end inside a chunk, exactly at a
    chunk boundary and before the first chunk's end (trailing chunks skipped), ragged last chunk (T % 256 != 0)."""
    g = torch.Generator().manual_seed(T + hs)
    C = nh * hs
    q, k, v = (torch.randn(nb, T, C, generator=g) for _ in range(3))
    lens = torch.randint(1, T + 1, (nb,), generator=g)
    lens[0] = min(T, 256)                                    # boundary case
    if nb > 1:
        lens[1] = 70                                         # everything beyond the first chunk is skipped
    kmask = (torch.arange(T)[None] < lens[:, None]).to(torch.uint8)
    qo, ko, vt = _tc_inputs(cuda, q, k, v, op)
    out = K.new_operand(nb * T, C, op, cuda)
    ws = torch.empty(K.attention_tc_workspace_bytes(1, nb, T, T, nh, hs), dtype=torch.uint8, device=cuda)
    assert ws.numel() > 0
    K.attention_tc([{"q": qo, "k": ko, "vt": vt, "kmask": kmask.to(cuda), "out": out}], nb, T, T, nh, hs, 1 / math.sqrt(hs), op,
                   workspace=ws)
    torch.cuda.synchronize()
    ref = _ref(_rt(q, op), _rt(k, op), _rt(v, op), kmask, nh, hs, 1 / math.sqrt(hs))
    err = (_read(out, C, op).view(nb, T, C) - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err


def test_tc_alignment_attention_long_with_cross_key(cuda):
    """The Alignment layers at T = 2304 (2305 tokens per modality, 8 heads, the time-aligned token of the other modality as a
    per-query extra key): chunked tcgen05 kernel + merge with the extra key, two groups in one launch."""
    op, tol = K.BF16X2, 6e-5
    g = torch.Generator().manual_seed(5)
    nb, N, nh, hs = 1, 577, 8, 64
    C = nh * hs
    qkv = torch.randn(2, nb, N, 3 * C, generator=g)
    lens = torch.tensor([400])
    kmask = torch.cat([torch.ones(nb, 1), (torch.arange(N - 1)[None] < lens[:, None]).float()], 1).to(torch.uint8)
    d32 = qkv.reshape(2 * nb * N, 3 * C).to(cuda)
    dop = K.pack_operand(d32, op)
    hm = nb * N
    out = K.new_operand(2 * hm, C, op, cuda)
    vts = []
    for m in range(2):
        vt = K.new_operand(nb * C, N, op, cuda)
        K.transpose_cast(K.View(d32[m * hm:(m + 1) * hm], 2 * C, C), 3 * C, vt, nb, N, C, op)
        vts.append(vt)
    groups = []
    for m in range(2):
        own, oth = dop[m * hm:(m + 1) * hm], d32[(1 - m) * hm:(2 - m) * hm]
        groups.append({"q": K.View(own, 0, C), "k": K.View(own, C, C), "vt": vts[m], "kmask": kmask.to(cuda),
                       "q32": K.View(d32[m * hm:(m + 1) * hm], 0, C), "xk": K.View(oth, C, C), "xv": K.View(oth, 2 * C, C),
                       "x_first": 1, "out": out[m * hm:(m + 1) * hm]})
    ws = torch.empty(K.attention_tc_workspace_bytes(2, nb, N, N, nh, hs), dtype=torch.uint8, device=cuda)
    K.attention_tc(groups, nb, N, N, nh, hs, 0.125, op, workspace=ws)
    torch.cuda.synchronize()
    for m in range(2):
        q, k, v = qkv[m, ..., :C], qkv[m, ..., C:2 * C], qkv[m, ..., 2 * C:]
        xk, xv = qkv[1 - m, ..., C:2 * C], qkv[1 - m, ..., 2 * C:]
        ref = _ref(_rt(q, op), _rt(k, op), _rt(v, op), kmask, nh, hs, 0.125, xk, xv, 1)
        err = (_read(out[m * hm:(m + 1) * hm], C, op).view(nb, N, C) - ref).abs().max().item()
        assert err < tol * 4, err
