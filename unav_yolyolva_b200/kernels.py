"""Thin Python wrappers over the C ABI: torch tensors in, raw device pointers + leading dimensions out.

Every function launches hand-written sm_100a kernels from ``libunav_b200.so`` on torch's *current* CUDA
stream (so calls can be captured into a CUDA graph).  Tensors are only used as memory handles.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _cabi as A
from ._cabi import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, BF16, BF16X2, F16, F16X2, F32, GEMM_SIMT,  # noqa: F401
                    GEMM_TCGEN05, SPLIT_DTYPES)

OP_TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16, BF16X2: torch.bfloat16, F16: torch.float16, F16X2: torch.float16}


# kernel names as ncu lists them, by unav_gemm_last_variant()
GEMM_KERNELS = {0: "gemm_tcgen05_kernel<64, 64>", 1: "gemm_tcgen05_kernel<128, 32>", 2: "gemm_tcgen05_kernel<128, 64>",
                3: "gemm_tcgen05_pair_kernel<256>", 4: "gemm_tcgen05_kernel<64, 32>", 5: "gemm_tcgen05_pair_kernel<128>",
                6: "gemm_tcgen05_kernel<256, 32>", 7: "gemm_tcgen05_ppair_kernel<256, 8>"}


def with_passes(op_dtype: int, passes: int = 0) -> int:
    """op_dtype argument carrying a pass count for split operands (include/unav_b200.h UNAV_PASSES): 0 = all three."""
    return op_dtype | (passes << 8) if op_dtype in SPLIT_DTYPES else op_dtype


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-launch tracing (bench.py's roofline leg): CUDA events on the launching stream around each
# C-ABI call, plus the algorithmic FLOPs / bytes of the call.  Off (None) on the product path.
_TRACE: Optional[list] = None
_TRACE_EXTERNAL = False


def start_trace(external: bool = False) -> None:
    """external=True: the events become event-record NODES when the calls are being captured into a CUDA graph, so the
    same spans can be read after every replay (per-kernel times without the CPU launch gaps of an eager pass)."""
    global _TRACE, _TRACE_EXTERNAL
    _TRACE = []
    _TRACE_EXTERNAL = external


def stop_trace(read: bool = True) -> list:
    """Returns [(kernel, ms, flops, bytes, note)] for the calls issued since start_trace(); with read=False the raw
    [(kernel, ev_a, ev_b, flops, bytes, note)] records (for spans captured into a graph: read them after a replay)."""
    global _TRACE
    rec, _TRACE = _TRACE or [], None
    if not read:
        return rec
    torch.cuda.synchronize()
    return read_trace(rec)


def read_trace(rec: list) -> list:
    return [(name, a.elapsed_time(b), fl, by, note) for name, a, b, fl, by, note in rec]


class _Span:
    def __init__(self, name, flops=0, nbytes=0, note=""):
        self.args = (name, flops, nbytes, note)

    def __enter__(self):
        if _TRACE is not None:
            kw = {"external": True} if _TRACE_EXTERNAL else {}
            self.a = torch.cuda.Event(enable_timing=True, **kw)
            self.b = torch.cuda.Event(enable_timing=True, **kw)
            self.a.record()
        return self

    def rename(self, name):
        self.args = (name,) + self.args[1:]

    def __exit__(self, *exc):
        if _TRACE is not None and exc[0] is None:
            self.b.record()
            name, fl, by, note = self.args
            _TRACE.append((name, self.a, self.b, fl, by, note))
        return False


def _p(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _ld(t) -> int:
    if t is None:
        return 0
    if t.dim() == 1:
        return t.shape[0]
    assert t.stride(-1) == 1, "innermost dimension must be contiguous"
    return t.stride(-2)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def op_cols(K: int, op_dtype: int) -> int:
    """Physical row length (elements) of an operand buffer with K logical columns."""
    if op_dtype == F32:
        return round_up(K, 4)
    kp = round_up(K, 8)
    return 2 * kp if op_dtype in SPLIT_DTYPES else kp


def new_operand(rows: int, K: int, op_dtype: int, device) -> torch.Tensor:
    return torch.zeros(rows, op_cols(K, op_dtype), dtype=OP_TORCH_DTYPE[op_dtype], device=device)


def pack_operand_into(src: torch.Tensor, dst: torch.Tensor, op_dtype: int) -> None:
    """src [N, K] FP32 device rows (last dim contiguous) -> dst operand buffer [N, op_cols(K)] (one ``pack_operand_kernel`` launch;
    padding columns are zeroed by the kernel)."""
    N, K = src.shape
    assert src.is_cuda and src.dtype == torch.float32 and src.stride(1) == 1 and dst.shape[0] == N and dst.stride(1) == 1
    lib = A.load(op_dtype)
    with _Span("pack_operand", 0, N * K * (4 + (4 if op_dtype in SPLIT_DTYPES else (4 if op_dtype == F32 else 2)))):
        A.check(lib.unav_pack_operand(_p(src), src.stride(0), _p(dst), dst.stride(0), N, K, op_dtype, _stream()), "unav_pack_operand")


def pack_operand(w: torch.Tensor, op_dtype: int) -> torch.Tensor:
    """[N, K] FP32 -> operand buffer (weight packing at load time; not on the per-batch path)."""
    N, K = w.shape
    w = w.detach()
    if w.dtype != torch.float32 or w.stride(1) != 1:
        w = w.float().contiguous()
    out = torch.empty(N, op_cols(K, op_dtype), dtype=OP_TORCH_DTYPE[op_dtype], device=w.device)
    pack_operand_into(w, out, op_dtype)
    return out


class View:
    """A column window of an operand / FP32 buffer: base tensor, first column, (logical) width."""

    __slots__ = ("t", "col", "width")

    def __init__(self, t: torch.Tensor, col: int = 0, width: Optional[int] = None):
        self.t, self.col = t, col
        self.width = width if width is not None else t.shape[1] - col

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.col * self.t.element_size()

    @property
    def ld(self) -> int:
        return self.t.stride(0)

    def rows(self, r0: int, r1: Optional[int] = None) -> "View":
        return View(self.t[r0:r1], self.col, self.width)


def _vp(v) -> Optional[int]:
    if v is None:
        return None
    return v.ptr if isinstance(v, View) else _p(v)


def _vld(v) -> int:
    if v is None:
        return 0
    return v.ld if isinstance(v, View) else _ld(v)


# ------------------------------------------------------------------------------------------- GEMM
def gemm(groups: Sequence[dict], M: int, N: int, K: int, op_dtype: int, act: int = ACT_NONE,
         res_masked: bool = False, backend: int = GEMM_TCGEN05, passes: int = 0) -> None:
    """groups: dicts with keys A, W (required) and bias, rowmask, rowscale, gate, gate_groups, gate_width,
    res, colscale, out_f32, out_op (tensors or Views).  passes: MMA passes over split operands (0 = all 3)."""
    n = len(groups)
    arr = (A.GemmGroup * n)()
    for i, g in enumerate(groups):
        s = arr[i]
        s.A, s.lda = _vp(g["A"]), _vld(g["A"])
        s.W, s.ldw = _vp(g["W"]), _vld(g["W"])
        s.bias = _vp(g.get("bias"))
        s.rowmask = _vp(g.get("rowmask"))
        s.rowscale = _vp(g.get("rowscale"))
        s.gate = _vp(g.get("gate"))
        s.gate_groups = g.get("gate_groups", 1)
        s.gate_width = g.get("gate_width", 1)
        s.res, s.ldres = _vp(g.get("res")), _vld(g.get("res"))
        s.colscale = _vp(g.get("colscale"))
        s.out_f32, s.ld_f32 = _vp(g.get("out_f32")), _vld(g.get("out_f32"))
        s.out_op, s.ld_op = _vp(g.get("out_op")), _vld(g.get("out_op"))
        s.out_opT, s.ld_opT = _vp(g.get("out_opT")), _vld(g.get("out_opT"))
        s.t_seg, s.t_col0, s.t_ncols = g.get("t_seg", 0), g.get("t_col0", 0), g.get("t_ncols", 0)
        s.conv_T = g.get("conv_T", 0)
    lib = A.load(op_dtype)
    es = 4 if op_dtype == F32 else (4 if op_dtype in SPLIT_DTYPES else 2)
    es_in = 2 if (op_dtype in SPLIT_DTYPES and passes == 1) else es     # one pass reads the hi halves only
    out_b = sum((4 if g.get("out_f32") is not None else 0) + (es if g.get("out_op") is not None else 0) for g in groups)
    with _Span("gemm_tcgen05" if backend == GEMM_TCGEN05 else "gemm_simt_kernel", 2.0 * M * N * K * n,
               n * (M * K + N * K) * es_in + M * N * out_b, f"{n}x[{M},{N},{K}]" + (f"p{passes}" if passes else "")) as sp:
        A.check(lib.unav_gemm(arr, n, M, N, K, with_passes(op_dtype, passes), act, int(res_masked), backend, _stream()),
                "unav_gemm")
        if _TRACE is not None and backend == GEMM_TCGEN05:      # the library chose the tile: name the span after the kernel
            sp.rename(GEMM_KERNELS.get(lib.unav_gemm_last_variant(), "gemm_tcgen05"))


# -------------------------------------------------------------------------------------- LayerNorm
def layernorm_rows(groups: Sequence[dict], M: int, C_: int, op_dtype: int, act: int = ACT_NONE,
                   eps: float = 1e-5) -> None:
    n = len(groups)
    arr = (A.LnGroup * n)()
    for i, g in enumerate(groups):
        s = arr[i]
        s.x, s.ldx = _vp(g["x"]), _vld(g["x"])
        s.add, s.ldadd = _vp(g.get("add")), _vld(g.get("add"))
        s.w, s.b = _p(g["w"]), _p(g["b"])
        s.post = _p(g.get("post"))
        s.post_rows = g.get("post_rows", 0)
        s.rowmask = _p(g.get("rowmask"))
        s.edge = _p(g.get("edge"))
        s.out_f32, s.ld_f32 = _vp(g.get("out_f32")), _vld(g.get("out_f32"))
        s.out_op, s.ld_op = _vp(g.get("out_op")), _vld(g.get("out_op"))
        s.out_im2col, s.ld_im2col = _vp(g.get("out_im2col")), _vld(g.get("out_im2col"))
        s.x_seg_rows = g.get("x_seg_rows", 0)
        s.x_seg_stride = g.get("x_seg_stride", 0)
        s.x_row_off = g.get("x_row_off", 0)
    lib = A.load(op_dtype)
    with _Span("layernorm_rows", 8.0 * M * C_ * n, n * M * C_ * 8, f"{n}x[{M},{C_}]"):
        A.check(lib.unav_layernorm_rows(arr, n, M, C_, eps, act, op_dtype, _stream()), "unav_layernorm_rows")


# ---------------------------------------------------------------------------- depthwise conv + LN
def dwconv_ln(groups: Sequence[dict], nseg: int, seg_len_in: int, stride: int, C_: int, op_dtype: int,
              eps: float = 1e-5) -> None:
    """groups: dicts with x, mask_out, pre (list of (w, b), <= 2) and outs (list of dicts: dw, ln_w, ln_b,
    src, out_f32, out_op)."""
    n = len(groups)
    arr = (A.DwLnGroup * n)()
    n_pre = len(groups[0].get("pre", []))
    n_out = len(groups[0]["outs"])
    for i, g in enumerate(groups):
        s = arr[i]
        s.x, s.ldx = _vp(g["x"]), _vld(g["x"])
        s.mask_out = _p(g.get("mask_out"))
        for j, (w, b) in enumerate(g.get("pre", [])):
            s.pre_w[j], s.pre_b[j] = _p(w), _p(b)
        assert len(g["outs"]) == n_out and len(g.get("pre", [])) == n_pre
        for j, o in enumerate(g["outs"]):
            d = s.out[j]
            d.dw, d.ln_w, d.ln_b = _p(o["dw"]), _p(o.get("ln_w")), _p(o.get("ln_b"))
            d.src = o.get("src", -1)
            d.out_f32, d.ld_f32 = _vp(o.get("out_f32")), _vld(o.get("out_f32"))
            d.out_op, d.ld_op = _vp(o.get("out_op")), _vld(o.get("out_op"))
    lib = A.load(op_dtype)
    rows = nseg * (seg_len_in // stride)
    with _Span("dwconv_ln", 20.0 * rows * C_ * n * n_out, n * rows * C_ * (4 * stride + 4 * n_out), f"{n}x[{rows},{C_}]x{n_out}"):
        A.check(lib.unav_dwconv_ln(arr, n, nseg, seg_len_in, stride, C_, n_pre, n_out, eps, op_dtype, _stream()),
                "unav_dwconv_ln")


# -------------------------------------------------------------------------------------- attention
def attention(groups: Sequence[dict], nb: int, Tq: int, Tk: int, nh: int, hs: int, scale: float,
              op_dtype: int) -> None:
    n = len(groups)
    arr = (A.AttnGroup * n)()
    for i, g in enumerate(groups):
        s = arr[i]
        s.q, s.ldq = _vp(g["q"]), _vld(g["q"])
        s.k, s.ldk = _vp(g["k"]), _vld(g["k"])
        s.v, s.ldv = _vp(g["v"]), _vld(g["v"])
        s.kmask = _p(g.get("kmask"))
        s.xk, s.xv = _vp(g.get("xk")), _vp(g.get("xv"))
        s.ldx = _vld(g.get("xk"))
        s.x_first = g.get("x_first", 0)
        s.out, s.ldo = _vp(g["out"]), _vld(g["out"])
    lib = A.load(op_dtype)
    with _Span("attention", 4.0 * n * nb * nh * Tq * Tk * hs, n * nb * nh * hs * (Tq * 8 + Tk * 8), f"{n}x[{nb},{nh},{Tq},{Tk},{hs}]"):
        A.check(lib.unav_attention(arr, n, nb, Tq, Tk, nh, hs, scale, op_dtype, _stream()), "unav_attention")


def attention_tc_workspace_bytes(ngroups: int, nb: int, Tq: int, Tk: int, nh: int, hs: int) -> int:
    return int(A.load().unav_attention_tc_workspace_bytes(ngroups, nb, Tq, Tk, nh, hs))


def attention_tc(groups: Sequence[dict], nb: int, Tq: int, Tk: int, nh: int, hs: int, scale: float, op_dtype: int,
                 passes: int = 0, workspace=None) -> None:
    """tcgen05 attention: groups carry q, k (operand rows), vt (operand, transposed values), kmask, out and optionally
    q32 / xk / xv (FP32 rows) + x_first for the per-query extra key.  Tk > 256 (config 4) runs key-chunked and needs
    ``workspace`` (uint8 tensor of ``attention_tc_workspace_bytes``)."""
    n = len(groups)
    arr = (A.AttnTcGroup * n)()
    for i, g in enumerate(groups):
        s = arr[i]
        s.q, s.ldq = _vp(g["q"]), _vld(g["q"])
        s.k, s.ldk = _vp(g["k"]), _vld(g["k"])
        s.vt, s.ldvt = _vp(g["vt"]), _vld(g["vt"])
        s.kmask = _p(g.get("kmask"))
        s.q32, s.ldq32 = _vp(g.get("q32")), _vld(g.get("q32"))
        s.xk, s.xv, s.ldx = _vp(g.get("xk")), _vp(g.get("xv")), _vld(g.get("xk"))
        s.x_first = g.get("x_first", 0)
        s.out, s.ldo = _vp(g["out"]), _vld(g["out"])
        s.qmask = _p(g.get("qmask"))
    lib = A.load(op_dtype)
    with _Span("attention_tc", 4.0 * n * nb * nh * Tq * Tk * hs, n * nb * nh * hs * (Tq * 8 + Tk * 8), f"{n}x[{nb},{nh},{Tq},{Tk},{hs}]"):
        if Tk > 256:
            assert workspace is not None, "attention_tc: Tk > 256 needs a workspace"
            A.check(lib.unav_attention_tc_long(arr, n, nb, Tq, Tk, nh, hs, scale, with_passes(op_dtype, passes), _p(workspace),
                                               workspace.numel() * workspace.element_size(), _stream()), "unav_attention_tc_long")
        else:
            A.check(lib.unav_attention_tc(arr, n, nb, Tq, Tk, nh, hs, scale, with_passes(op_dtype, passes), _stream()),
                    "unav_attention_tc")


def maxsig_gate(x, G, head_bias, gate, nb: int, T: int, nwords: int, H: int, hc: int) -> None:
    lib = A.load()
    with _Span("maxsig_gate", 2.0 * nb * T * nwords * H * hc, nb * (T + nwords) * H * hc * 4, f"[{nb},{T},{nwords},{H}x{hc}]"):
        A.check(lib.unav_maxsig_gate(_vp(x), _vld(x), _vp(G), _vld(G), _p(head_bias), _p(gate), nb, T, nwords, H,
                                     hc, _stream()), "unav_maxsig_gate")


def maxsig_gate_tc(x_buf, x_col0: int, G_buf, g_col0: int, head_bias, gate, nb: int, T: int, nwords: int, H: int, hc: int,
                   op_dtype: int, passes: int = 0) -> None:
    """tcgen05 MaxSigmoid gate: x_buf / G_buf are whole operand buffers, x_col0 / g_col0 the first column of the window."""
    lib = A.load(op_dtype)
    with _Span("maxsig_gate_tc", 2.0 * nb * T * nwords * H * hc, nb * (T + nwords) * H * hc * 4, f"[{nb},{T},{nwords},{H}x{hc}]"):
        A.check(lib.unav_maxsig_gate_tc(_p(x_buf), _ld(x_buf), x_col0, _p(G_buf), _ld(G_buf), g_col0, _p(head_bias), _p(gate),
                                        nb, T, nwords, H, hc, with_passes(op_dtype, passes), _stream()), "unav_maxsig_gate_tc")


def pool_match(u0, u1, u2, T0: int, T1: int, T2: int, Wm, bm, q, nb: int, C_: int, Tq: int, P: int = 4) -> None:
    lib = A.load()
    with _Span("pool_match", 2.0 * nb * Tq * 3 * P * C_, nb * C_ * (T0 + T1 + T2 + Tq) * 4):
        A.check(lib.unav_pool_match(_p(u0), _p(u1), _p(u2), T0, T1, T2, _ld(u0), _p(Wm), _p(bm), _p(q), _ld(q), nb,
                                    C_, Tq, P, _stream()), "unav_pool_match")


def rowcopy(jobs: Sequence[dict], op_dtype: int) -> None:
    n = len(jobs)
    arr = (A.CopyJob * n)()
    for i, j in enumerate(jobs):
        s = arr[i]
        s.src, s.ld_src = _vp(j["src"]), _vld(j["src"])
        s.dst, s.ld_dst = _vp(j["dst"]), _vld(j["dst"])
        s.nseg, s.seg_len_in, s.seg_len_out = j["nseg"], j["seg_len_in"], j["seg_len_out"]
        s.dst_seg_stride = j.get("dst_seg_stride", j["seg_len_out"])
        s.dst_row_off = j.get("dst_row_off", 0)
        s.num, s.den = j.get("num", 1), j.get("den", 1)
        s.ntaps = j.get("ntaps", 1)
        s.tap_stride = j.get("tap_stride", j["C"])
        s.C = j["C"]
    lib = A.load(op_dtype)
    es = 4 if op_dtype in (F32, BF16X2, F16X2) else 2
    nbytes = sum(j["nseg"] * j["seg_len_out"] * j.get("ntaps", 1) * j["C"] * (4 + es) for j in jobs)
    with _Span("rowcopy", 0, nbytes, f"{n} jobs"):
        A.check(lib.unav_rowcopy(arr, n, op_dtype, _stream()), "unav_rowcopy")


def transpose_cast(inp, ld_in: int, out, nb: int, R: int, Cc: int, op_dtype: int) -> None:
    lib = A.load(op_dtype)
    with _Span("transpose_cast", 0, nb * R * Cc * 8, f"[{nb},{R},{Cc}]"):
        A.check(lib.unav_transpose_cast(_vp(inp), ld_in, _vp(out), _vld(out), nb, R, Cc, op_dtype, _stream()),
                "unav_transpose_cast")


def align_embed(x0, cls_v, cls_a, pos_v, pos_a, type_v, type_a, tokens, nb: int, T: int, C_: int) -> None:
    lib = A.load()
    with _Span("align_embed", 0, 2 * nb * (T + 1) * C_ * 12):
        A.check(lib.unav_align_embed(_p(x0), _p(cls_v), _p(cls_a), _p(pos_v), _p(pos_a), _p(type_v), _p(type_a),
                                     _p(tokens), nb, T, C_, _stream()), "unav_align_embed")


def collate_pad(ragged, offsets, lens, out, mask, B: int, C_: int, T: int, pad: float = 0.0) -> None:
    """ragged f32 (videos' [C, len] blocks back to back), offsets i64 [B], lens i32 [B] -> out [B, C, T] f32, mask [B, T] u8."""
    lib = A.load()
    with _Span("collate_pad", 0, B * C_ * T * 8):
        A.check(lib.unav_collate_pad(_p(ragged), _p(offsets), _p(lens), _p(out), _p(mask), B, C_, T, pad, _stream()),
                "unav_collate_pad")


def build_masks(mask, out_true, out_up, out_cls, out_heads, nb: int, nb_src: int, T: int, L: int) -> None:
    lib = A.load()
    with _Span("build_masks", 0, nb * T * 4):
        A.check(lib.unav_build_masks(_p(mask), _p(out_true), _p(out_up), _p(out_cls), _p(out_heads), nb, nb_src, T, L,
                                     _stream()), "unav_build_masks")


def decode(logits, offsets, masks, points, level_off: List[int], B: int, ncls: int, class_aware: bool,
           pre_nms_thresh: float, pre_nms_topk: int, duration_thresh: float, cand_segs, cand_scores,
           cand_labels, cap: int) -> None:
    L = len(level_off) - 1
    arr = (C.c_int * (L + 1))(*level_off)
    lib = A.load()
    rows = level_off[-1]
    # algorithmic bytes: logits + offsets in, 20 B per candidate slot out
    with _Span("decode", 0, B * (rows * ncls * 12 + cap * 20), f"[{B},{rows},{ncls}]"):
        A.check(lib.unav_decode(_p(logits), _p(offsets), _p(masks), _p(points), arr, B, L, ncls, int(class_aware),
                                pre_nms_thresh, pre_nms_topk, duration_thresh, _p(cand_segs), _p(cand_scores),
                                _p(cand_labels), cap, _stream()), "unav_decode")


def softnms_workspace_bytes(B: int, ncls: int, max_seg_num: int) -> int:
    return int(A.load().unav_softnms_workspace_bytes(B, ncls, max_seg_num))


def softnms_batched(cand_segs, cand_scores, cand_labels, B: int, cap: int, ncls: int, iou_threshold: float,
                    sigma: float, min_score: float, method: int, max_seg_num: int, max_per_class: int,
                    vid_meta, out_segs, out_scores, out_labels, out_counts, workspace) -> None:
    lib = A.load()
    # algorithmic bytes: 20 B per candidate in (seg 8 + score 4 + label 8) + 20 B per kept detection
    with _Span("softnms", 0, B * (cap * 20 + max_seg_num * 20), f"[{B},{cap},{ncls}]"):
      A.check(lib.unav_softnms_batched(_p(cand_segs), _p(cand_scores), _p(cand_labels), B, cap, ncls,
                                     iou_threshold, sigma, min_score, method, max_seg_num, max_per_class,
                                     _p(vid_meta), _p(out_segs), _p(out_scores), _p(out_labels),
                                     _p(out_counts), _p(workspace), workspace.numel() * workspace.element_size(),
                                     _stream()), "unav_softnms_batched")


def gemm_variant_counts() -> dict:
    """Cumulative tcgen05 GEMM launches per kernel variant (names as ncu lists them), summed over the loaded libraries."""
    tot = [0] * 8
    for lib in A.loaded():
        buf = (C.c_longlong * 8)()
        A.check(lib.unav_gemm_variant_counts(buf, 8), "unav_gemm_variant_counts")
        tot = [a + int(b) for a, b in zip(tot, buf)]
    return {GEMM_KERNELS.get(i, f"variant{i}"): n for i, n in enumerate(tot)}


def launch_count() -> int:
    return sum(int(lib.unav_launch_count()) for lib in A.loaded())
