"""ctypes binding of ``include/unav_b200.h`` (the C-ABI drop-in boundary).

The shared library ``csrc/libunav_b200.so`` is built in-tree by ``csrc/build.py`` (called from
``__graft_entry__.build()``).  There is NO fallback: if the library is missing or a call fails, a
``UnavError`` is raised — the product path never routes through PyTorch eager ops or the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libunav_b200.so")              # BF16 halves (default)
LIB_PATH_F16 = os.path.join(_HERE, "csrc", "libunav_b200_f16.so")      # FP16 halves (same sources, -DUNAV_HALF_F16)

F32, BF16, BF16X2, F16, F16X2 = 0, 1, 2, 3, 4
SPLIT_DTYPES = (BF16X2, F16X2)
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU = 0, 1, 2, 3
GEMM_SIMT, GEMM_TCGEN05 = 0, 1
MAX_GROUPS, MAX_COPY_JOBS = 8, 16

c_ll, c_vp, c_i, c_f = C.c_longlong, C.c_void_p, C.c_int, C.c_float


class UnavError(RuntimeError):
    pass


class GemmGroup(C.Structure):
    _fields_ = [("A", c_vp), ("lda", c_ll), ("W", c_vp), ("ldw", c_ll), ("bias", c_vp), ("rowmask", c_vp),
                ("rowscale", c_vp), ("gate", c_vp), ("res", c_vp), ("ldres", c_ll), ("colscale", c_vp),
                ("out_f32", c_vp), ("ld_f32", c_ll), ("out_op", c_vp), ("ld_op", c_ll),
                ("gate_groups", c_i), ("gate_width", c_i), ("out_opT", c_vp), ("ld_opT", c_ll), ("t_seg", c_i),
                ("t_col0", c_i), ("t_ncols", c_i), ("conv_T", c_i)]


class LnGroup(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_ll), ("add", c_vp), ("ldadd", c_ll), ("w", c_vp), ("b", c_vp),
                ("post", c_vp), ("rowmask", c_vp), ("edge", c_vp), ("out_f32", c_vp), ("ld_f32", c_ll),
                ("out_op", c_vp), ("ld_op", c_ll), ("out_im2col", c_vp), ("ld_im2col", c_ll),
                ("x_seg_rows", c_i), ("x_seg_stride", c_i), ("x_row_off", c_i), ("post_rows", c_i)]


class DwLnOut(C.Structure):
    _fields_ = [("dw", c_vp), ("ln_w", c_vp), ("ln_b", c_vp), ("out_f32", c_vp), ("ld_f32", c_ll),
                ("out_op", c_vp), ("ld_op", c_ll), ("src", c_i), ("pad_", c_i)]


class DwLnGroup(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_ll), ("mask_out", c_vp), ("pre_w", c_vp * 2), ("pre_b", c_vp * 2),
                ("out", DwLnOut * 3)]


class AttnGroup(C.Structure):
    _fields_ = [("q", c_vp), ("ldq", c_ll), ("k", c_vp), ("ldk", c_ll), ("v", c_vp), ("ldv", c_ll),
                ("kmask", c_vp), ("xk", c_vp), ("xv", c_vp), ("ldx", c_ll), ("out", c_vp), ("ldo", c_ll),
                ("x_first", c_i), ("pad_", c_i)]


class AttnTcGroup(C.Structure):
    _fields_ = [("q", c_vp), ("ldq", c_ll), ("k", c_vp), ("ldk", c_ll), ("vt", c_vp), ("ldvt", c_ll), ("kmask", c_vp),
                ("q32", c_vp), ("ldq32", c_ll), ("xk", c_vp), ("xv", c_vp), ("ldx", c_ll), ("out", c_vp), ("ldo", c_ll),
                ("x_first", c_i), ("pad_", c_i), ("qmask", c_vp)]


class CopyJob(C.Structure):
    _fields_ = [("src", c_vp), ("ld_src", c_ll), ("dst", c_vp), ("ld_dst", c_ll), ("nseg", c_i),
                ("seg_len_in", c_i), ("seg_len_out", c_i), ("dst_seg_stride", c_i), ("dst_row_off", c_i),
                ("num", c_i), ("den", c_i), ("ntaps", c_i), ("tap_stride", c_i), ("C", c_i)]


_PROTOS = {
    "unav_version": (C.c_char_p, []),
    "unav_last_error": (C.c_char_p, []),
    "unav_check_device": (c_i, [c_i]),
    "unav_launch_count": (c_ll, []),
    "unav_gemm_last_variant": (c_i, []),
    "unav_gemm_variant_counts": (c_i, [C.POINTER(c_ll), c_i]),
    "unav_set_phase_trace": (c_i, [c_vp, c_i]),
    "unav_set_pdl": (c_i, [c_i]),
    "unav_gemm": (c_i, [C.POINTER(GemmGroup), c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "unav_layernorm_rows": (c_i, [C.POINTER(LnGroup), c_i, c_i, c_i, c_f, c_i, c_i, c_vp]),
    "unav_dwconv_ln": (c_i, [C.POINTER(DwLnGroup), c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_vp]),
    "unav_attention": (c_i, [C.POINTER(AttnGroup), c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_vp]),
    "unav_attention_tc": (c_i, [C.POINTER(AttnTcGroup), c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_vp]),
    "unav_attention_tc_workspace_bytes": (C.c_size_t, [c_i, c_i, c_i, c_i, c_i, c_i]),
    "unav_attention_tc_long": (c_i, [C.POINTER(AttnTcGroup), c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_vp, C.c_size_t, c_vp]),
    "unav_maxsig_gate": (c_i, [c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "unav_maxsig_gate_tc": (c_i, [c_vp, c_ll, c_i, c_vp, c_ll, c_i, c_vp, c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "unav_pool_match": (c_i, [c_vp, c_vp, c_vp, c_i, c_i, c_i, c_ll, c_vp, c_vp, c_vp, c_ll, c_i, c_i, c_i, c_i, c_vp]),
    "unav_rowcopy": (c_i, [C.POINTER(CopyJob), c_i, c_i, c_vp]),
    "unav_transpose_cast": (c_i, [c_vp, c_ll, c_vp, c_ll, c_i, c_i, c_i, c_i, c_vp]),
    "unav_align_embed": (c_i, [c_vp] * 8 + [c_i, c_i, c_i, c_vp]),
    "unav_pack_operand": (c_i, [c_vp, c_ll, c_vp, c_ll, c_ll, c_i, c_i, c_vp]),
    "unav_map_match": (c_i, [c_vp, c_vp, c_vp, c_vp, c_i, c_vp, c_i, c_i, c_i, c_vp, c_vp, c_vp]),
    "unav_collate_pad": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_f, c_vp]),
    "unav_build_masks": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_vp]),
    "unav_decode": (c_i, [c_vp, c_vp, c_vp, c_vp, C.POINTER(c_i), c_i, c_i, c_i, c_i, c_f, c_i, c_f,
                          c_vp, c_vp, c_vp, c_i, c_vp]),
    "unav_softnms_workspace_bytes": (C.c_size_t, [c_i, c_i, c_i]),
    "unav_softnms_batched": (c_i, [c_vp, c_vp, c_vp, c_i, c_i, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_vp,
                                   c_vp, c_vp, c_vp, c_vp, c_vp, C.c_size_t, c_vp]),
}

EXPORTS = tuple(_PROTOS)
_libs = {False: None, True: None}


def load(op_dtype: int = F32) -> C.CDLL:
    """Load the library that serves `op_dtype` (once each) and attach prototypes: libunav_b200.so for F32 / BF16 /
    BF16X2, libunav_b200_f16.so for F16 / F16X2.  Raises UnavError if it is not built."""
    f16 = (op_dtype & 0xff) in (F16, F16X2)
    if _libs[f16] is not None:
        return _libs[f16]
    path = LIB_PATH_F16 if f16 else LIB_PATH
    if not os.path.exists(path):
        raise UnavError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / eager fallback for the hot path)")
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _libs[f16] = lib
    return lib


def loaded():
    """The libraries loaded so far (the BF16-halves one first)."""
    load()
    return [l for l in (_libs[False], _libs[True]) if l is not None]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = " | ".join(l.unav_last_error().decode(errors="replace") for l in loaded())
        raise UnavError(f"{what or 'unav call'} failed with code {rc}: {msg}")


def ptr(t) -> Optional[int]:
    """Device (or host) address of a torch tensor, None for None."""
    return None if t is None else t.data_ptr()
