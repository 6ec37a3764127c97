"""ctypes wrapper of oracle/nms_ref.c — TEST INFRASTRUCTURE ONLY (see the C file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nms_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libnms_ref.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", SRC, "-o", SO, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int64)
        _lib.softnms_ref.restype = C.c_int
        _lib.softnms_ref.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, fp, ip]
        _lib.batched_nms_ref.restype = C.c_int
        _lib.batched_nms_ref.argtypes = [fp, fp, ip, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float,
                                         fp, fp, ip, ip]
        _lib.to_seconds_ref.restype = None
        _lib.to_seconds_ref.argtypes = [fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def softnms(segs, scores, iou_threshold, sigma, min_score, method=2):
    segs = np.ascontiguousarray(segs, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    n = scores.shape[0]
    dets = np.zeros((max(n, 1), 3), np.float32)
    inds = np.zeros(max(n, 1), np.int64)
    k = lib().softnms_ref(_f(segs), _f(scores), n, iou_threshold, sigma, min_score, method, _f(dets), _i(inds))
    return dets[:k], inds[:k]


def batched_nms(segs, scores, labels, iou_threshold, min_score, max_seg_num, use_soft_nms=True, sigma=0.5):
    """Returns (segs [k,2], scores [k], labels [k], src [k]) as numpy arrays."""
    segs = np.ascontiguousarray(segs, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    n = scores.shape[0]
    o_segs = np.zeros((max_seg_num, 2), np.float32)
    o_scores = np.zeros(max_seg_num, np.float32)
    o_labels = np.zeros(max_seg_num, np.int64)
    o_src = np.zeros(max_seg_num, np.int64)
    k = lib().batched_nms_ref(_f(segs), _f(scores), _i(labels), n, iou_threshold, min_score, max_seg_num,
                              int(use_soft_nms), sigma, _f(o_segs), _f(o_scores), _i(o_labels), _i(o_src))
    return o_segs[:k], o_scores[:k], o_labels[:k], o_src[:k]


def to_seconds(segs, stride, nframes, fps, duration):
    segs = np.ascontiguousarray(segs, np.float32).copy()
    lib().to_seconds_ref(_f(segs), segs.size, stride, nframes, fps, duration)
    return segs
