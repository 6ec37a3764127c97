"""``batched_nms`` operator of the reference (/root/reference/libs/utils/nms.py:103-190), device resident.

The reference loops over classes in Python and calls a single-threaded CPU extension per class
(``nms_1d_cpu.softnms``, libs/utils/csrc/nms_cpu.cpp:67-160), forcing a device->host copy.  Here the whole
operator is one ``unav_softnms_batched`` call: one CUDA block per class, then a per-video merge.  Same
signature, same ordering contract (scores descending, ties: lower class id, earlier emission), bit-exact
indices/labels on tie-free inputs (tests/test_gpu_nms.py).
"""
from __future__ import annotations

import torch

from .. import kernels as K


def nms_method_code(nms_method: str, multiclass: bool) -> int:
    """The `method` argument of ``unav_softnms_batched`` for a test_cfg (2 = gaussian soft-NMS, 3 = hard NMS).  The ONE place
    both model entry points (the fused engine and ``PtTransformer.inference``) validate the NMS configuration: the branches of
    the reference that are off both shipped configs — ``nms_method='none'`` (multimodal_meta_archs.py:836 skips NMS and keeps
    every candidate) and class-agnostic NMS + seg voting (nms.py:67-101, :161-180) — raise instead of silently running
    per-class NMS."""
    if nms_method == "none":
        raise NotImplementedError("nms_method='none' (no NMS, all candidates returned) is not on the device path")
    if not multiclass:
        raise NotImplementedError("class-agnostic NMS + seg voting (libs/utils/nms.py:67-101, :161-180) is not on the device "
                                  "path: multiclass_nms=True in the reference configs")
    if nms_method not in ("soft", "hard"):
        raise ValueError(f"unknown nms_method {nms_method!r}")
    return 2 if nms_method == "soft" else 3


def batched_nms(segs, scores, cls_idxs, iou_threshold, min_score, max_seg_num, use_soft_nms=True,
                multiclass=True, sigma=0.5, voting_thresh=0.75):
    num_segs = segs.shape[0]
    if num_segs == 0:   # same corner case as nms.py:118-121
        return torch.zeros([0, 2]), torch.zeros([0, ]), torch.zeros([0, ], dtype=cls_idxs.dtype)
    if not torch.cuda.is_available():
        raise RuntimeError("batched_nms: no CUDA device; the B200 operator has no CPU fallback")
    in_dev = segs.device
    dev = in_dev if in_dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    segs_d = segs.detach().to(dev, torch.float32).contiguous()
    scores_d = scores.detach().to(dev, torch.float32).contiguous()
    method = nms_method_code("soft" if use_soft_nms else "hard", multiclass)
    labels_d = cls_idxs.detach().to(dev, torch.int32).contiguous()
    per_class = torch.bincount(labels_d.to(torch.int64))
    ncls = int(per_class.numel())
    # true per-class bound: the one-CTA-per-video kernel holds every candidate in shared memory (<= ~12.5 k); above that the
    # per-class kernels need only the largest class to fit (the reference op accepts any N: up to 19 900 candidates with its
    # default pre_nms_topk = 5000)
    max_per_class = int(per_class.max().item())
    if max_per_class * 16 > 200 * 1024:
        raise NotImplementedError(f"batched_nms: {max_per_class} candidates in one class exceed the shared-memory budget of "
                                  "the per-class kernel (12 800)")
    K_out = int(max_seg_num)
    out_segs = torch.empty(1, K_out, 2, dtype=torch.float32, device=dev)
    out_scores = torch.empty(1, K_out, dtype=torch.float32, device=dev)
    out_labels = torch.empty(1, K_out, dtype=torch.int64, device=dev)
    out_counts = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(K.softnms_workspace_bytes(1, ncls, K_out), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        K.softnms_batched(segs_d, scores_d, labels_d, 1, num_segs, ncls, float(iou_threshold), float(sigma),
                          float(min_score), method, K_out, max_per_class, None, out_segs, out_scores, out_labels, out_counts,
                          ws)
    n = int(out_counts.item())
    r_segs, r_scores, r_labels = out_segs[0, :n], out_scores[0, :n], out_labels[0, :n]
    return r_segs.to(in_dev), r_scores.to(in_dev), r_labels.to(in_dev, cls_idxs.dtype)
