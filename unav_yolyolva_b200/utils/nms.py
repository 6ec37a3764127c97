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
    if not multiclass:
        raise NotImplementedError("class-agnostic NMS + seg voting (nms.py:67-101, :161-180) is not on the hot "
                                  "path: multiclass_nms=True in the reference configs")
    labels_d = cls_idxs.detach().to(dev, torch.int32).contiguous()
    ncls = int(cls_idxs.max().item()) + 1
    K_out = int(max_seg_num)
    out_segs = torch.empty(1, K_out, 2, dtype=torch.float32, device=dev)
    out_scores = torch.empty(1, K_out, dtype=torch.float32, device=dev)
    out_labels = torch.empty(1, K_out, dtype=torch.int64, device=dev)
    out_counts = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(K.softnms_workspace_bytes(1, ncls, K_out), dtype=torch.uint8, device=dev)
    method = 2 if use_soft_nms else 3
    with torch.cuda.device(dev):
        K.softnms_batched(segs_d, scores_d, labels_d, 1, num_segs, ncls, float(iou_threshold), float(sigma),
                          float(min_score), method, K_out, 0, None, out_segs, out_scores, out_labels, out_counts,
                          ws)
    n = int(out_counts.item())
    r_segs, r_scores, r_labels = out_segs[0, :n], out_scores[0, :n], out_labels[0, :n]
    return r_segs.to(in_dev), r_scores.to(in_dev), r_labels.to(in_dev, cls_idxs.dtype)
