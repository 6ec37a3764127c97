"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the reference's detection-mAP evaluator
(/root/reference/libs/utils/metrics.py): ``compute_average_precision_detection`` (:306-407), ``segment_iou``
(:417-437), ``interpolated_prec_rec`` (:440-453) and the per-class loop of ``wrapper_compute_average_precision``
(:154-170), without pandas / joblib.  Pinned against the reference's own ``ANETdetection.evaluate`` run in the build
container (tests/golden/map_case.npz, tests/golden/make_golden_map.py).  Only tests/ may import this module."""
import numpy as np


def segment_iou(target, cands):
    """:417-437 — float64, same operation order."""
    tt1 = np.maximum(target[0], cands[:, 0])
    tt2 = np.minimum(target[1], cands[:, 1])
    inter = (tt2 - tt1).clip(0)
    union = (cands[:, 1] - cands[:, 0]) + (target[1] - target[0]) - inter
    return inter.astype(float) / union


def interpolated_prec_rec(prec, rec):
    """:440-453."""
    mprec = np.hstack([[0], prec, [0]])
    mrec = np.hstack([[0], rec, [1]])
    for i in range(len(mprec) - 1)[::-1]:
        mprec[i] = max(mprec[i], mprec[i + 1])
    idx = np.where(mrec[1::] != mrec[0:-1])[0] + 1
    return np.sum((mrec[idx] - mrec[idx - 1]) * mprec[idx])


def class_ap(gt_video, gt_seg, pr_video, pr_seg, pr_score, tious):
    """One class (:306-407).  gt_video [G] int, gt_seg [G,2] f64, pr_video [P] int, pr_seg [P,2] f64, pr_score [P] f64.
    Returns (ap [nt], tp [nt, P] in the score order used, order [P])."""
    nt = len(tious)
    ap = np.zeros(nt)
    if len(pr_score) == 0:
        return ap, np.zeros((nt, 0)), np.zeros(0, dtype=np.int64)
    npos = float(len(gt_video))
    lock_gt = np.ones((nt, len(gt_video))) * -1
    order = pr_score.argsort()[::-1]                      # :338
    tp = np.zeros((nt, len(order)))
    fp = np.zeros((nt, len(order)))
    for idx, pi in enumerate(order):
        gi = np.nonzero(gt_video == pr_video[pi])[0]      # the video's ground truth of this class, original order
        if len(gi) == 0:                                  # :354-357
            fp[:, idx] = 1
            continue
        tiou = segment_iou(pr_seg[pi], gt_seg[gi])
        srt = tiou.argsort()[::-1]                        # :363
        for t, thr in enumerate(tious):
            for j in srt:
                if tiou[j] < thr:
                    fp[t, idx] = 1
                    break
                if lock_gt[t, gi[j]] >= 0:
                    continue
                tp[t, idx] = 1
                lock_gt[t, gi[j]] = idx
                break
            if fp[t, idx] == 0 and tp[t, idx] == 0:
                fp[t, idx] = 1
    tp_c = np.cumsum(tp, axis=1).astype(np.single)
    fp_c = np.cumsum(fp, axis=1).astype(np.single)
    rec = tp_c / npos
    prec = tp_c / (tp_c + fp_c)
    for t in range(nt):
        ap[t] = interpolated_prec_rec(prec[t, :], rec[t, :])
    return ap, tp, order


def average_precision(gt_video, gt_seg, gt_label, pr_video, pr_seg, pr_label, pr_score, tious, ncls):
    """ap [nt, ncls] as ``ANETdetection.wrapper_compute_average_precision`` fills it (labels already 0..ncls-1)."""
    ap = np.zeros((len(tious), ncls))
    for c in range(ncls):
        g, p = gt_label == c, pr_label == c
        ap[:, c] = class_ap(gt_video[g], gt_seg[g], pr_video[p], pr_seg[p], pr_score[p], tious)[0]
    return ap
