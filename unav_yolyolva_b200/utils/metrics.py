"""Detection mAP with the matching on the device.

``ANETdetection`` keeps the reference's constructor and ``evaluate`` contract
(/root/reference/libs/utils/metrics.py:89-150, :257-305): ground truth from the annotation JSON (``database`` ->
``subset`` / ``annotations`` with ``segment`` and ``label_id``, duplicates within 1e-3 s removed, :13-30), label ids
re-indexed to 0..C-1 in sorted order (:121-123), predictions as a dict of arrays / tensors, a pandas DataFrame or a JSON
file; returns ``(mAP per tIoU, average mAP)`` and leaves the per-class matrix in ``self.ap``.

What changes is how the work is done: the reference walks every class's detections with pandas ``iterrows`` under joblib
(:340-398); here the detections are ranked and grouped on the host with numpy (vectorised), ONE launch of
``unav_map_match`` does the greedy tIoU matching of every (class, video, threshold) chain in FP64, and the interpolated
precision/recall integral (:440-453) is evaluated with the reference's exact array expressions, so ``self.ap`` is
bit-identical to the reference's (same ``numpy.argsort`` ranking, hence the same order of equal scores).
There is no CPU fallback for the matching: a CUDA device is required.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np
import torch

from .. import _cabi as A


def remove_duplicate_annotations(ants, tol=1e-3):
    """metrics.py:13-30."""
    valid = []
    for ev in ants:
        s, e, l = ev["segment"][0], ev["segment"][1], ev["label_id"]
        if not any(abs(s - p["segment"][0]) <= tol and abs(e - p["segment"][1]) <= tol and l == p["label_id"] for p in valid):
            valid.append(ev)
    return valid


def _label_id(ev, label, label_offset):
    """metrics.py:47-53 (the tuple / list branch is restated literally, including its ``label_offset**i + x`` sum)."""
    if isinstance(ev[label], (tuple, list)):
        return sum(label_offset ** i + int(x) for i, x in enumerate(ev[label][::-1]))
    return int(ev[label])


def load_gt_seg_from_json(json_file, split=None, label="label_id", label_offset=0):
    """metrics.py:32-66, as column lists instead of a DataFrame."""
    with open(json_file, "r", encoding="utf8") as f:
        db = json.load(f)["database"]
    vids, starts, stops, labels = [], [], [], []
    for k, v in db.items():
        if split is not None and v["subset"].lower() != split:
            continue
        for ev in remove_duplicate_annotations(v["annotations"]):
            vids.append(k)
            starts.append(float(ev["segment"][0]))
            stops.append(float(ev["segment"][1]))
            labels.append(_label_id(ev, label, label_offset))
    return {"video-id": vids, "t-start": starts, "t-end": stops, "label": labels}


def load_pred_seg_from_json(json_file, label="label_id", label_offset=0):
    """metrics.py:68-99."""
    with open(json_file, "r", encoding="utf8") as f:
        db = json.load(f)["database"]
    out = {"video-id": [], "t-start": [], "t-end": [], "label": [], "score": []}
    for k, v in db.items():
        for ev in v:
            out["video-id"].append(k)
            out["t-start"].append(float(ev["segment"][0]))
            out["t-end"].append(float(ev["segment"][1]))
            out["label"].append(_label_id(ev, label, label_offset))
            out["score"].append(float(ev["scores"]))
    return out


def _f64(x):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.asarray(x if not isinstance(x, np.ndarray) else x.tolist(), dtype=np.float64)   # .tolist(): metrics.py:276-280


def interpolated_ap(prec: np.ndarray, rec: np.ndarray) -> float:
    """metrics.py:440-453 with the right-to-left running maximum vectorised (same values: max() does not round)."""
    mprec = np.hstack([[0], prec, [0]])
    mrec = np.hstack([[0], rec, [1]])
    mprec = np.maximum.accumulate(mprec[::-1])[::-1]
    idx = np.where(mrec[1::] != mrec[0:-1])[0] + 1
    return np.sum((mrec[idx] - mrec[idx - 1]) * mprec[idx])


class ANETdetection(object):
    def __init__(self, ant_file, split=None, model_name=None, tiou_thresholds=np.linspace(0.1, 0.5, 5), label="label_id",
                 label_offset=0, num_workers=8, dataset_name=None, device="cuda"):
        self.tiou_thresholds = np.asarray(tiou_thresholds, dtype=np.float64)
        self.ap = None
        self.num_workers = num_workers          # kept for signature compatibility; the device does the parallel part
        self.dataset_name = dataset_name if dataset_name is not None else os.path.basename(ant_file).replace(".json", "")
        self.split = split
        self.device = torch.device(device)
        gt = load_gt_seg_from_json(ant_file, split=self.split, label=label, label_offset=label_offset)
        self.activity_index = {j: i for i, j in enumerate(sorted(set(gt["label"])))}
        self.ground_truth = gt
        self._gt_label = np.array([self.activity_index[l] for l in gt["label"]], dtype=np.int64)
        self._gt_seg = (np.stack([np.asarray(gt["t-start"], dtype=np.float64), np.asarray(gt["t-end"], dtype=np.float64)], 1)
                        if gt["label"] else np.zeros((0, 2)))
        self._vid_index: Dict[str, int] = {}
        self._gt_video = np.array([self._vid_index.setdefault(v, len(self._vid_index)) for v in gt["video-id"]], dtype=np.int64)

    # ------------------------------------------------------------------------------------------------------------
    def _columns(self, preds) -> Tuple[List[str], np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        if isinstance(preds, str) and os.path.isfile(preds):
            preds = load_pred_seg_from_json(preds)
        elif not isinstance(preds, dict):       # a pandas DataFrame (metrics.py:267-268)
            assert "label" in preds
            preds = {k: preds[k].tolist() for k in ("video-id", "t-start", "t-end", "label", "score")}
        vids = list(preds["video-id"])
        lab = preds["label"]
        if torch.is_tensor(lab):
            lab = lab.detach().cpu().numpy()
        lab = np.asarray(lab).astype(np.int64)
        # `preds['label'].replace(self.activity_index)` (:287): ids without ground truth keep their raw value
        lab = np.array([self.activity_index.get(int(l), int(l)) for l in lab], dtype=np.int64)
        return vids, _f64(preds["t-start"]), _f64(preds["t-end"]), lab, _f64(preds["score"])

    def match(self, vids, t0, t1, lab, score):
        """Ranks and groups the detections, runs the device matching.  Returns per class c (only classes with ground
        truth and detections): (tp [nt, n_c] uint8 in descending-score order, npos)."""
        nt, ncls = len(self.tiou_thresholds), len(self.activity_index)
        vmap = dict(self._vid_index)
        pv = np.array([vmap.setdefault(v, len(vmap)) for v in vids], dtype=np.int64)
        V = max(len(vmap), 1)
        keep = (lab >= 0) & (lab < ncls)
        cls_rank = np.zeros(len(lab), dtype=np.int64)          # position in the class's descending-score order
        cls_start = np.zeros(ncls + 1, dtype=np.int64)
        for c in range(ncls):
            p = np.nonzero(lab == c)[0]
            order = score[p].argsort()[::-1]                    # metrics.py:338, the very same call
            cls_rank[p[order]] = np.arange(len(p))
            cls_start[c + 1] = cls_start[c] + len(p)
        sel = np.nonzero(keep)[0]
        if len(sel) == 0:
            return {}
        dkey = lab[sel] * V + pv[sel]
        o = np.lexsort((cls_rank[sel], dkey))                   # (class, video) groups, score order inside
        sel, dkey = sel[o], dkey[o]
        ukey, first = np.unique(dkey, return_index=True)
        det_ptr = np.append(first, len(sel)).astype(np.int32)
        gkey = self._gt_label * V + self._gt_video
        go = np.argsort(gkey, kind="stable")                    # original order inside a group
        gkey_s = gkey[go]
        gt_ptr = np.stack([np.searchsorted(gkey_s, ukey, "left"), np.searchsorted(gkey_s, ukey, "right")], 1).astype(np.int32)
        det_seg = np.stack([t0[sel], t1[sel]], 1)
        gt_seg = self._gt_seg[go]
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("ANETdetection matching runs on a CUDA device (no CPU fallback)")
        with torch.cuda.device(dev):
            d_det = torch.from_numpy(np.ascontiguousarray(det_seg)).to(dev)
            d_gt = torch.from_numpy(np.ascontiguousarray(gt_seg)).to(dev) if len(gt_seg) else torch.zeros(1, 2, dtype=torch.float64, device=dev)
            d_dp = torch.from_numpy(det_ptr).to(dev)
            d_gp = torch.from_numpy(np.ascontiguousarray(gt_ptr)).to(dev)
            d_th = torch.from_numpy(self.tiou_thresholds).to(dev)
            tp = torch.zeros(nt, len(sel), dtype=torch.uint8, device=dev)
            lock = torch.zeros(nt * max(len(gt_seg), 1), dtype=torch.uint8, device=dev)
            lib = A.load()
            A.check(lib.unav_map_match(d_det.data_ptr(), d_gt.data_ptr(), d_dp.data_ptr(), d_gp.data_ptr(), len(ukey), d_th.data_ptr(),
                                       nt, len(sel), len(gt_seg), tp.data_ptr(), lock.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "unav_map_match")
            tp = tp.cpu().numpy()
        out = {}
        npos = np.bincount(self._gt_label, minlength=ncls)
        for c in range(ncls):
            m = lab[sel] == c
            if not m.any():
                continue
            tpc = np.zeros((nt, int(m.sum())), dtype=np.uint8)
            tpc[:, cls_rank[sel[m]]] = tp[:, m]
            out[c] = (tpc, float(npos[c]))
        return out

    def wrapper_compute_average_precision(self, vids, t0, t1, lab, score):
        """metrics.py:154-170 + :399-407: ap [nt, ncls]."""
        nt = len(self.tiou_thresholds)
        ap = np.zeros((nt, len(self.activity_index)))
        for c, (tp, npos) in self.match(vids, t0, t1, lab, score).items():
            tpd = tp.astype(np.float64)
            tp_c = np.cumsum(tpd, axis=1).astype(np.single)
            fp_c = np.cumsum(1.0 - tpd, axis=1).astype(np.single)
            rec = tp_c / npos
            prec = tp_c / (tp_c + fp_c)
            for t in range(nt):
                ap[t, c] = interpolated_ap(prec[t, :], rec[t, :])
        return ap

    def evaluate(self, preds, verbose=True):
        """metrics.py:257-305."""
        self.ap = None
        self.ap = self.wrapper_compute_average_precision(*self._columns(preds))
        mAP = self.ap.mean(axis=1)
        average_mAP = mAP.mean()
        if verbose:
            print("[RESULTS] Action detection results on {:s}.".format(self.dataset_name))
            block = ""
            for tiou, tiou_mAP in zip(self.tiou_thresholds, mAP):
                block += "\n|tIoU = {:.2f}: mAP = {:.2f} (%)".format(tiou, tiou_mAP * 100)
            print(block)
            print("Avearge mAP: {:.2f} (%)".format(average_mAP * 100))
        return mAP, average_mAP
