"""mAP evaluator with the matching on the device (unav_yolyolva_b200/utils/metrics.py, unav_map_match) vs the reference's
ANETdetection (golden, bit-exact) and vs the numpy oracle on a larger random case."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import map_ref
from unav_yolyolva_b200.utils import ANETdetection

pytestmark = pytest.mark.gpu


def _write_json(tmp_path, db):
    jf = os.path.join(tmp_path, "ants.json")
    json.dump({"database": db}, open(jf, "w"))
    return jf


def test_anetdetection_matches_reference_golden(golden_dir, tmp_path):
    d = np.load(os.path.join(golden_dir, "map_case.npz"))
    jf = os.path.join(tmp_path, "ants.json")
    open(jf, "w").write(str(d["json"]))
    ev = ANETdetection(jf, "test", tiou_thresholds=d["tious"], num_workers=1, device="cuda:0")
    preds = {"video-id": [str(v) for v in d["video_id"]], "t-start": torch.from_numpy(d["t_start"]),
             "t-end": torch.from_numpy(d["t_end"]), "label": torch.from_numpy(d["label"]), "score": torch.from_numpy(d["score"])}
    mAP, avg = ev.evaluate(preds, verbose=False)
    assert np.array_equal(ev.ap, d["ap"]), np.abs(ev.ap - d["ap"]).max()
    assert np.array_equal(mAP, d["mAP"]) and avg == float(d["average_mAP"])


def test_anetdetection_large_random_case_vs_oracle(tmp_path):
    """~60 k detections over 300 videos and 100 classes (the size of an UnAV-100 test split at 200 detections per
    video): device matching vs the numpy restatement of the reference loop, bit-exact AP matrix."""
    rng = np.random.default_rng(3)
    nvid, ncls, per = 300, 100, 200
    db, pv, p0, p1, pl, ps = {}, [], [], [], [], []
    for v in range(nvid):
        dur = float(rng.uniform(20, 60))
        ants = []
        for _ in range(int(rng.integers(1, 12))):
            s = float(rng.uniform(0, dur - 2)); e = float(min(dur, s + rng.uniform(0.5, 15)))
            lab = int(rng.integers(0, ncls))
            ants.append({"segment": [s, e], "label_id": lab, "label": str(lab)})
        db[f"v{v}"] = {"subset": "test", "duration": dur, "annotations": ants}
        for k in range(per):
            a = ants[k % len(ants)]
            w = a["segment"][1] - a["segment"][0]
            if k < 3 * len(ants):
                s, e = a["segment"][0] + rng.normal(0, 0.2) * w, a["segment"][1] + rng.normal(0, 0.2) * w
                lab = a["label_id"]
            else:
                s = float(rng.uniform(0, dur - 1)); e = s + float(rng.uniform(0.3, 10)); lab = int(rng.integers(0, ncls))
            pv.append(f"v{v}"); p0.append(np.float32(max(0.0, min(s, e)))); p1.append(np.float32(max(s, e))); pl.append(lab)
            ps.append(np.float32(rng.random()))
    jf = _write_json(tmp_path, db)
    tious = np.linspace(0.1, 0.9, 9)
    ev = ANETdetection(jf, "test", tiou_thresholds=tious, device="cuda:0")
    preds = {"video-id": pv, "t-start": np.array(p0), "t-end": np.array(p1), "label": np.array(pl), "score": np.array(ps)}
    mAP, avg = ev.evaluate(preds, verbose=False)
    # oracle on the same columns
    gt = ev.ground_truth
    vid = {}
    gv = np.array([vid.setdefault(v, len(vid)) for v in gt["video-id"]])
    pvi = np.array([vid.setdefault(v, len(vid)) for v in pv])
    gl = np.array([ev.activity_index[l] for l in gt["label"]])
    pli = np.array([ev.activity_index.get(int(l), int(l)) for l in pl])
    ap = map_ref.average_precision(gv, np.stack([gt["t-start"], gt["t-end"]], 1).astype(np.float64), gl, pvi,
                                   np.stack([np.array(p0), np.array(p1)], 1).astype(np.float64), pli,
                                   np.array(ps).astype(np.float64), tious, len(ev.activity_index))
    assert np.array_equal(ev.ap, ap), np.abs(ev.ap - ap).max()
    assert 0.0 < avg < 1.0


def test_anetdetection_needs_cuda(golden_dir, tmp_path):
    d = np.load(os.path.join(golden_dir, "map_case.npz"))
    jf = os.path.join(tmp_path, "ants.json")
    open(jf, "w").write(str(d["json"]))
    ev = ANETdetection(jf, "test", tiou_thresholds=d["tious"], device="cpu")
    with pytest.raises(RuntimeError):
        ev.evaluate({"video-id": ["vid_001"], "t-start": np.array([1.0]), "t-end": np.array([2.0]), "label": np.array([2]),
                     "score": np.array([0.5])}, verbose=False)
