"""Golden fixture for the mAP evaluator: the REAL reference ``ANETdetection`` (libs/utils/metrics.py) on a synthetic
annotation file and synthetic detections.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_map.py        -> tests/golden/map_case.npz
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402


def make_case(seed=11, nvid=40, ncls=12, per_video=60):
    rng = np.random.default_rng(seed)
    db, preds = {}, {"video-id": [], "t-start": [], "t-end": [], "label": [], "score": []}
    for v in range(nvid):
        dur = float(rng.uniform(20, 60))
        ants = []
        for _ in range(int(rng.integers(1, 7))):
            s = float(rng.uniform(0, dur - 2))
            e = float(min(dur, s + rng.uniform(0.5, 15)))
            lab = int(rng.integers(0, ncls)) * 3 + 2            # non-contiguous label ids: exercises activity_index
            ants.append({"segment": [s, e], "label_id": lab, "label": str(lab)})
        if v % 7 == 0:
            ants.append(dict(ants[0]))                          # exact duplicate: removed by remove_duplicate_annotations
        subset = "test" if v % 5 else "train"                   # some videos outside the evaluated split
        db[f"vid_{v:03d}"] = {"subset": subset, "duration": dur, "annotations": ants}
        for k in range(per_video):
            if k < 2 * len(ants):                               # jittered copies of the ground truth, sometimes wrong class
                a = ants[k % len(ants)]
                w = a["segment"][1] - a["segment"][0]
                s = a["segment"][0] + rng.normal(0, 0.15) * w
                e = a["segment"][1] + rng.normal(0, 0.15) * w
                lab = a["label_id"] if rng.random() < 0.8 else int(rng.integers(0, ncls)) * 3 + 2
            else:
                s = float(rng.uniform(0, dur - 1)); e = s + float(rng.uniform(0.3, 10)); lab = int(rng.integers(0, ncls)) * 3 + 2
            s, e = float(np.float32(max(0.0, min(s, e)))), float(np.float32(max(s, e)))
            preds["video-id"].append(f"vid_{v:03d}")
            preds["t-start"].append(s); preds["t-end"].append(e); preds["label"].append(lab)
            preds["score"].append(float(np.float32(rng.random())))
    # a few exact score ties inside one class and one video nobody annotated
    preds["score"][5] = preds["score"][3]
    for k in range(3):
        preds["video-id"].append("vid_unknown"); preds["t-start"].append(1.0); preds["t-end"].append(2.0 + k)
        preds["label"].append(2); preds["score"].append(0.5)
    return {"database": db}, preds


def main():
    import_reference()
    from libs.utils import ANETdetection
    db, preds = make_case()
    tmp = tempfile.mkdtemp()
    jf = os.path.join(tmp, "synthetic_map.json")
    json.dump(db, open(jf, "w"))
    tious = np.linspace(0.1, 0.9, 9)
    ev = ANETdetection(jf, "test", tiou_thresholds=tious, num_workers=1)
    arr = {k: (np.array(v) if k != "video-id" else v) for k, v in preds.items()}
    mAP, avg = ev.evaluate({"video-id": arr["video-id"], "t-start": arr["t-start"], "t-end": arr["t-end"],
                            "label": arr["label"], "score": arr["score"]}, verbose=False)
    np.savez_compressed(os.path.join(HERE, "map_case.npz"), json=json.dumps(db), video_id=np.array(preds["video-id"]),
                        t_start=arr["t-start"], t_end=arr["t-end"], label=arr["label"], score=arr["score"], tious=tious,
                        ap=ev.ap, mAP=mAP, average_mAP=avg)
    print("wrote map_case.npz: ap", ev.ap.shape, "mAP", np.round(mAP * 100, 2), "avg", round(avg * 100, 3))


if __name__ == "__main__":
    main()
