"""Wall time of the mAP evaluation on a test-split-sized synthetic case (300 videos x 200 detections, 100 classes).
    python scripts/map_speed.py ours        (GPU box: unav_yolyolva_b200.utils.ANETdetection, matching on the device)
    python scripts/map_speed.py reference   (build container: the reference evaluator, pandas iterrows, num_workers=1)"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def case():
    rng = np.random.default_rng(3)
    nvid, ncls, per = 300, 100, 200
    db, pv, p0, p1, pl, ps = {}, [], [], [], [], []
    for v in range(nvid):
        dur = float(rng.uniform(20, 60)); ants = []
        for _ in range(int(rng.integers(1, 12))):
            s = float(rng.uniform(0, dur - 2)); e = float(min(dur, s + rng.uniform(0.5, 15))); lab = int(rng.integers(0, ncls))
            ants.append({"segment": [s, e], "label_id": lab, "label": str(lab)})
        db[f"v{v}"] = {"subset": "test", "duration": dur, "annotations": ants}
        for k in range(per):
            a = ants[k % len(ants)]; w = a["segment"][1] - a["segment"][0]
            if k < 3 * len(ants):
                s, e = a["segment"][0] + rng.normal(0, 0.2) * w, a["segment"][1] + rng.normal(0, 0.2) * w; lab = a["label_id"]
            else:
                s = float(rng.uniform(0, dur - 1)); e = s + float(rng.uniform(0.3, 10)); lab = int(rng.integers(0, ncls))
            pv.append(f"v{v}"); p0.append(np.float32(max(0.0, min(s, e)))); p1.append(np.float32(max(s, e))); pl.append(lab)
            ps.append(np.float32(rng.random()))
    jf = os.path.join(tempfile.mkdtemp(), "ants.json")
    json.dump({"database": db}, open(jf, "w"))
    return jf, {"video-id": pv, "t-start": np.array(p0), "t-end": np.array(p1), "label": np.array(pl), "score": np.array(ps)}

which = sys.argv[1] if len(sys.argv) > 1 else "ours"
jf, preds = case()
tious = np.linspace(0.1, 0.9, 9)
if which == "reference":
    from oracle.ref_harness import import_reference
    import_reference()
    from libs.utils import ANETdetection
    ev = ANETdetection(jf, "test", tiou_thresholds=tious, num_workers=1)   # joblib workers cannot unpickle the stubbed import path
else:
    from unav_yolyolva_b200.utils import ANETdetection
    ev = ANETdetection(jf, "test", tiou_thresholds=tious, device="cuda:0")
    ev.evaluate(dict(preds), verbose=False)          # warm-up (library load, context)
t0 = time.perf_counter()
mAP, avg = ev.evaluate(dict(preds), verbose=False)
dt = time.perf_counter() - t0
print(f"{which}: {len(preds['score'])} detections, evaluate() {dt * 1e3:.1f} ms, average mAP {avg * 100:.4f} %, checksum {float(ev.ap.sum()):.12f}")
