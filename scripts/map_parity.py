"""mAP parity of the CUDA path vs the FP32 oracle with the REFERENCE's own evaluator (SURVEY.md §8d parity gate:
mAP at every tIoU 0.1..0.9 and both averages within 0.1 point).

Two stages, because the reference tree (libs/utils/metrics.py::ANETdetection) only exists in the build container:

  on the GPU box :  python scripts/map_parity.py dump  [--videos 64] [--mode bf16x3]   -> gpurun_out/map_dets.npz
  in the container: python scripts/map_parity.py eval                                  -> profiles/r01_map_parity.md

`dump` runs the same seeded synthetic videos through (a) the oracle restatement on the CPU (FP32) + the reference's
compiled NMS where available and (b) the B200 engine, and stores both detection sets.  `eval` builds a synthetic
annotation file whose ground truth is the oracle's detections ranked 1/4/9 per video with 5 % boundary jitter
(random GT gives mAP 0 everywhere) and scores both sets with ANETdetection.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "map_dets.npz")


def dump(args):
    import torch
    from oracle import model_ref as R
    from oracle import nms_ref
    from unav_yolyolva_b200 import synth
    from unav_yolyolva_b200.config import TEST_CFG, default_model_cfg
    from unav_yolyolva_b200.modeling import make_multimodal_meta_arch
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda", 0)
    sd = synth.trained_like_state_dict()
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    modes = args.mode.split(",")
    B = 16
    o_seg, o_sc, o_lb, durs = [], [], [], []
    g = {m: ([], [], []) for m in modes}
    pts = R.make_points(224)
    for first in range(0, args.videos, B):
        batch = synth.make_batch(B, 224, first_index=first, with_gt=False)
        for m in modes:
            model.precision = m
            res, _ = model(batch)
            g[m][0].append(res["segments"].cpu().numpy()); g[m][1].append(res["scores"].cpu().numpy()); g[m][2].append(res["labels"].cpu().numpy())
        with torch.no_grad():
            logits, offsets, masks = R.forward_logits(sd, batch["visual"], batch["audio"], batch["mask"])
        for i in range(B):
            segs, scores, labels, _ = R.decode_single_video(pts, [m[i] for m in masks], [x[i] for x in logits], [x[i] for x in offsets])
            r = nms_ref.batched_nms(segs.numpy(), scores.numpy(), labels.numpy(), TEST_CFG["iou_threshold"], TEST_CFG["min_score"],
                                    TEST_CFG["max_seg_num"], True, TEST_CFG["nms_sigma"])
            o_seg.append(nms_ref.to_seconds(r[0], batch["feat_stride"][i], batch["feat_num_frames"][i], batch["fps"][i], batch["duration"][i]))
            o_sc.append(r[1]); o_lb.append(r[2]); durs.append(batch["duration"][i])
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    arrays = dict(oracle_segs=np.stack(o_seg), oracle_scores=np.stack(o_sc), oracle_labels=np.stack(o_lb),
                  durations=np.array(durs), modes=np.array(modes))
    for m in modes:
        arrays[f"{m}_segs"], arrays[f"{m}_scores"], arrays[f"{m}_labels"] = (np.concatenate(x) for x in g[m])
        same = (arrays[f"{m}_labels"] == arrays["oracle_labels"]).mean()
        dseg = np.abs(arrays[f"{m}_segs"] - arrays["oracle_segs"])[arrays[f"{m}_labels"] == arrays["oracle_labels"]]
        print(f"mode {m}: detection labels identical in {same * 100:.2f} % of the {arrays['oracle_labels'].size} ranked slots; "
              f"max |segment diff| on those {dseg.max():.3e} s")
    np.savez_compressed(OUT, **arrays)
    print(f"dumped {len(durs)} videos")


def evaluate(args):
    import tempfile
    from oracle.ref_harness import import_reference
    import_reference()
    from libs.utils import ANETdetection
    d = np.load(OUT)
    n = d["oracle_segs"].shape[0]
    rng = np.random.default_rng(7)
    db = {}
    for v in range(n):
        ants = []
        for rank in (0, 3, 8):
            s, e = d["oracle_segs"][v, rank]
            w = max(e - s, 0.2)
            s2, e2 = s + rng.normal(0, 0.05) * w, e + rng.normal(0, 0.05) * w
            ants.append({"segment": [float(max(0.0, min(s2, e2))), float(max(s2, e2))], "label_id": int(d["oracle_labels"][v, rank]),
                         "label": str(int(d["oracle_labels"][v, rank]))})
        db[f"synth_{v:06d}"] = {"subset": "test", "duration": float(d["durations"][v]), "annotations": ants}
    tmp = tempfile.mkdtemp()
    jf = os.path.join(tmp, "synthetic_unav100.json")
    json.dump({"database": db}, open(jf, "w"))
    tious = np.linspace(0.1, 0.9, 9)

    def score(prefix):
        K = d[prefix + "_segs"].shape[1]
        res = {"video-id": [f"synth_{v:06d}" for v in range(n) for _ in range(K)],
               "t-start": d[prefix + "_segs"][..., 0].reshape(-1), "t-end": d[prefix + "_segs"][..., 1].reshape(-1),
               "label": d[prefix + "_labels"].reshape(-1), "score": d[prefix + "_scores"].reshape(-1)}
        ev = ANETdetection(jf, "test", tiou_thresholds=tious, num_workers=1)
        mAP, avg = ev.evaluate(res, verbose=False)
        return np.asarray(mAP) * 100.0

    m_o = score("oracle")
    modes = [str(m) for m in d["modes"]]
    lines = ["# Round 1 — mAP parity (reference evaluator `libs/utils/metrics.py::ANETdetection`)", "",
             f"{n} synthetic videos; ground truth = oracle detections ranked 1/4/9 per video with 5 % boundary jitter "
             "(`scripts/map_parity.py`).  Gate: |difference| <= 0.1 point at every tIoU and on both averages.", ""]
    ok = True
    for m in modes:
        m_g = score(m)
        same = (d[f"{m}_labels"] == d["oracle_labels"]).mean() * 100
        lines += [f"## engine mode `{m}` — labels identical to the oracle's in {same:.2f} % of the {d['oracle_labels'].size} ranked detection slots", "",
                  "| tIoU | oracle (CPU FP32) mAP % | B200 engine mAP % | difference |", "|---|---:|---:|---:|"]
        for t, a, b in zip(tious, m_o, m_g):
            lines.append(f"| {t:.1f} | {a:.3f} | {b:.3f} | {b - a:+.3f} |")
        lines.append(f"| avg 0.1:0.9 | {m_o.mean():.3f} | {m_g.mean():.3f} | {m_g.mean() - m_o.mean():+.3f} |")
        lines.append(f"| avg 0.5:0.9 | {m_o[4:].mean():.3f} | {m_g[4:].mean():.3f} | {m_g[4:].mean() - m_o[4:].mean():+.3f} |")
        okm = bool(np.all(np.abs(m_g - m_o) <= 0.1))
        ok5 = bool(abs(m_g[4:].mean() - m_o[4:].mean()) <= 0.1)
        if m in ("bf16x3", "fp32", "f16x3"):      # the modes that claim full parity
            ok &= okm
        lines += ["", f"Per-tIoU gate (every row within 0.1): {'PASSED' if okm else 'FAILED'}; north_star gate (mAP@[0.5:0.9] within "
                  f"0.1): {'PASSED' if ok5 else 'FAILED'}.", ""]
    out = os.path.join(ROOT, "profiles", os.environ.get("UNAV_MAP_PARITY_OUT", "r01_map_parity.md"))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return 0 if ok else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["dump", "eval"])
    ap.add_argument("--videos", type=int, default=64)
    ap.add_argument("--mode", default="bf16x3", help="comma-separated engine modes")
    a = ap.parse_args()
    sys.exit(dump(a) if a.stage == "dump" else evaluate(a))
