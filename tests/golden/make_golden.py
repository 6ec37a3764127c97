"""Generate the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference through ``oracle/ref_harness.py``, loads the name-keyed
trained-like weights of ``unav_yolyolva_b200/synth.py`` (strict ``load_state_dict`` — this also
pins the parameter naming contract), feeds seeded synthetic batches and stores what the reference
produced:

* ``state_dict_manifest.json``   names/shapes of ``PtTransformer.state_dict()`` (1235 entries)
* ``model_b2.npz``               B=2, T=224: logits / offsets / masks of all 6 levels, the final
                                  detections of ``model(batch)``, and strided samples of the
                                  Alignment and backbone outputs
* ``nms_cases.npz``              ``libs.utils.batched_nms`` (compiled ``nms_1d_cpu``) on random
                                  overlapping candidate sets, incl. ragged / single / empty classes
* ``decode_b2.npz``              ``inference_single_video`` candidates for the B=2 batch
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402


def nms_case(seed, n, ncls, skew=False):
    """Config-5 style candidates: centres clustered around 5 events per class, wide overlap."""
    g = torch.Generator().manual_seed(seed)
    if skew:
        labels = torch.where(torch.rand(n, generator=g) < 0.5, 0, torch.randint(0, ncls, (n,), generator=g))
    else:
        labels = torch.randint(0, ncls, (n,), generator=g)
    ev = torch.rand(ncls, 5, generator=g) * 200.0
    which = torch.randint(0, 5, (n,), generator=g)
    centre = ev[labels, which] + torch.randn(n, generator=g) * 2.0
    half = 0.5 + torch.rand(n, generator=g) * 19.5
    segs = torch.stack((centre - half, centre + half), dim=1).float()
    scores = (0.001 + torch.rand(n, generator=g) * 0.999).float()
    return segs, scores, labels.long()


def main():
    torch.set_num_threads(8)
    libs = import_reference()
    from libs.core import load_config
    from libs.modeling import make_multimodal_meta_arch
    from libs.utils import batched_nms

    from unav_yolyolva_b200 import synth

    cfg = load_config("/root/reference/configs/avel_unav100.yaml")
    torch.manual_seed(0)
    model = make_multimodal_meta_arch(cfg["model_name"], **cfg["model"])
    manifest = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)
    sd = synth.trained_like_state_dict(manifest)
    model.load_state_dict(sd, strict=True)
    model.eval()

    B, T = 2, 224
    batch = synth.make_batch(B, T, first_index=0)
    cap = {}

    def hook(name):
        def fn(mod, inp, out):
            cap[name] = out
        return fn

    model.alignment.register_forward_hook(hook("alignment"))
    model.backbone.register_forward_hook(hook("backbone"))
    model.cls_head.register_forward_hook(hook("cls"))
    model.reg_head.register_forward_hook(hook("reg"))
    orig_single = model.inference_single_video
    decoded = []

    def spy(points, fpn_masks, out_cls_logits, out_offsets):
        r = orig_single(points, fpn_masks, out_cls_logits, out_offsets)
        decoded.append({k: v.clone() for k, v in r.items()})
        return r

    model.inference_single_video = spy
    with torch.no_grad():
        results, losses = model(batch)

    out = {}
    for l in range(6):
        out[f"logits_{l}"] = cap["cls"][l].permute(0, 2, 1).contiguous().numpy()            # [B,T_l,100]
        r = cap["reg"][l].permute(0, 2, 1)
        out[f"offsets_{l}"] = r.reshape(B, r.shape[1], 100, 2).contiguous().numpy()         # [B,T_l,100,2]
        out[f"mask_{l}"] = cap["backbone"][2][l].squeeze(1).numpy()
        out[f"featV_{l}"] = cap["backbone"][0][l][:, ::7, :].contiguous().numpy()
        out[f"featA_{l}"] = cap["backbone"][1][l][:, ::7, :].contiguous().numpy()
    out["align_V"] = cap["alignment"][0][0][:, ::7, ::5].contiguous().numpy()
    out["align_A"] = cap["alignment"][1][0][:, ::7, ::5].contiguous().numpy()
    out["segments"] = results["segments"].numpy()
    out["scores"] = results["scores"].numpy()
    out["labels"] = results["labels"].numpy()
    out["lengths"] = np.array(batch["lengths"])
    np.savez_compressed(os.path.join(HERE, "model_b2.npz"), **out)

    dec = {}
    for i, d in enumerate(decoded):
        dec[f"segs_{i}"] = d["segments"].numpy()
        dec[f"scores_{i}"] = d["scores"].numpy()
        dec[f"labels_{i}"] = d["labels"].numpy()
    np.savez_compressed(os.path.join(HERE, "decode_b2.npz"), **dec)

    tc = cfg["model"]["test_cfg"]
    cases = {}
    specs = [(11, 3000, 100, False), (12, 2500, 100, True), (13, 441, 1, False),
             (14, 37, 100, False), (15, 1, 100, False), (16, 900, 7, False)]
    for ci, (seed, n, ncls, skew) in enumerate(specs):
        segs, scores, labels = nms_case(seed, n, ncls, skew)
        o_segs, o_scores, o_labels = batched_nms(
            segs.clone(), scores.clone(), labels.clone(), tc["iou_threshold"], 1e-4 if ci == 0 else tc["min_score"],
            tc["max_seg_num"], use_soft_nms=True, multiclass=True, sigma=tc["nms_sigma"],
            voting_thresh=tc["voting_thresh"])
        cases[f"in_segs_{ci}"] = segs.numpy()
        cases[f"in_scores_{ci}"] = scores.numpy()
        cases[f"in_labels_{ci}"] = labels.numpy()
        cases[f"min_score_{ci}"] = np.float32(1e-4 if ci == 0 else tc["min_score"])
        cases[f"out_segs_{ci}"] = o_segs.numpy()
        cases[f"out_scores_{ci}"] = o_scores.numpy()
        cases[f"out_labels_{ci}"] = o_labels.numpy()
    # hard-NMS variant (NMSop path, nms.py:8-35) on one case
    segs, scores, labels = nms_case(21, 1200, 20, False)
    o = batched_nms(segs.clone(), scores.clone(), labels.clone(), 0.5, tc["min_score"], tc["max_seg_num"],
                    use_soft_nms=False, multiclass=True, sigma=tc["nms_sigma"], voting_thresh=0.75)
    cases["hard_in_segs"], cases["hard_in_scores"], cases["hard_in_labels"] = segs.numpy(), scores.numpy(), labels.numpy()
    cases["hard_out_segs"], cases["hard_out_scores"], cases["hard_out_labels"] = [t.numpy() for t in o]
    np.savez_compressed(os.path.join(HERE, "nms_cases.npz"), **cases)
    print("golden fixtures written:", sorted(os.listdir(HERE)))
    print("results[0][:5]", results["segments"][0, :5], results["scores"][0, :5], results["labels"][0, :5])
    print("score range", float(results["scores"].min()), float(results["scores"].max()))


if __name__ == "__main__":
    main()
