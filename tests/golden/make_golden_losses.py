"""Golden fixture for the loss-only tail of the eval forward: the REAL reference's
``results, losses = model(batch)`` (libs/modeling/multimodal_meta_archs.py:426-522) on a seeded batch of 3 videos with
event targets, the name-keyed synthetic weights, and the intermediate selections of
``Alignment.select_contrastive_embedding`` (multimodal_backbones.py:1080-1124).
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_losses.py       -> tests/golden/losses_b3.npz

The class of each video's first event (``key_labels``) is chosen from the reference's own per-frame class predictions so
that the non-key selection (frames outside the dilated event whose predicted class equals the event's class) is non-empty
for most videos; the labels are stored in the fixture and the batch is rebuilt from them by
``synth.add_event_targets``.  Two consecutive forwards are recorded: ``loss_normalizer`` is state that the reference also
updates in eval (:637-640).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402
from unav_yolyolva_b200 import synth  # noqa: E402

B, T, FIRST = 3, 224, 40


def main():
    torch.set_num_threads(8)
    import_reference()
    from libs.core import load_config
    from libs.modeling import make_multimodal_meta_arch

    cfg = load_config("/root/reference/configs/avel_unav100.yaml")
    torch.manual_seed(0)
    model = make_multimodal_meta_arch(cfg["model_name"], **cfg["model"])
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model.eval()

    # pass 1: per-frame class predictions of the Alignment heads on this batch
    batch = synth.add_event_targets(synth.make_batch(B, T, first_index=FIRST), FIRST)
    cap = {}
    al = model.alignment
    h1 = al.fc_video_cls.register_forward_hook(lambda m, i, o: cap.__setitem__("v", o.argmax(2)))
    h2 = al.fc_text_cls.register_forward_hook(lambda m, i, o: cap.__setitem__("t", o.argmax(2)))
    with torch.no_grad():
        model(batch)
    h1.remove(); h2.remove()
    key_labels = []
    for i in range(B):
        L = batch["lengths"][i]
        grown = torch.nn.functional.max_pool1d(batch["start_end"][i][None, None], 9, 1, 4)[0, 0] > 0
        outside = (~grown[:L - 1])
        votes = torch.bincount(torch.cat((cap["v"][i, :L - 1][outside], cap["t"][i, :L - 1][outside])), minlength=100)
        key_labels.append(int(votes.argmax()))
    print("key labels", key_labels)

    # pass 2+3: the recorded forwards
    model.loss_normalizer = cfg["model"]["train_cfg"]["init_loss_norm"]
    batch = synth.add_event_targets(synth.make_batch(B, T, first_index=FIRST), FIRST, key_labels=key_labels)
    picked = {"video": [], "text": []}
    orig = al.select_contrastive_embedding
    calls = []

    def spy(score, embedding, mask, label, cls_prd, cls_gt):
        k, n = orig(score, embedding, mask, label, cls_prd, cls_gt)
        calls.append(([int(x.shape[0]) for x in k], [int(x.shape[0]) for x in n]))
        return k, n

    al.select_contrastive_embedding = spy
    out = {"key_labels": np.array(key_labels), "first_index": np.array(FIRST),
           "init_loss_norm": np.array(float(model.loss_normalizer))}
    with torch.no_grad():
        for rep in range(2):
            results, losses = model(batch)
            for k, v in losses.items():
                out[f"{k}_{rep}"] = np.array(float(v), dtype=np.float64)
            out[f"loss_normalizer_{rep}"] = np.array(float(model.loss_normalizer))
    out["n_key_video"], out["n_nonkey_video"] = np.array(calls[0][0]), np.array(calls[0][1])
    out["n_key_text"], out["n_nonkey_text"] = np.array(calls[1][0]), np.array(calls[1][1])
    out["segments"] = results["segments"].numpy()
    out["scores"] = results["scores"].numpy()
    out["labels"] = results["labels"].numpy()
    np.savez_compressed(os.path.join(HERE, "losses_b3.npz"), **out)
    for k in sorted(out):
        if out[k].size <= 4:
            print(k, out[k])


if __name__ == "__main__":
    main()
