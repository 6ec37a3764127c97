"""The reference's UNMODIFIED ``eval.py`` (eval.py:22-103 -> libs/utils/train_utils.py:378-466, ``nn.DataParallel`` at
eval.py:61) driving the B200 hot path through the two documented shim files, on a GPU — and the mAP
it prints against the reference's own model evaluated on the CPU over the SAME feature files, annotation file, config and
checkpoint (gate: <= 0.1 point at every tIoU and on the average).

The reference tree reaches the GPU box as ``oracle/_ref/reference`` (git-ignored copy made by ``oracle/ref_harness.ship_reference``
from ``__graft_entry__.build()``); nothing here reads ``/root/reference`` at run time.  Everything under ``oracle/`` is the
checker: the product run is the ``eval.py`` subprocess, whose ``libs.modeling`` / ``libs.utils.nms`` resolve to this package."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_dropin as ED
from oracle import ref_harness, ref_runner
from unav_yolyolva_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_runner.reference_available(), reason="reference tree not shipped (oracle/_ref/reference missing)")
def test_unmodified_eval_py_on_gpu_matches_reference_cpu_map(cuda, tmp_path):
    n, first = 64, 200
    tmp = str(tmp_path)
    feat_dir, anno, ckpt, cfg = (os.path.join(tmp, x) for x in ("feats", "anno.json", "ckpt/model.pth.tar", "cfg.yaml"))
    items = ED.write_features(feat_dir, n, first_index=first)
    sd = synth.trained_like_state_dict()
    # ground truth from the reference's own detections on these videos (in-process CPU forward of the unmodified model)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    ref_model = ref_runner.build_reference_model(sd)
    segs, labels = [], []
    for b0 in range(0, n, 8):
        res, _ = ref_runner.reference_forward(ref_model, synth.make_batch(8, 224, first_index=first + b0))
        segs.append(res["segments"].numpy()); labels.append(res["labels"].numpy())
    del ref_model
    ED.write_annotations(anno, items, np.concatenate(segs), np.concatenate(labels))
    ED.write_checkpoint(ckpt, sd)
    ED.write_config(cfg, anno, feat_dir, os.path.join(tmp, "out"), batch_size=8, workers=2)

    overlay = ED.make_overlay(os.path.join(tmp, "checkout"))
    assert os.path.realpath(os.path.join(overlay, "eval.py")) == os.path.realpath(os.path.join(ref_harness.reference_root(), "eval.py"))
    ours = ED.run_eval_py(overlay, cfg, ckpt)                 # GPU: eval.py + DataParallel(['cuda:0']) + valid_one_epoch, our model
    ref = ED.run_reference_cpu(cfg, ckpt)                      # CPU FP32: the reference's model, same files
    print("tIoU   ours(GPU)  reference(CPU)")
    for k in sorted(ref):
        print(f"{k:>5}  {ours[k]:8.2f}  {ref[k]:8.2f}")
    assert set(ours) == set(ref) and len(ref) == 10
    assert ref["avg"] > 10.0, "the synthetic ground truth should give a non-trivial mAP"
    for k in ref:
        assert abs(ours[k] - ref[k]) <= 0.1 + 1e-6, f"mAP at tIoU {k}: ours {ours[k]} vs reference {ref[k]}"
