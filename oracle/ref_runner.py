"""Drive the UNMODIFIED reference model — TEST INFRASTRUCTURE ONLY (never imported by the product path).

``bench.py --impl reference`` / ``cpu_baseline`` and the parity tests use these helpers to run the reference's own
``PtTransformer`` (``libs/modeling/multimodal_meta_archs.py:263-875``) + its compiled ``nms_1d_cpu`` through its own public API:
``make_multimodal_meta_arch(cfg['model_name'], **cfg['model'])`` (eval.py:60) and ``model(video_list)`` under
``torch.no_grad()`` (libs/utils/train_utils.py:409-412).  The reference tree is ``/root/reference`` in the build container
and ``oracle/_ref/reference`` (shipped by ``ref_harness.ship_reference``) on the GPU box.
"""
from __future__ import annotations

import os

import torch

from . import ref_harness


def reference_available() -> bool:
    return ref_harness.reference_root() is not None and ref_harness._find_ref_nms_so() is not None


def reference_config(max_seq_len: int = 224) -> dict:
    """The reference's own merged config (libs/core/config.py over configs/avel_unav100.yaml)."""
    ref_harness.import_reference()
    from libs.core import load_config
    cfg = load_config(os.path.join(ref_harness.reference_root(), "configs", "avel_unav100.yaml"))
    assert cfg["model"]["max_seq_len"] == max_seq_len
    return cfg


def build_reference_model(state_dict, device="cpu"):
    """The reference's PtTransformer with ``state_dict`` loaded strictly, in eval mode (eval.py:60-72 without DataParallel)."""
    ref_harness.import_reference()
    from libs.modeling import make_multimodal_meta_arch
    cfg = reference_config()
    model = make_multimodal_meta_arch(cfg["model_name"], **cfg["model"])
    model.load_state_dict(state_dict, strict=True)
    return model.to(device).eval()


@torch.no_grad()
def reference_forward(model, batch):
    """``model(video_list)`` exactly as ``valid_one_epoch`` calls it; returns (results, losses).  ``batch`` is a collate dict
    with the GT tensors (the reference's eval forward evaluates the losses unconditionally, multimodal_meta_archs.py:504-509)."""
    return model(batch)
