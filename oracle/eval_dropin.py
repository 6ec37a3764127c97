"""Harness for running the reference's UNMODIFIED ``eval.py`` — TEST INFRASTRUCTURE ONLY (never imported by the product path).

SURVEY.md section 4 level 4 / VERDICT r01 item 8: `north_star` says the hot path "drops into `eval.py` unchanged".  This module
builds everything such a run needs out of synthetic data, in a scratch directory:

* a feature folder in the reference dataset's on-disk format (``<id>_rgb.npy`` / ``<id>_flow.npy`` [L, 1024],
  ``<id>_vggish.npy`` [L, 128]; libs/datasets/unav100.py:227-248) holding the same seeded videos as ``synth.make_items``;
* an annotation JSON in the format ``_load_json_db`` / ``load_gt_seg_from_json`` read (unav100.py:118-170,
  libs/utils/metrics.py:33-71) whose ground truth is derived from the reference's own detections (ranks 1/4/9 per video,
  5 % boundary jitter) so that mAP is non-trivial and sensitive to ranking / boundary changes;
* a ``state_dict_ema`` checkpoint with DataParallel's ``module.`` prefix (eval.py:64-72);
* a YAML config = the reference's ``configs/avel_unav100.yaml`` values with the paths, ``devices: [0]`` and a small
  loader (the file is an INPUT of eval.py, not code);
* the drop-in overlay of INTEGRATION.md section 1: a directory of symlinks to the reference checkout in which exactly two
  files are replaced by shims (``libs/modeling/__init__.py``, ``libs/utils/nms.py``); ``eval.py`` itself is a symlink to the
  reference's file, byte for byte.

``run_eval_py`` then executes ``python eval.py --config ... --ckpt ...`` in a subprocess and parses the mAP table that the
reference's ``ANETdetection.evaluate`` prints.  ``run_reference_cpu`` executes the same evaluation with the reference's own
model on the CPU (``oracle/ref_eval_cpu.py``) for the ground-truth numbers.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys
from typing import Dict, List

import numpy as np
import torch

from . import ref_harness

ROOT = os.path.dirname(ref_harness.ORACLE_DIR)

SHIM_MODELING = """\
from unav_yolyolva_b200.modeling import (MaskedConv1D, MaskedMHCA, LayerNorm, TransformerBlock, Scale, AffineDropPath,
                                         make_multimodal_backbone, make_multimodal_meta_arch, make_dependency_block)
"""
SHIM_NMS = "from unav_yolyolva_b200.utils.nms import batched_nms\n"


def make_overlay(dst: str) -> str:
    """Reference checkout with the two shim files, built from symlinks; returns ``dst``."""
    ref = ref_harness.reference_root()
    assert ref is not None, "no reference tree on this machine"
    libs = os.path.join(dst, "libs")
    os.makedirs(libs)
    src = os.path.join(ref, "libs")
    for name in os.listdir(src):
        s = os.path.join(src, name)
        if name in ("modeling", "utils"):
            os.makedirs(os.path.join(libs, name))
            for f in os.listdir(s):
                if (name, f) in (("modeling", "__init__.py"), ("utils", "nms.py")) or f == "__pycache__":
                    continue
                os.symlink(os.path.join(s, f), os.path.join(libs, name, f))
        elif name != "__pycache__":
            os.symlink(s, os.path.join(libs, name))
    with open(os.path.join(libs, "modeling", "__init__.py"), "w") as f:
        f.write(SHIM_MODELING)
    with open(os.path.join(libs, "utils", "nms.py"), "w") as f:
        f.write(SHIM_NMS)
    os.symlink(os.path.join(ref, "eval.py"), os.path.join(dst, "eval.py"))          # the reference's file itself
    return dst


def write_features(feat_dir: str, n_videos: int, first_index: int = 0) -> List[dict]:
    """The seeded synthetic videos of ``synth.make_items`` as .npy files in the reference dataset's layout."""
    from unav_yolyolva_b200 import synth
    os.makedirs(feat_dir, exist_ok=True)
    items = synth.make_items(n_videos, first_index=first_index)
    for it in items:
        vis = it["feats"]["visual"].numpy().T                      # [L, 2048] = hstack(rgb, flow) (unav100.py:234)
        aud = it["feats"]["audio"].numpy().T                       # [L, 128]
        vid = it["video_id"]
        np.save(os.path.join(feat_dir, vid + "_rgb.npy"), np.ascontiguousarray(vis[:, :1024]))
        np.save(os.path.join(feat_dir, vid + "_flow.npy"), np.ascontiguousarray(vis[:, 1024:]))
        np.save(os.path.join(feat_dir, vid + "_vggish.npy"), np.ascontiguousarray(aud))
    return items


def write_annotations(json_file: str, items: List[dict], det_segments: np.ndarray, det_labels: np.ndarray, seed: int = 7) -> None:
    """Ground truth = the given detections ranked 1 / 4 / 9 per video with 5 % boundary jitter, subset 'test'."""
    rng = np.random.default_rng(seed)
    db = {}
    for v, it in enumerate(items):
        ants = []
        dur = float(it["duration"])
        for rank in (0, 3, 8):
            s, e = (float(x) for x in det_segments[v, rank])
            w = max(e - s, 0.2)
            s2, e2 = s + rng.normal(0, 0.05) * w, e + rng.normal(0, 0.05) * w
            # inside [0.6 s, duration - 0.7 s]: collate_fcn indexes per-step tensors with seg / 1.28 (data_utils.py:150-157),
            # which must stay within [0, L)
            lo, hi = max(0.6, min(s2, e2)), min(max(s2, e2), dur - 0.7)
            if hi - lo < 0.1:
                lo, hi = max(0.6, hi - 0.5), max(hi, 1.1)
            lab = int(det_labels[v, rank])
            ants.append({"segment": [lo, hi], "label_id": lab, "label": str(lab)})
        db[it["video_id"]] = {"subset": "test", "duration": dur, "annotations": ants}
    with open(json_file, "w") as f:
        json.dump({"database": db}, f)


def write_checkpoint(path: str, state_dict: Dict[str, torch.Tensor]) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({"state_dict_ema": {"module." + k: v for k, v in state_dict.items()}}, path)


def write_config(path: str, json_file: str, feat_dir: str, out_dir: str, batch_size: int = 8, workers: int = 2) -> None:
    """configs/avel_unav100.yaml with this run's paths and ``devices: [0]``: the default ['cuda:1'] cannot work under the
    CUDA_VISIBLE_DEVICES=0 pin of eval.py:19-20, and with torch >= 2 eval.py:66-69's
    ``storage.cuda(cfg['devices'][0])`` only accepts an integer index (``torch.device('cuda', 'cuda:0')`` raises TypeError)."""
    cfg = f"""dataset_name: unav100
devices: [0]
dataset: {{
  json_file: {json_file},
  feat_folder: {feat_dir},
  file_prefix: ~,
  file_ext: .npy,
  max_seq_len: 224,
}}
model: {{
  input_dim_V: 512,
  input_dim_A: 512,
  use_abs_pe: True,
  class_aware: True,
  use_dependency: False,
  intra_contr_weight: 1.0,
  inter_contr_weight: 0.001,
  score_V_weight: 0.001,
  score_A_weight: 0.001,
}}
opt: {{
  learning_rate: 0.0001,
  epochs: 40,
  weight_decay: 0.0001,
  warmup_epochs: 5,
}}
loader: {{
  batch_size: {batch_size},
  num_workers: {workers},
}}
train_cfg: {{
  loss_weight: 1,
  evaluate: True,
  eval_freq: 2,
}}
test_cfg: {{
  pre_nms_topk: 2000,
  max_seg_num: 100,
  min_score: 0.001,
  multiclass_nms: True,
  nms_sigma : 0.4,
  iou_threshold: 0.7,
}}
output_folder: {out_dir}
"""
    with open(path, "w") as f:
        f.write(cfg)


_ROW = re.compile(r"\|tIoU = ([0-9.]+): mAP = ([0-9.]+) \(%\)")
_AVG = re.compile(r"Avearge mAP: ([0-9.]+) \(%\)")


def parse_map_table(stdout: str) -> Dict[str, float]:
    """The table ``ANETdetection.evaluate(verbose=True)`` prints (libs/utils/metrics.py:283-293), in percent."""
    rows = {m.group(1): float(m.group(2)) for m in _ROW.finditer(stdout)}
    avg = _AVG.search(stdout)
    if not rows or avg is None:
        raise RuntimeError("no mAP table in the output:\n" + stdout[-3000:])
    rows["avg"] = float(avg.group(1))
    return rows


def _env(extra_path: List[str], cuda_visible=None) -> dict:
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(extra_path + [ref_harness.STUBS, ref_harness.REF_OUT, ROOT])
    if cuda_visible is not None:
        env["CUDA_VISIBLE_DEVICES"] = cuda_visible
    return env


def run_eval_py(workdir: str, config: str, ckpt: str, timeout: int = 900) -> Dict[str, float]:
    """``python eval.py --config ... --ckpt ...`` in ``workdir`` (an overlay from ``make_overlay`` or the reference root)."""
    out = subprocess.run([sys.executable, "eval.py", "--config", config, "--ckpt", ckpt, "--print-freq", "1000000"], cwd=workdir,
                         env=_env([workdir]), capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"eval.py failed ({out.returncode}):\n{out.stdout[-2000:]}\n{out.stderr[-4000:]}")
    return parse_map_table(out.stdout)


def run_reference_cpu(config: str, ckpt: str, timeout: int = 1800) -> Dict[str, float]:
    """The same evaluation with the reference's own model on the CPU (FP32 ground truth, SURVEY.md section 8c caveat 1)."""
    out = subprocess.run([sys.executable, os.path.join(ref_harness.ORACLE_DIR, "ref_eval_cpu.py"), "--config", config, "--ckpt", ckpt],
                         cwd=ROOT, env=_env([ref_harness.reference_root()], cuda_visible=""), capture_output=True, text=True,
                         timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"ref_eval_cpu.py failed ({out.returncode}):\n{out.stdout[-2000:]}\n{out.stderr[-4000:]}")
    return parse_map_table(out.stdout)
