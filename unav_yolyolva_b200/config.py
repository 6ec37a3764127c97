"""Model / test configuration of the reference's ``configs/avel_unav100.yaml`` merged over its DEFAULTS.

This restates the *values* the reference produces with ``load_config`` (/root/reference/libs/core/config.py:4-155
over /root/reference/configs/avel_unav100.yaml:1-43) so that tests and ``bench.py`` can build the model on a
machine without the reference tree.  Under the reference's own ``eval.py`` the reference's ``libs.core`` is
used unchanged; ``tests/test_boundary.py`` checks (in the build container) that both give the same kwargs.
"""
from __future__ import annotations

import copy

TRAIN_CFG = {
    "loss_weight": 1, "evaluate": True, "eval_freq": 2, "cls_prior_prob": 0.01, "init_loss_norm": 250,
    "clip_grad_l2norm": 1.0, "head_empty_cls": [], "dropout": 0.0, "droppath": 0.1, "label_smoothing": 0.0,
}
TEST_CFG = {
    "pre_nms_topk": 2000, "max_seg_num": 100, "min_score": 0.001, "multiclass_nms": True, "nms_sigma": 0.4,
    "iou_threshold": 0.7, "pre_nms_thresh": 0.001, "nms_method": "soft", "duration_thresh": 0.05,
    "ext_score_file": None, "voting_thresh": 0.75,
}
MODEL_CFG = {
    "input_dim_V": 512, "input_dim_A": 512, "use_abs_pe": True, "class_aware": True, "use_dependency": False,
    "intra_contr_weight": 1.0, "inter_contr_weight": 0.001, "score_V_weight": 0.001, "score_A_weight": 0.001,
    "backbone_type": "convTransformer", "dependency_type": "DependencyBlock", "backbone_arch": (2, 3, 5),
    "scale_factor": 2,
    "regression_range": [(0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000)],
    "n_head": 4, "embd_kernel_size": 3, "embd_dim": 512, "embd_with_ln": True, "head_dim": 512,
    "head_kernel_size": 3, "head_num_layers": 3, "head_with_ln": True, "num_classes": 100, "max_seq_len": 224,
}
DATASET_CFG = {"feat_stride": 8, "num_frames": 24, "default_fps": 25, "num_classes": 100, "max_seq_len": 224}


def default_model_cfg(max_seq_len: int = 224, **test_overrides) -> dict:
    """kwargs for ``make_multimodal_meta_arch('LocPointTransformer', **cfg)``."""
    cfg = copy.deepcopy(MODEL_CFG)
    cfg["max_seq_len"] = max_seq_len
    cfg["train_cfg"] = copy.deepcopy(TRAIN_CFG)
    cfg["test_cfg"] = copy.deepcopy(TEST_CFG)
    cfg["test_cfg"].update(test_overrides)
    return cfg
