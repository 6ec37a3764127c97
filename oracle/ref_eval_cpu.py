"""The reference's evaluation on the CPU — TEST INFRASTRUCTURE ONLY.

`eval.py` itself cannot run without a GPU (it maps the checkpoint with ``storage.cuda(...)``, eval.py:66-69), so the FP32 CPU
ground truth for the drop-in test is produced by this driver, which performs eval.py's steps 0-5 (eval.py:22-103) with the
reference's own functions and classes — ``load_config``, ``make_dataset`` / ``make_data_loader``, ``make_multimodal_meta_arch``,
``nn.DataParallel``, ``ANETdetection``, ``valid_one_epoch`` — and only differs in ``map_location='cpu'`` and an empty
``device_ids`` list.  Run with the reference tree and oracle/stubs on PYTHONPATH and CUDA_VISIBLE_DEVICES="" (see
``oracle/eval_dropin.run_reference_cpu``).
"""
import argparse
import os

import torch
import torch.nn as nn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--ckpt", required=True)
    args = ap.parse_args()
    from libs.core import load_config
    from libs.datasets import make_dataset, make_data_loader
    from libs.modeling import make_multimodal_meta_arch
    from libs.utils import ANETdetection, fix_random_seed, valid_one_epoch
    assert not torch.cuda.is_available(), "run with CUDA_VISIBLE_DEVICES=\"\" (CPU FP32 ground truth)"
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = load_config(args.config)
    fix_random_seed(0, include_cuda=False)
    val_dataset = make_dataset(cfg["dataset_name"], False, cfg["test_split"], **cfg["dataset"])
    val_loader = make_data_loader(val_dataset, False, None, **cfg["loader"], **cfg["dataset"])
    model = make_multimodal_meta_arch(cfg["model_name"], **cfg["model"])
    model = nn.DataParallel(model, device_ids=[])
    checkpoint = torch.load(args.ckpt, map_location="cpu")
    model.load_state_dict(checkpoint["state_dict_ema"])
    det_eval = ANETdetection(val_dataset.json_file, val_dataset.split[0],
                             tiou_thresholds=val_dataset.get_attributes()["tiou_thresholds"])
    valid_one_epoch(val_loader, model, -1, evaluator=det_eval, output_file=None,
                    ext_score_file=cfg["test_cfg"]["ext_score_file"], tb_writer=None, print_freq=10 ** 9)


if __name__ == "__main__":
    main()
