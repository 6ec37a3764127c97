"""INTEGRATION.md section 1 checked in the build container: a checkout of the reference with the TWO documented shim files
(`libs/modeling/__init__.py`, `libs/utils/nms.py`) imports, builds the model through the reference's own config loader and
registry call, and loads a DataParallel-prefixed state_dict strictly — i.e. everything `eval.py` does before the first
forward (eval.py:15-18, 41-72).  The overlay is built from symlinks in a temporary directory at test time; no reference
source enters the repository.  Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle.ref_harness import REF_ROOT, STUBS, have_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM_MODELING = """\
from unav_yolyolva_b200.modeling import (MaskedConv1D, MaskedMHCA, LayerNorm, TransformerBlock, Scale, AffineDropPath,
                                         make_multimodal_backbone, make_multimodal_meta_arch, make_dependency_block)
"""
SHIM_NMS = "from unav_yolyolva_b200.utils.nms import batched_nms\n"


def _overlay(tmp):
    libs = os.path.join(tmp, "libs")
    os.makedirs(libs)
    src = os.path.join(REF_ROOT, "libs")
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
    open(os.path.join(libs, "modeling", "__init__.py"), "w").write(SHIM_MODELING)
    open(os.path.join(libs, "utils", "nms.py"), "w").write(SHIM_NMS)


@pytest.mark.skipif(not have_reference(), reason="reference tree not on this machine")
def test_reference_checkout_with_the_two_shims(tmp_path):
    _overlay(str(tmp_path))
    code = textwrap.dedent("""
        import torch, torch.nn as nn
        from libs.core import load_config
        from libs.modeling import make_multimodal_meta_arch
        from libs.utils import valid_one_epoch, ANETdetection, fix_random_seed, batched_nms
        import libs.utils.train_utils as tu
        import unav_yolyolva_b200.modeling as ours
        from unav_yolyolva_b200 import synth
        assert tu.MaskedConv1D is ours.MaskedConv1D and tu.LayerNorm is ours.LayerNorm       # train_utils.py:17 resolves to the shim
        assert batched_nms.__module__ == "unav_yolyolva_b200.utils.nms"
        cfg = load_config("%s/configs/avel_unav100.yaml")
        model = make_multimodal_meta_arch(cfg["model_name"], **cfg["model"])                  # eval.py:60
        assert type(model).__module__ == "unav_yolyolva_b200.modeling.multimodal_meta_archs"
        model = nn.DataParallel(model, device_ids=[])                                         # eval.py:61 (no GPU here)
        sd = synth.trained_like_state_dict(prefix="module.")
        r = model.load_state_dict(sd, strict=True)                                            # eval.py:72
        assert not r.missing_keys and not r.unexpected_keys
        print("SHIM_OK", len(sd))
    """ % REF_ROOT)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), STUBS, ROOT]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "SHIM_OK 1235" in out.stdout, out.stderr[-2000:]


@pytest.mark.skipif(not have_reference(), reason="reference tree not on this machine")
def test_eval_harness_reference_cpu_run_and_shipped_copy(tmp_path):
    """The eval.py harness of tests/test_gpu_eval_dropin.py, CPU half: synthetic dataset + annotations + checkpoint + config ->
    the reference's own evaluation loop on the CPU prints a non-trivial mAP table; and the copy that ships to the GPU box
    (oracle/_ref/reference) is byte-identical to the reference's Python tree."""
    import filecmp

    import numpy as np
    import torch

    from oracle import eval_dropin as ED
    from oracle import ref_harness, ref_runner
    from unav_yolyolva_b200 import synth
    ship = ref_harness.ship_reference()
    for rel in ("eval.py", "libs/modeling/multimodal_meta_archs.py", "libs/utils/nms.py", "libs/utils/train_utils.py",
                "libs/datasets/data_utils.py", "configs/avel_unav100.yaml"):
        assert filecmp.cmp(os.path.join(ship, rel), os.path.join(REF_ROOT, rel), shallow=False), rel
    tmp = str(tmp_path)
    n = 4
    items = ED.write_features(os.path.join(tmp, "feats"), n, first_index=300)
    sd = synth.trained_like_state_dict()
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    res, _ = ref_runner.reference_forward(ref_runner.build_reference_model(sd), synth.make_batch(n, 224, first_index=300))
    ED.write_annotations(os.path.join(tmp, "anno.json"), items, res["segments"].numpy(), res["labels"].numpy())
    ED.write_checkpoint(os.path.join(tmp, "ckpt", "model.pth.tar"), sd)
    ED.write_config(os.path.join(tmp, "cfg.yaml"), os.path.join(tmp, "anno.json"), os.path.join(tmp, "feats"), os.path.join(tmp, "out"),
                    batch_size=4, workers=1)
    tab = ED.run_reference_cpu(os.path.join(tmp, "cfg.yaml"), os.path.join(tmp, "ckpt", "model.pth.tar"))
    assert len(tab) == 10 and tab["avg"] > 10.0, tab
    overlay = ED.make_overlay(os.path.join(tmp, "checkout"))
    assert os.path.islink(os.path.join(overlay, "eval.py"))
    assert open(os.path.join(overlay, "libs", "utils", "nms.py")).read() == ED.SHIM_NMS
