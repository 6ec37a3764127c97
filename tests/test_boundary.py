"""Drop-in boundary: registry names, constructor kwargs, state_dict contract, loud failure without CUDA."""
import os

import pytest
import torch

from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.config import default_model_cfg
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch, models


@pytest.fixture(scope="module")
def model():
    return make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())


def test_registry_names():
    assert "LocPointTransformer" in models.multimodal_meta_archs
    assert "convTransformer" in models.multimodal_backbones
    assert "DependencyBlock" in models.dependency_blocks


def test_state_dict_matches_reference_manifest(model):
    man = synth.load_manifest()
    sd = model.state_dict()
    assert len(man) == 1235
    assert set(sd) == set(man)
    for k, shape in man.items():
        assert list(sd[k].shape) == shape, k
    assert sum(p.numel() for p in model.parameters()) == 97373087
    # aliased modules share storage
    assert sd["alignment.multiway_list.0.norm1_fused.weight"].data_ptr() == sd["alignment.multiway_list.1.norm1_fused.weight"].data_ptr()
    assert len({v.data_ptr() for v in sd.values()}) == 1197


def test_strict_load_with_dataparallel_prefix(model):
    sd = synth.trained_like_state_dict(prefix="module.")
    dp = torch.nn.DataParallel(model, device_ids=[]) if not torch.cuda.is_available() else None
    if dp is None:
        pytest.skip("DataParallel pass-through check is for the CPU box")
    assert dp.load_state_dict(sd, strict=True).missing_keys == []


def test_non_persistent_pos_embd(model):
    assert "backbone.pos_embd" not in model.state_dict()
    assert tuple(model.backbone.pos_embd.shape) == (1, 512, 224)


def test_forward_fails_loudly_without_cuda(model):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    model.eval()
    batch = synth.make_batch(1)
    with pytest.raises(RuntimeError):
        model(batch)


def test_config_matches_reference_loader():
    from oracle.ref_harness import have_reference, import_reference
    if not have_reference():
        pytest.skip("reference tree not on this machine")
    import_reference()
    from libs.core import load_config
    import copy
    from libs.core import config as ref_config
    cfg = load_config("/root/reference/configs/avel_unav100.yaml", defaults=copy.deepcopy(ref_config.DEFAULTS))["model"]
    mine = default_model_cfg()
    assert set(cfg) == set(mine)
    for k in cfg:
        assert cfg[k] == mine[k] or list(cfg[k]) == list(mine[k]), k


def test_engine_rejects_structures_outside_the_fused_plan():
    """HotPathEngine hard-codes the avel_unav100 structure; any other valid reference config must raise, not mis-pack."""
    import copy

    import pytest

    from unav_yolyolva_b200.config import default_model_cfg
    from unav_yolyolva_b200.engine import HotPathEngine
    from unav_yolyolva_b200.modeling import make_multimodal_meta_arch
    from unav_yolyolva_b200.utils.nms import nms_method_code
    HotPathEngine._validate_structure(make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg()))
    for key, val in (("head_num_layers", 2), ("use_abs_pe", False), ("class_aware", False), ("head_with_ln", False),
                     ("backbone_arch", (3, 3, 5))):
        cfg = copy.deepcopy(default_model_cfg())
        cfg[key] = val
        with pytest.raises(NotImplementedError):
            HotPathEngine._validate_structure(make_multimodal_meta_arch("LocPointTransformer", **cfg))
    assert nms_method_code("soft", True) == 2 and nms_method_code("hard", True) == 3
    for bad in (("none", True), ("soft", False)):
        with pytest.raises(NotImplementedError):
            nms_method_code(*bad)
