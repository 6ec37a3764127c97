"""CudaPrefetcher (unav_yolyolva_b200/ingest.py): the overlapped upload must hand the model exactly the bytes a
synchronous ``.to(device)`` would, in order, for ragged last batches too, and the detections must not change."""
import pytest
import torch

from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.config import default_model_cfg
from unav_yolyolva_b200.ingest import CudaPrefetcher
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch

pytestmark = pytest.mark.gpu


def _batches(sizes):
    first = 0
    for n in sizes:
        yield synth.make_batch(n, 224, first_index=first, with_gt=False)
        first += n


def test_prefetcher_preserves_order_and_bytes():
    dev = torch.device("cuda", 0)
    sizes = [3, 3, 3, 2, 3, 1]
    ref = list(_batches(sizes))
    seen = 0
    for got, want in zip(CudaPrefetcher(_batches(sizes), dev), ref):
        for k in ("visual", "audio", "mask"):
            assert got[k].is_cuda
            assert torch.equal(got[k].cpu(), want[k]), k
        assert got["video_id"] == want["video_id"]
        seen += 1
    assert seen == len(sizes)


def test_prefetcher_same_detections_as_sync_upload():
    dev = torch.device("cuda", 0)
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model = model.to(dev).eval()
    sizes = [4, 4, 4, 4, 4]
    want = []
    for b in _batches(sizes):
        res, _ = model(b)
        want.append({k: res[k].cpu() for k in ("segments", "scores", "labels")})
    # a slow consumer and a fast one: sleep on the stream before some steps so uploads run ahead / behind
    for j, b in enumerate(CudaPrefetcher(_batches(sizes), dev)):
        if j % 2:
            torch.cuda._sleep(20_000_000)
        res, _ = model(b)
        for k in ("segments", "scores", "labels"):
            assert torch.equal(res[k].cpu(), want[j][k]), (j, k)


def test_prefetcher_rejects_cpu_device():
    with pytest.raises(RuntimeError):
        CudaPrefetcher(iter([]), "cpu")


def test_submit_result_matches_forward_and_pipelines():
    dev = torch.device("cuda", 0)
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model = model.to(dev).eval()
    sizes = [4, 4, 2, 4, 4, 4]
    want = []
    for b in _batches(sizes):
        res, _ = model(b)
        want.append({k: res[k].cpu() for k in ("segments", "scores", "labels")})
    got, prev = [], None
    for b in CudaPrefetcher(_batches(sizes), dev):
        cur = model.submit(b)
        if prev is not None:
            got.append({k: v.clone() for k, v in prev.result().items()})
        prev = cur
    got.append({k: v.clone() for k, v in prev.result().items()})
    assert len(got) == len(want)
    for j, (g, w) in enumerate(zip(got, want)):
        for k in ("segments", "scores", "labels"):
            assert g[k].device.type == "cpu" and torch.equal(g[k], w[k]), (j, k)
