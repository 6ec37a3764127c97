"""CudaPrefetcher (unav_yolyolva_b200/ingest.py): the overlapped upload must hand the model exactly the bytes a
synchronous ``.to(device)`` would, in order, for ragged last batches too, and the detections must not change."""
import os

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


# ------------------------------------------------------------------------------------------------------------------
# device-side collate
def test_device_collate_matches_reference_golden(golden_dir):
    """unav_collate_pad through DeviceCollator vs what the reference's collate_fcn produced (bit-exact)."""
    import os
    import numpy as np
    from unav_yolyolva_b200.ingest import DeviceCollator
    d = np.load(os.path.join(golden_dir, "collate_b3.npz"))
    n = len(d["lens"])
    items = [{"video_id": f"v{i}", "feats": {"visual": torch.from_numpy(d[f"visual_{i}"]), "audio": torch.from_numpy(d[f"audio_{i}"])},
              "fps": 25.0, "duration": 1.0, "feat_stride": 8, "feat_num_frames": 24} for i in range(n)]
    out = DeviceCollator(int(d["max_seq_len"]), "cuda:0")(items)
    torch.cuda.synchronize()
    assert np.array_equal(out["visual"].cpu().numpy(), d["visual"])
    assert np.array_equal(out["audio"].cpu().numpy(), d["audio"])
    assert out["mask"].dtype == torch.bool and np.array_equal(out["mask"].cpu().numpy(), d["mask"])
    assert out["video_id"] == ["v0", "v1", "v2"]


def test_device_collate_full_size_and_model_parity():
    """Ragged items of the full-size synthetic videos: padded batch identical to the host collate (synth.make_batch),
    through the prefetcher with ragged batch sizes, and identical detections from the model."""
    from oracle import collate_ref
    from unav_yolyolva_b200.ingest import DeviceCollator
    dev = torch.device("cuda", 0)
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model = model.to(dev).eval()
    sizes, first = [4, 4, 3, 4], 0
    lists, batches = [], []
    for n in sizes:
        lists.append(synth.make_items(n, first))
        batches.append(synth.make_batch(n, 224, first_index=first, with_gt=False))
        first += n
    want = []
    for b in batches:
        res, _ = model(b)
        want.append({k: res[k].cpu() for k in ("segments", "scores", "labels")})
    coll = DeviceCollator(224, dev)
    for j, got in enumerate(CudaPrefetcher(iter(lists), dev, collate=coll)):
        ref_v, ref_m = collate_ref.collate_pad([x["feats"]["visual"].numpy() for x in lists[j]], 224)
        assert torch.equal(got["visual"].cpu(), torch.from_numpy(ref_v)) and torch.equal(got["visual"].cpu(), batches[j]["visual"])
        assert torch.equal(got["audio"].cpu(), batches[j]["audio"])
        assert torch.equal(got["mask"].cpu(), batches[j]["mask"]) and torch.equal(got["mask"].cpu(), torch.from_numpy(ref_m))
        assert got["duration"] == batches[j]["duration"] and got["video_id"] == batches[j]["video_id"]
        res, _ = model(got)
        for k in ("segments", "scores", "labels"):
            assert torch.equal(res[k].cpu(), want[j][k]), (j, k)


def test_device_collate_pinned_items_upload_direct_and_pack_thread():
    """Items whose features already live in pinned host memory are uploaded from where they lie (no host repack); pageable
    items are packed into the staging slot by the prefetcher's background thread.  Both give the host collate's bytes, in
    order, for more batches than there are slots."""
    from unav_yolyolva_b200.ingest import DeviceCollator
    dev = torch.device("cuda", 0)
    sizes = [3, 4, 2, 4, 4, 1, 3]
    for pinned in (False, True):
        lists, first = [], 40
        for n in sizes:
            items = synth.make_items(n, first)
            if pinned:
                for it in items:
                    it["feats"] = {k: v.pin_memory() for k, v in it["feats"].items()}
            lists.append(items)
            first += n
        for threaded in (True, False):
            coll = DeviceCollator(224, dev, direct_pinned=pinned)
            first = 40
            for j, got in enumerate(CudaPrefetcher(iter(lists), dev, collate=coll, depth=3, pack_thread=threaded)):
                ref = synth.make_batch(sizes[j], 224, first_index=first, with_gt=False)
                first += sizes[j]
                assert torch.equal(got["visual"].cpu(), ref["visual"]) and torch.equal(got["audio"].cpu(), ref["audio"])
                assert torch.equal(got["mask"].cpu(), ref["mask"]) and got["video_id"] == ref["video_id"]
            assert j == len(sizes) - 1


def test_device_collate_rejects_cpu_and_mismatched_lengths():
    from unav_yolyolva_b200.ingest import DeviceCollator
    with pytest.raises(RuntimeError):
        DeviceCollator(224, "cpu")
    it = synth.make_items(2, 0)
    it[1]["feats"]["audio"] = it[1]["feats"]["audio"][:, :-1]
    with pytest.raises(ValueError):
        DeviceCollator(224, "cuda:0")(it)


def test_valid_one_epoch_dropin(tmp_path):
    """runner.valid_one_epoch (the reference's evaluation loop, overlapped) over un-collated items with the device-side
    collate and the device mAP evaluator: same detections and the same mAP as the synchronous loop."""
    import json
    import numpy as np
    from unav_yolyolva_b200 import runner
    from unav_yolyolva_b200.ingest import DeviceCollator
    from unav_yolyolva_b200.utils import ANETdetection
    dev = torch.device("cuda", 0)
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model = model.to(dev).eval()
    sizes, first, lists, batches = [4, 4, 4, 2], 0, [], []
    for n in sizes:
        lists.append(synth.make_items(n, first))
        batches.append(synth.make_batch(n, 224, first_index=first, with_gt=False))
        first += n
    # synchronous reference loop -> detections; ground truth = its top detections, so mAP is not trivially zero
    sync = {"video-id": [], "t-start": [], "t-end": [], "label": [], "score": []}
    db = {}
    for b in batches:
        res, _ = model(b)
        for i, vid in enumerate(b["video_id"]):
            seg, lab, sc = res["segments"][i].cpu(), res["labels"][i].cpu(), res["scores"][i].cpu()
            sync["video-id"].extend([vid] * seg.shape[0])
            sync["t-start"].append(seg[:, 0]); sync["t-end"].append(seg[:, 1]); sync["label"].append(lab); sync["score"].append(sc)
            db[vid] = {"subset": "test", "duration": b["duration"][i],
                       "annotations": [{"segment": [float(seg[r, 0]), float(seg[r, 1]) + 0.3], "label_id": int(lab[r]), "label": str(int(lab[r]))}
                                       for r in (0, 2, 5)]}
    for k in ("t-start", "t-end", "label", "score"):
        sync[k] = torch.cat(sync[k]).numpy()
    jf = os.path.join(tmp_path, "ants.json")
    json.dump({"database": db}, open(jf, "w"))
    ev = ANETdetection(jf, "test", tiou_thresholds=np.linspace(0.1, 0.9, 9), device=dev)
    _, want = ev.evaluate(sync, verbose=False)
    ap_want = ev.ap.copy()
    mAP, losses = runner.valid_one_epoch(iter(lists), model, 0, evaluator=ev, collate=DeviceCollator(224, dev), print_freq=100)
    assert mAP == want and np.array_equal(ev.ap, ap_want) and mAP > 0.05
    assert set(losses) >= {"cls_loss", "reg_loss", "final_loss"}
