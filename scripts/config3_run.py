"""BASELINE.json config 3 on N GPUs of one box: a UnAV-100 test-split-sized synthetic workload (2 158 videos)
sharded by video index, one all-gather of the detections, mAP on rank 0 with the device evaluator.

    python scripts/config3_run.py                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/config3_run.py

Prints one JSON line on rank 0: videos/s of the sharded pass (host features already in memory; device collate, upload,
forward, decode, soft-NMS, gather inside the timed region; max over ranks) and of the mAP evaluation, plus a checksum of the
gathered detections that must not depend on N.
"""
import json
import os
import sys
import tempfile
import time
import zlib

os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")      # before torch / libgomp initialise (see bench.py)

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unav_yolyolva_b200 import runner, synth  # noqa: E402
from unav_yolyolva_b200.config import default_model_cfg  # noqa: E402
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch  # noqa: E402
from unav_yolyolva_b200.utils import ANETdetection  # noqa: E402


def main():
    N = int(os.environ.get("UNAV_N_VIDEOS", 2158))
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    model.load_state_dict(synth.trained_like_state_dict(), strict=True)
    model = model.to(dev).eval()
    torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // max(1, world))))
    BS = int(os.environ.get("UNAV_CONFIG3_BATCH", 16))
    res = runner.split_benchmark(model, N, BS, dev, rank, world)
    mine = runner.shard_indices(N, rank, world)
    cache = {i: synth.make_items(1, i)[0] for i in mine}
    load = lambda idxs: [cache[i] for i in idxs]
    dets, valid = runner.evaluate_split(model, N, load, batch_size=BS)
    torch.cuda.synchronize()
    dt = torch.tensor([res["pass_s"]], device=dev)
    if rank == 0:
        d = dets.cpu()
        ids = [f"synth_{i:06d}" for i in range(N)]
        db = {}
        for v in range(N):
            L = synth.video_length(v)
            db[ids[v]] = {"subset": "test", "duration": (L * 8 + 24) / 25.0, "annotations": [
                {"segment": [float(d[v, r, 0]), float(d[v, r, 1]) * 1.05 + 0.1], "label_id": int(d[v, r, 3]), "label": str(int(d[v, r, 3]))}
                for r in (0, 3, 8)]}
        with tempfile.TemporaryDirectory() as td:
            jf = os.path.join(td, "gt.json")
            json.dump({"database": db}, open(jf, "w"))
            ev = ANETdetection(jf, "test", tiou_thresholds=np.linspace(0.1, 0.9, 9), device=dev)
            t1 = time.perf_counter()
            mAPs, avg = ev.evaluate(runner.detections_to_anet(dets, ids), verbose=False)
            t_map = time.perf_counter() - t1
        print(json.dumps({"workload": "config 3: %d synthetic videos sharded i %% world" % N, "n_gpus": world, "batch_per_step": BS,
                          "videos_per_s": N / float(dt), "pass_s": float(dt), "all_gathered": bool(valid.all()),
                          "checksum_crc32": zlib.crc32(d.numpy().tobytes()), "breakdown": res["rank0_breakdown"], "map_eval_s": t_map, "avg_mAP": float(avg),
                          "mAP_at_tiou": [float(x) for x in np.atleast_1d(mAPs)]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
