"""N>1 host logic on CPU: world_size-2 gloo run of the shard + all-gather plumbing used by bench.py / the
multi-GPU runner (the data path itself has no collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unav_yolyolva_b200 import runner


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_videos, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = runner.shard_indices(n_videos, rank, world)
    K = 5
    # fake "detections": row v of video v is filled with v so the gathered tensor can be checked exactly
    local = torch.stack([torch.full((K, 4), float(v)) for v in idx]) if idx else torch.zeros(0, K, 4)
    dets, valid = runner.gather_detections(local, torch.tensor(idx, dtype=torch.int64), n_videos)
    ok = bool(valid.all()) and all(bool((dets[v] == v).all()) for v in range(n_videos))
    if rank == 0:
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_videos = 7          # ragged: rank 0 gets 4 videos, rank 1 gets 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_videos, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)


def test_shard_indices_partition():
    for n, w in [(2158, 8), (7, 2), (3, 4), (0, 2)]:
        parts = [runner.shard_indices(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) == runner.padded_shard_len(n, w) or n == 0


def test_anet_packing_roundtrip():
    segs = torch.rand(3, 4, 2); scores = torch.rand(3, 4); labels = torch.randint(0, 100, (3, 4))
    d = runner.pack_detections(segs, scores, labels)
    r = runner.detections_to_anet(d, ["a", "b", "c"])
    assert len(r["video-id"]) == 12 and r["label"].tolist() == labels.reshape(-1).tolist()


def test_anet_packing_drops_padding_rows_and_invalid_videos():
    """Zero-score padding rows (a video with fewer than K detections) and videos that never arrived (valid mask of
    gather_detections) must not become class-0 predictions at [0, 0]."""
    segs = torch.rand(3, 4, 2) + 0.1; scores = torch.rand(3, 4) + 0.01; labels = torch.randint(0, 100, (3, 4))
    d = runner.pack_detections(segs, scores, labels)
    d[1, 2:] = 0                                   # video b produced only 2 detections
    r = runner.detections_to_anet(d, ["a", "b", "c"], valid=torch.tensor([True, True, False]))
    assert r["video-id"] == ["a"] * 4 + ["b"] * 2
    assert r["score"].tolist() == d[0, :, 2].tolist() + d[1, :2, 2].tolist()
    assert (r["score"] > 0).all()


def test_tag_last_lookahead():
    """runner.tag_last: items pass through unchanged, flags[i] is known before item i is consumed, only the last is True."""
    for n in (0, 1, 2, 5):
        flags, seen = [], []
        for i, x in enumerate(runner.tag_last(iter(range(n)), flags)):
            assert len(flags) == i + 1          # the flag of item i exists when item i arrives
            seen.append(x)
        assert seen == list(range(n))
        assert flags == [False] * (n - 1) + [True] * (1 if n else 0)
