"""Multi-GPU plumbing: one process per GPU, videos sharded by index, one all-gather of the detections.

The hot path has no cross-video operation, so ranks never exchange activations; the only
collective is the final all-gather of the fixed-shape detection tensor for mAP (reference behaviour for
comparison: ``nn.DataParallel`` re-broadcasts all 97 M parameters every forward, /root/reference/eval.py:61).
Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU tensors (tests/test_multirank_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

import torch
import torch.distributed as dist


def shard_indices(n_videos: int, rank: int, world: int) -> List[int]:
    """Videos {i : i % world == rank} (strided so every rank sees the same length distribution)."""
    return list(range(rank, n_videos, world))


def padded_shard_len(n_videos: int, world: int) -> int:
    return (n_videos + world - 1) // world


def pack_detections(segments: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """[n,K,2], [n,K], [n,K] -> [n,K,4] float32 (seg0, seg1, score, label); labels < 2^24 are exact in f32."""
    return torch.cat([segments.float(), scores.float()[..., None], labels.float()[..., None]], dim=-1)


def gather_detections(local: torch.Tensor, video_index: torch.Tensor, n_videos: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank detections.

    local [n_local, K, 4] and video_index [n_local] (int32/64) live on the rank's device.  Every rank pads to
    ceil(n_videos / world) rows (index -1), one all_gather moves [world, n_pad, K, 4] + [world, n_pad]; returns
    (detections [n_videos, K, 4] ordered by video index, valid mask [n_videos]).
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    n_pad = padded_shard_len(n_videos, world)
    K = local.shape[1]
    buf = torch.zeros(n_pad, K, 4, dtype=torch.float32, device=local.device)
    idx = torch.full((n_pad,), -1, dtype=torch.int64, device=local.device)
    n_local = local.shape[0]
    buf[:n_local] = local
    idx[:n_local] = video_index.to(torch.int64)
    if world > 1:
        all_buf = torch.empty(world * n_pad, K, 4, dtype=torch.float32, device=local.device)
        all_idx = torch.empty(world * n_pad, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(all_buf, buf)
        dist.all_gather_into_tensor(all_idx, idx)
    else:
        all_buf, all_idx = buf, idx
    # scatter by video index without boolean indexing (no host sync): padded rows go to a spare slot n_videos
    flat_idx = all_idx.reshape(-1)
    safe = torch.where(flat_idx >= 0, flat_idx, torch.full_like(flat_idx, n_videos))
    out = torch.zeros(n_videos + 1, K, 4, dtype=torch.float32, device=local.device)
    out.index_copy_(0, safe, all_buf.reshape(-1, K, 4))
    valid = torch.zeros(n_videos + 1, dtype=torch.bool, device=local.device)
    valid.index_fill_(0, safe, True)
    out, valid = out[:n_videos], valid[:n_videos]
    return out, valid


def detections_to_anet(dets: torch.Tensor, video_ids: List[str], valid: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """[n,K,4] -> the flat result dict ``valid_one_epoch`` hands to ``ANETdetection.evaluate``
    (/root/reference/libs/utils/train_utils.py:400-449).  Rows that are not detections are dropped: the zero padding of a
    video with fewer than K detections (score 0 — a real detection's score is > pre_nms_thresh > 0) and, with ``valid``
    (the mask returned by ``gather_detections``), videos that never arrived; otherwise they would count as class-0
    predictions at [0, 0]."""
    d = dets.cpu()
    n, K = d.shape[:2]
    keep = d[..., 2] > 0
    if valid is not None:
        keep &= valid.cpu().reshape(n, 1)
    keep = keep.reshape(-1).numpy()
    ids = np.repeat(np.asarray(video_ids, dtype=object), K)
    return {
        "video-id": ids[keep].tolist(),
        "t-start": d[..., 0].reshape(-1).numpy()[keep],
        "t-end": d[..., 1].reshape(-1).numpy()[keep],
        "score": d[..., 2].reshape(-1).numpy()[keep],
        "label": d[..., 3].reshape(-1).long().numpy()[keep],
    }


def tag_last(loader, flags: list):
    """Yield the items of ``loader`` unchanged while appending to ``flags``, in order, whether each one is the last — with one
    HOST item of look-ahead, so that a consumer behind a prefetcher (which holds device slots) learns it without holding a
    slot longer.  ``flags[i]`` is set before item i is yielded."""
    it = iter(loader)
    try:
        prev = next(it)
    except StopIteration:
        return
    for cur in it:
        flags.append(False)
        yield prev
        prev = cur
    flags.append(True)
    yield prev


def valid_one_epoch(val_loader, model, curr_epoch, ext_score_file=None, evaluator=None, output_file=None, tb_writer=None,
                    print_freq=2, device=None, collate=None, losses="last"):
    """Drop-in for the reference's evaluation loop (/root/reference/libs/utils/train_utils.py:380-463) with the three
    serial stages of that loop overlapped: the upload of batch j+1 (``CudaPrefetcher``, optionally with the device-side
    collate), the forward of batch j (``PtTransformer.submit``: soft-NMS and the copy of the detections to pinned host
    memory on a side stream) and the host-side bookkeeping of batch j-1.  Same arguments and return value
    ``(mAP, losses)``; ``model`` may be the bare ``PtTransformer`` or wrapped in ``nn.DataParallel`` (``model.module``).
    ``evaluator``: the reference's ``ANETdetection`` or ``unav_yolyolva_b200.utils.ANETdetection`` (matching on the
    device).  ``ext_score_file`` post-processing is out of scope and raises.
    ``losses``: the reference returns the loss dict of the LAST batch (:466; it evaluates the losses of every batch and drops
    them).  "last" evaluates it for the last batch only, "all" for every batch (then ``loss_normalizer``, which the
    reference also updates in eval, :637-640, follows the reference's state exactly), "none" returns zeros.  Batches without
    GT tensors (the device-side collate does not produce them) give zeros."""
    import pickle
    import time

    import numpy as np

    from .ingest import CudaPrefetcher
    assert (evaluator is not None) or (output_file is not None)
    if ext_score_file is not None:
        raise NotImplementedError("external-score post-processing (libs/utils/postprocessing.py) is not on the hot path")
    net = model.module if hasattr(model, "module") else model
    net.eval()
    dev = torch.device(device) if device is not None else net.device
    results = {"video-id": [], "t-start": [], "t-end": [], "label": [], "score": []}

    def unpack(handle, video_ids):
        out = handle.result()                               # host tensors in pinned memory, this step only
        for vid_idx, vid in enumerate(video_ids):
            n = out["segments"][vid_idx].shape[0]
            if n > 0:
                results["video-id"].extend([vid] * n)
                results["t-start"].append(out["segments"][vid_idx][:, 0].clone())
                results["t-end"].append(out["segments"][vid_idx][:, 1].clone())
                results["label"].append(out["labels"][vid_idx].clone())
                results["score"].append(out["scores"][vid_idx].clone())

    start = time.time()
    pending = []
    depth = max(1, getattr(net, "streams", 1))                 # steps kept in flight before the oldest result is read
    from .losses import GT_KEYS, LOSS_KEYS
    assert losses in ("last", "all", "none")
    last_losses = None

    flags = []                      # is_last per raw batch, in loader order (the prefetcher draws them one ahead)
    tagged = lambda loader: tag_last(loader, flags)
    for iter_idx, batch in enumerate(CudaPrefetcher(tagged(val_loader), dev, collate=collate, depth=depth + 1)):
        is_last = flags[iter_idx]
        want = (losses == "all" or (losses == "last" and is_last)) and all(k in batch for k in GT_KEYS)
        with torch.no_grad():
            h = net.submit(batch, with_losses=want)
            pending.append((h, batch["video_id"]))
            if h.losses is not None:
                last_losses = h.losses
        if len(pending) > depth:
            unpack(*pending.pop(0))
        if iter_idx != 0 and iter_idx % print_freq == 0:
            print("Test: [{0:05d}]\tTime {1:.3f} s / batch".format(iter_idx, (time.time() - start) / print_freq))
            start = time.time()
    while pending:
        unpack(*pending.pop(0))
    for k in ("t-start", "t-end", "label", "score"):
        results[k] = torch.cat(results[k]).numpy() if results[k] else np.zeros(0)
    if evaluator is not None:
        _, mAP = evaluator.evaluate(results, verbose=True)
    else:
        with open(output_file, "wb") as f:
            pickle.dump(results, f)
        mAP = 0.0
    if tb_writer is not None:
        tb_writer.add_scalar("validation/mAP", mAP, curr_epoch)
    if last_losses is None:
        last_losses = {k: torch.zeros((), device=dev) for k in LOSS_KEYS}
    return mAP, last_losses


def run_shard(model, indices: List[int], load_items, batch_size: int = 16, device=None, stats: Optional[dict] = None) -> torch.Tensor:
    """Detections of the videos ``indices`` (this rank's shard) as one device tensor [len(indices), K, 4]
    (seg0, seg1, score, label), through the overlapped public pipeline: ``load_items(list_of_indices)`` returns un-collated
    dataset items (/root/reference/libs/datasets/datasets.py:28-46 layout), ``DeviceCollator`` pads them on the device,
    ``CudaPrefetcher`` uploads batch j+1 under the forward of batch j, ``PtTransformer.submit`` keeps ``model.streams``
    batches in flight.  Videos with fewer than K detections keep zero rows (score 0).
    ``stats`` (optional dict) receives the host-side seconds spent per stage (collate + upload enqueue, submit, waiting for the
    device), the number of batches and the device time from the first submit to the last detection (CUDA events)."""
    import time

    from .ingest import CudaPrefetcher, DeviceCollator
    net = model.module if hasattr(model, "module") else model
    dev = torch.device(device) if device is not None else net.device
    K_ = net.test_max_seg_num
    out = torch.zeros(len(indices), K_, 4, dtype=torch.float32, device=dev)
    chunks = [indices[i:i + batch_size] for i in range(0, len(indices), batch_size)]
    depth = max(1, getattr(net, "streams", 1))
    cur = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(cur)                                       # `out` is zero-filled on the caller's stream
    ar = torch.arange(K_, device=dev)[None]
    pending, tails, row = [], {}, 0
    loader = (load_items(c) for c in chunks)
    # the collator owns pinned staging + device slots: keep ONE per model and device, so that a warm-up pass really warms
    # (a fresh collator per call re-allocated ~100 MB of pinned memory inside every pass: tens of milliseconds, which is what
    # a 17-batch shard on 8 GPUs cannot hide)
    coll = getattr(net, "_shard_collator", None)
    if coll is None or coll.device != dev or coll.T != net.max_seq_len:
        coll = DeviceCollator(net.max_seq_len, dev, max_div_factor=net.max_div_factor)
        net._shard_collator = coll
    t_fetch = t_submit = t_wait = t_tail = 0.0
    if stats is not None:
        net._profile = {}
    ev_first = torch.cuda.Event(enable_timing=True) if stats is not None else None
    it = iter(CudaPrefetcher(loader, dev, collate=coll, depth=depth + 1))
    j = 0
    while True:
        ta = time.perf_counter()
        try:
            batch = next(it)                                # returns batch j, enqueues collate + upload of batch j+1
        except StopIteration:
            break
        tb = time.perf_counter()
        if ev_first is not None and j == 0:
            ev_first.record(cur)
        n = len(chunks[j])
        with torch.no_grad():
            h = net.submit(batch)
            tb2 = time.perf_counter()
            P = h.plan
            ns = P["nms_stream"]
            with torch.cuda.stream(ns):                     # device -> device, behind this step's soft-NMS
                ns.wait_event(ready)
                d = pack_detections(P["out_segs"], P["out_scores"], P["out_labels"])
                out[row:row + n] = d * (ar < P["out_counts"][:, None])[..., None]
                ev = torch.cuda.Event()
                ev.record(ns)
                tails[id(ns)] = ev
        pending.append(h)
        row += n
        tc = time.perf_counter()
        if len(pending) > depth:
            pending.pop(0)._event.synchronize()             # bounds the host's lead over the device
        td = time.perf_counter()
        t_fetch += tb - ta; t_submit += tb2 - tb; t_tail += tc - tb2; t_wait += td - tc
        j += 1
    for ev in tails.values():
        cur.wait_event(ev)
    if stats is not None:
        ev_last = torch.cuda.Event(enable_timing=True)
        ev_last.record(cur)
        ev_last.synchronize()
        stats.update(net._profile)
        net._profile = None
        stats.update({"batches": len(chunks), "host_collate_upload_s": t_fetch, "host_submit_s": t_submit, "host_pack_out_s": t_tail,
                      "host_wait_device_s": t_wait,
                      "device_first_submit_to_last_detection_s": ev_first.elapsed_time(ev_last) / 1e3 if chunks else 0.0})
    return out


def evaluate_split(model, n_videos: int, load_items, batch_size: int = 16, device=None, stats: Optional[dict] = None):
    """BASELINE.json config 3: a whole split sharded by video index over the ranks of the current process group (or one
    process), one all-gather of the detections.  Returns (detections [n_videos, K, 4] ordered by video index, valid [n_videos])
    on every rank; rank 0 hands them to ``ANETdetection.evaluate`` via ``detections_to_anet``."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    idx = shard_indices(n_videos, rank, world)
    local = run_shard(model, idx, load_items, batch_size, device, stats=stats)
    return gather_detections(local, torch.tensor(idx, dtype=torch.int64, device=local.device), n_videos)


def split_benchmark(model, n_videos: int, batch_size: int, device, rank: int, world: int, ms_per_step: Optional[float] = None) -> dict:
    """BASELINE.json configs[2] as a measurement: the test-split-sized synthetic workload sharded by video index over the
    ``world`` ranks through the public pipeline, one all-gather of the detections — STRONG scaling (total work fixed).
    Window: CUDA events on each rank's stream from before the first batch's collate to after the all-gather, max over ranks;
    weights packed, graphs captured, staging buffers allocated and NCCL warmed before it; the synthetic features are generated
    in host memory outside the window.  Every rank must call it; the dict is meaningful on every rank (checksum on rank 0's copy)."""
    import time
    import zlib

    import os

    from . import synth
    dev = torch.device(device)
    net = model.module if hasattr(model, "module") else model
    # Host threads of this rank: the pass is host bound once the ranks share the box's cores, and a torch
    # intra-op pool per rank (OpenMP workers that spin after every parallel region) only adds to it — the packing copies and
    # the launch path are one thread each.  UNAV_SPLIT_THREADS overrides; single-process runs keep their setting.
    nt = int(os.environ.get("UNAV_SPLIT_THREADS", "1" if world > 1 else "0"))
    prev_threads = torch.get_num_threads()
    if nt > 0:
        torch.set_num_threads(nt)
    mine = shard_indices(n_videos, rank, world)
    items = [synth.make_items(1, i)[0] for i in mine]          # this rank's shard of the synthetic dataset, in host memory
    cache = dict(zip(mine, items))
    load = lambda idxs: [cache[i] for i in idxs]
    K_ = net.test_max_seg_num
    warm = mine[:(max(1, getattr(net, "streams", 1)) + 2) * batch_size]
    for _ in range(2):                                        # plans, host rings, every collate slot, allocator
        run_shard(model, warm, load, batch_size=batch_size, device=dev)
    gather_detections(torch.zeros(1, K_, 4, device=dev), torch.tensor([rank], device=dev), world)      # NCCL warm-up
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    st: dict = {}
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    c0.record()
    dets, valid = evaluate_split(model, n_videos, load, batch_size=batch_size, device=dev, stats=st)
    c1.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    tt = torch.tensor([c0.elapsed_time(c1) / 1e3, wall], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt[0])
    out = {"workload": f"configs[2]: {n_videos} synthetic videos sharded i % {world} over {world} GPU(s), detections all-gathered",
           "scaling": "strong", "videos": n_videos, "n_gpus": world, "videos_per_s": n_videos / sec, "pass_s": sec,
           "wall_s_max": float(tt[1]), "all_gathered": bool(valid.all().item()),
           "detections_crc32": zlib.crc32(dets.cpu().numpy().tobytes()),
           "rank0_breakdown": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()},
           "window": "CUDA events from before the first collate to after the all-gather, max over ranks; init / capture / "
                     "staging allocation / NCCL warm-up excluded"}
    if ms_per_step is not None:
        out["ideal_pass_s_at_device_rate"] = (len(mine) / batch_size) * ms_per_step / 1e3
    out["host_threads"] = torch.get_num_threads()
    if nt > 0:
        torch.set_num_threads(prev_threads)
    return out
