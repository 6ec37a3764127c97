"""Pinned, double-buffered upload of collate dicts (what replaces ``DataParallel.scatter``).

The reference moves every batch host->device synchronously inside ``nn.DataParallel.scatter`` right before the forward
(/root/reference/eval.py:61, libs/utils/train_utils.py:409).  ``CudaPrefetcher`` wraps any iterable of collate dicts
(/root/reference/libs/datasets/data_utils.py:214-229) and uploads batch j+1 on a side stream while batch j computes:

    for batch in CudaPrefetcher(val_loader, device):      # drop-in around the reference's DataLoader
        results, losses = model(batch)

Tensors are pinned (if the loader did not already use pin_memory) and copied with ``non_blocking=True`` into persistent
device slots; every other entry of the dict (lists of Python scalars / strings, ground truth) passes through unchanged.
The yielded tensors alias the slots: they stay valid until ``depth - 1`` further batches have been drawn, which is what a
``for batch in ...: model(batch)`` loop needs (the engine copies them into its static inputs at the start of the step).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional

import torch

from . import kernels as K

_TENSOR_KEYS = ("visual", "audio", "mask")


class CudaPrefetcher:
    """Iterator of device-resident collate dicts, one batch ahead of the consumer.

    The device side is ``depth`` persistent slots per tensor key (no allocator traffic in steady state: a fresh
    ``tensor.to(device)`` per batch makes the caching allocator cudaMalloc whenever the cross-stream free of the previous
    block has not retired yet, which shows up as multi-millisecond stalls).  Slot reuse is ordered by events: the upload
    of batch j+depth waits, on the copy stream, for everything the consumer enqueued for batch j.
    """

    def __init__(self, batches: Iterable, device, keys=_TENSOR_KEYS, depth: int = 2, collate: "Optional[DeviceCollator]" = None,
                 pack_thread: bool = True):
        """batches: collate dicts — or, with ``collate=DeviceCollator(...)``, the raw per-batch lists of dataset items
        (a DataLoader built with ``collate_fn=lambda items: items``), which are then padded on the device.
        ``pack_thread`` (with ``collate``): the host half of the collate — packing the videos' feature blocks into the pinned
        staging slot, ~16 MB of copies per batch of 16 — runs in a background thread up to ``depth - 1`` batches ahead, so the
        consumer's thread only enqueues the upload and the two pad kernels (the copies release the GIL).  With a short shard
        per GPU (configs[2] on 8 GPUs: 17 batches per rank) the host, not the device, is what bounds the pass."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CudaPrefetcher needs a CUDA device")
        if depth < 2:
            raise ValueError("CudaPrefetcher needs at least two slots")
        self.it: Iterator[dict] = iter(batches)
        self.keys = keys
        self.depth = depth
        self.collate = collate
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [dict() for _ in range(depth)]          # slot -> {key: device tensor}
        self._free = [None] * depth                            # slot -> event: consumer done with the slot
        self._n = 0                                            # batches uploaded so far
        self._ready: Optional[tuple] = None
        self._queue = None
        if collate is not None and pack_thread:
            import queue
            import threading
            self._queue = queue.Queue(maxsize=depth - 1)
            # a pinned slot may be repacked once the upload of its previous use has been ENQUEUED (then its event exists and
            # the packer waits for it): one permit per slot, returned by the consumer's thread after it recorded that event
            self._permits = [threading.Semaphore(1) for _ in range(depth)]
            self._thread = threading.Thread(target=self._pack_loop, name="unav-collate-pack", daemon=True)
            self._thread.start()
        self._preload()

    def _pack_loop(self) -> None:
        n = 0
        try:
            for batch in self.it:
                slot = n % self.depth
                n += 1
                self._permits[slot].acquire()
                self._queue.put((self.collate.pack_host(batch, slot), slot))
        except BaseException as e:                             # noqa: BLE001 - re-raised in the consumer's thread
            self._queue.put((e, -1))
            return
        self._queue.put((None, -1))

    def _slot_tensor(self, slot: int, key: str, like: torch.Tensor) -> torch.Tensor:
        cur = self._slots[slot].get(key)
        if cur is None or cur.shape != like.shape or cur.dtype != like.dtype:
            cur = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._slots[slot][key] = cur
        return cur

    def _preload(self) -> None:
        if self._queue is not None:
            packed, slot = self._queue.get()
            if packed is None:
                self._ready = None
                return
            if isinstance(packed, BaseException):
                self._ready = None
                raise packed
            self._n += 1
            with torch.cuda.stream(self.stream):
                if self._free[slot] is not None:
                    self.stream.wait_event(self._free[slot])
                out = self.collate.enqueue_device(packed, slot)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._permits[slot].release()
            self._ready = (out, ev, slot)
            return
        try:
            batch = next(self.it)
        except StopIteration:
            self._ready = None
            return
        slot = self._n % self.depth
        self._n += 1
        if self.collate is not None:
            with torch.cuda.stream(self.stream):
                if self._free[slot] is not None:
                    self.stream.wait_event(self._free[slot])
                out = self.collate.enqueue(batch, slot)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._ready = (out, ev, slot)
            return
        out = dict(batch)
        # device buffers are created on the consumer's stream (plain allocations, made once per shape)
        dst = {}
        for k in self.keys:
            t = out.get(k)
            if torch.is_tensor(t) and not t.is_cuda:
                if not t.is_pinned():
                    t = t.pin_memory()
                dst[k] = (self._slot_tensor(slot, k, t), t)
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])      # the consumer has finished with this slot's last batch
            for k, (d, t) in dst.items():
                d.copy_(t, non_blocking=True)
                out[k] = d
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready = (out, ev, slot)

    def __iter__(self):
        return self

    def __next__(self) -> dict:
        if self._ready is None:
            raise StopIteration
        batch, ev, slot = self._ready
        cur = torch.cuda.current_stream(self.device)
        # everything the consumer enqueued so far belongs to earlier batches: the slot handed out `depth - 1` batches
        # ago becomes reusable once that work retires
        prev = (slot - 1) % self.depth
        done = torch.cuda.Event()
        done.record(cur)
        self._free[prev] = done
        cur.wait_event(ev)                       # compute stream waits for this batch's upload only
        self._preload()                          # start uploading the next batch while this one computes
        return batch


class DeviceCollator:
    """``collate_fcn`` of the reference (/root/reference/libs/datasets/data_utils.py:123-229) for inference, with the
    padding and the mask done on the device.

    The reference pads every video to ``max_seq_len`` on the host (a Python loop of ``copy_`` per video, :178-198) and the
    padded ``[B, 2048+128, T]`` batch then crosses PCIe.  Here only the valid frames do: the videos' ``[C, len]`` feature
    blocks are packed back to back into one pinned staging buffer, uploaded with one copy, and ``unav_collate_pad`` writes
    the padded ``visual`` / ``audio`` tensors and the ``mask`` (``arange(T) < len``, :201).  Output keys are the ones the
    inference path reads (visual, audio, mask, video_id, fps, duration, feat_stride, feat_num_frames); the loss-only keys of
    the reference's dict (scores, start_end, m_labels, gt_*, points) are not produced.

    ``collator(items)`` enqueues on the current stream and returns device tensors that alias slot 0;
    ``CudaPrefetcher(loader, device, collate=collator)`` runs it one batch ahead on the copy stream.
    """

    def __init__(self, max_seq_len: int, device, max_div_factor: int = 1, padding_val: float = 0.0, direct_pinned: bool = False):
        # direct_pinned: upload features that already live in pinned host memory straight from where they lie (one async copy
        # per tensor) instead of packing them into the staging slot first.  Opt-in: in isolation it is 2x cheaper on the host
        # (scripts/h2d_probe.py: 0.5 vs 1.1 ms per batch of 16), but inside the three-batch pipeline the ~33 small copies per
        # batch queue on the copy engine between the engine's own input copies and made whole passes erratically 2x slower
        # (0.62 s / 1.28 s for the same 2 158-video pass on one B200), so the one-copy-per-batch packed path stays the default.
        self.direct_pinned = bool(direct_pinned)
        self.T = int(max_seq_len)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceCollator needs a CUDA device (no CPU fallback)")
        self.max_div_factor = int(max_div_factor)
        self.pad = float(padding_val)
        self._host: dict = {}                                   # slot -> pinned staging (header + feature blocks)
        self._dev: dict = {}                                    # slot -> device staging + padded outputs
        self._uploaded: dict = {}                               # slot -> event: the staging buffer's last upload has been read

    def pack_host(self, video_list: List[dict], slot: int = 0) -> dict:
        """Host half: pack the videos' feature blocks (+ offsets / lengths header) into the slot's pinned staging buffer.
        Thread-safe with respect to ``enqueue_device`` of OTHER slots; waits for the slot's previous upload to retire."""
        B = len(video_list)
        vis = [x["feats"]["visual"] for x in video_list]
        aud = [x["feats"]["audio"] for x in video_list]
        lens = [int(v.shape[-1]) for v in vis]
        Cv, Ca = int(vis[0].shape[0]), int(aud[0].shape[0])
        max_len = max(lens)
        if max_len <= self.T:                                          # data_utils.py:170-176 (eval branch)
            T = self.T
        else:
            st = self.max_div_factor
            T = (max_len + (st - 1)) // st * st
        nfloat = sum(lens) * (Cv + Ca)
        hdr = 4 * B
        meta = {k: [x[k] for x in video_list] for k in ("video_id", "fps", "duration", "feat_stride", "feat_num_frames")}
        base = 5 * B + (-5 * B) % 4                                    # first payload float, 16-byte aligned
        # Features that already live in PINNED host memory (a dataset cache resident in a pinned arena) are uploaded straight
        # from where they lie, one async copy per tensor: only the small header goes through the staging slot.  Repacking them
        # first costs a host memcpy of every byte, which on an 8-GPU box (8 ranks sharing the host's cores and memory
        # bandwidth) is what bounds a short shard (measured: 5 ms per batch of 16 against 3.4 ms of device work).
        direct = self.direct_pinned and all(t.is_pinned() and t.is_contiguous() and t.dtype == torch.float32 for t in vis + aud)
        need = (base if direct else nfloat + 2 * hdr + 2 * B)
        h = self._host.get(slot)
        if h is None or h.numel() < need:
            h = torch.empty(int(need * 1.25) + 1024, dtype=torch.float32).pin_memory()
            self._host[slot] = h
        if slot in self._uploaded:                                     # the host must not rewrite pinned memory a copy still reads
            self._uploaded[slot].synchronize()
        # header: [offsets_v i64 x B | offsets_a i64 x B | lens i32 x B], then the feature blocks
        hb = h.view(torch.uint8)
        off_v = hb[0:8 * B].view(torch.int64)
        off_a = hb[8 * B:16 * B].view(torch.int64)
        ln = hb[16 * B:20 * B].view(torch.int32)
        pos = base
        blocks = []
        for i, (v, L) in enumerate(zip(vis, lens)):
            n = Cv * L
            if direct:
                blocks.append((pos, n, v))
            else:
                h[pos:pos + n].view(Cv, L).copy_(v)
            off_v[i] = pos
            pos += n
        for i, (a, L) in enumerate(zip(aud, lens)):
            if int(a.shape[-1]) != L:
                raise ValueError("visual and audio features of a video must have the same length")
            n = Ca * L
            if direct:
                blocks.append((pos, n, a))
            else:
                h[pos:pos + n].view(Ca, L).copy_(a)
            off_a[i] = pos
            ln[i] = L
            pos += n
        return {"B": B, "Cv": Cv, "Ca": Ca, "T": T, "pos": pos, "host": h, "meta": meta, "blocks": blocks if direct else None,
                "hdr_floats": base}

    def enqueue_device(self, packed: dict, slot: int = 0) -> dict:
        """Device half, on the current stream: one upload of the packed staging buffer + the two pad kernels."""
        B, Cv, Ca, T, pos, h = (packed[k] for k in ("B", "Cv", "Ca", "T", "pos", "host"))
        d = self._dev.setdefault(slot, {})
        if d.get("key") != (B, Cv, Ca, T) or d["stage"].numel() < max(h.numel(), pos):
            d.clear()
            d["key"] = (B, Cv, Ca, T)
            # worst case (every video T frames long) + header: sized once per batch shape, never regrown inside a pass
            d["stage"] = torch.empty(max(h.numel(), pos, B * T * (Cv + Ca) + 8 * B + 64), dtype=torch.float32, device=self.device)
            d["visual"] = torch.empty(B, Cv, T, dtype=torch.float32, device=self.device)
            d["audio"] = torch.empty(B, Ca, T, dtype=torch.float32, device=self.device)
            d["mask"] = torch.empty(B, 1, T, dtype=torch.uint8, device=self.device)
        st = d["stage"]
        if packed.get("blocks") is None:
            st[:pos].copy_(h[:pos], non_blocking=True)
        else:                                                          # header from the slot, features from where they lie
            nh = packed["hdr_floats"]
            st[:nh].copy_(h[:nh], non_blocking=True)
            for o, n, t in packed["blocks"]:
                st[o:o + n].copy_(t.view(-1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._uploaded[slot] = ev
        sb = st.view(torch.uint8)
        d_off_v, d_off_a, d_ln = sb[0:8 * B].view(torch.int64), sb[8 * B:16 * B].view(torch.int64), sb[16 * B:20 * B].view(torch.int32)
        mask_u8 = d["mask"].view(B, T)
        K.collate_pad(st, d_off_v, d_ln, d["visual"], mask_u8, B, Cv, T, self.pad)
        K.collate_pad(st, d_off_a, d_ln, d["audio"], None, B, Ca, T, self.pad)
        return dict({"visual": d["visual"], "audio": d["audio"], "mask": d["mask"].view(torch.bool)}, **packed["meta"])

    def enqueue(self, video_list: List[dict], slot: int = 0) -> dict:
        return self.enqueue_device(self.pack_host(video_list, slot), slot)

    __call__ = enqueue
