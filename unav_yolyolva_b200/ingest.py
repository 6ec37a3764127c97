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

from typing import Iterable, Iterator, Optional

import torch

_TENSOR_KEYS = ("visual", "audio", "mask")


class CudaPrefetcher:
    """Iterator of device-resident collate dicts, one batch ahead of the consumer.

    The device side is ``depth`` persistent slots per tensor key (no allocator traffic in steady state: a fresh
    ``tensor.to(device)`` per batch makes the caching allocator cudaMalloc whenever the cross-stream free of the previous
    block has not retired yet, which shows up as multi-millisecond stalls).  Slot reuse is ordered by events: the upload
    of batch j+depth waits, on the copy stream, for everything the consumer enqueued for batch j.
    """

    def __init__(self, batches: Iterable[dict], device, keys=_TENSOR_KEYS, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CudaPrefetcher needs a CUDA device")
        if depth < 2:
            raise ValueError("CudaPrefetcher needs at least two slots")
        self.it: Iterator[dict] = iter(batches)
        self.keys = keys
        self.depth = depth
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [dict() for _ in range(depth)]          # slot -> {key: device tensor}
        self._free = [None] * depth                            # slot -> event: consumer done with the slot
        self._n = 0                                            # batches uploaded so far
        self._ready: Optional[tuple] = None
        self._preload()

    def _slot_tensor(self, slot: int, key: str, like: torch.Tensor) -> torch.Tensor:
        cur = self._slots[slot].get(key)
        if cur is None or cur.shape != like.shape or cur.dtype != like.dtype:
            cur = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._slots[slot][key] = cur
        return cur

    def _preload(self) -> None:
        try:
            batch = next(self.it)
        except StopIteration:
            self._ready = None
            return
        slot = self._n % self.depth
        self._n += 1
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
