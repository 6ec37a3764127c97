"""Where does the direct (pinned-arena) upload path spend its time?  Isolated from the model."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.ingest import CudaPrefetcher, DeviceCollator
dev = torch.device("cuda", 0)
N = 512
items = [synth.make_items(1, i)[0] for i in range(N)]
total = sum(it["feats"]["visual"].numel() + it["feats"]["audio"].numel() for it in items)
arena = torch.empty(total, dtype=torch.float32).pin_memory()
o = 0
pinned_items = []
for it in items:
    f = {}
    for k in ("visual", "audio"):
        t = it["feats"][k]
        d = arena[o:o + t.numel()].view(t.shape); d.copy_(t); f[k] = d; o += t.numel()
    pinned_items.append(dict(it, feats=f))
stage = torch.empty(40 << 20 >> 2, dtype=torch.float32, device=dev)
s = torch.cuda.Stream(dev)
def blocks(batch):
    pos = 0
    for it in batch:
        for k in ("visual", "audio"):
            t = it["feats"][k]; n = t.numel()
            stage[pos:pos + n].copy_(t.view(-1), non_blocking=True); pos += n
    return pos
for name, src in (("pinned arena views", pinned_items),):
    for rep in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            a.record(s)
            for j in range(0, N, 16):
                blocks(src[j:j + 16])
            b.record(s)
        th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"{name}: {N // 16} batches of 16: host enqueue {th * 1e3 / (N // 16):.3f} ms/batch, device {a.elapsed_time(b) / (N // 16):.3f} ms/batch", flush=True)
print("is_pinned of a view:", pinned_items[0]["feats"]["visual"].is_pinned())
t0 = time.perf_counter()
for it in pinned_items[:64]:
    it["feats"]["visual"].is_pinned()
print("is_pinned cost us:", (time.perf_counter() - t0) / 64 * 1e6)
for pinned in (False, True):
    lists = [(pinned_items if pinned else items)[j:j + 16] for j in range(0, N, 16)]
    for threaded in (True, False):
        coll = DeviceCollator(224, dev)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 0
            for got in CudaPrefetcher(iter(lists), dev, collate=coll, depth=4, pack_thread=threaded):
                n += 1
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"collator pinned={pinned} thread={threaded}: {dt * 1e3 / n:.3f} ms/batch", flush=True)
