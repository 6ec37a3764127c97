import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.config import default_model_cfg
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch
from unav_yolyolva_b200.ingest import CudaPrefetcher
dev = torch.device("cuda", 0)
model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
model.load_state_dict(synth.trained_like_state_dict(), strict=True)
model = model.to(dev).eval()
B = 16
hb = []
for j in range(4):
    b = synth.make_batch(B, 224, first_index=j * B, with_gt=False)
    for k in ("visual", "audio", "mask"): b[k] = b[k].pin_memory()
    hb.append(b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def loop(n, use_pf, do_flush, d2h, tag):
    src = (hb[j % 4] for j in range(n))
    it = CudaPrefetcher(src, dev) if use_pf else src
    torch.cuda.synchronize(); t0 = time.perf_counter()
    host = 0.0
    for batch in it:
        if do_flush: flush.zero_()
        h0 = time.perf_counter()
        res, _ = model(batch)
        if d2h: out = (res["segments"].cpu(), res["scores"].cpu(), res["labels"].cpu())
        host += time.perf_counter() - h0
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{tag:40s} {dt:7.3f} ms/step (model+d2h host-side {host / n * 1e3:.3f})", flush=True)
for _ in range(2): loop(10, True, True, True, "warm")
loop(30, True, True, True, "prefetch + flush + d2h")
loop(30, True, False, True, "prefetch + d2h (no flush)")
loop(30, False, False, True, "inline H2D + d2h (no flush)")
loop(30, True, False, False, "prefetch, no d2h, no flush")
# raw pieces
plan = model.engine._plans[(B, 0)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): plan["graph"].replay()
torch.cuda.synchronize(); print("graph replay only", (time.perf_counter() - t0) / 30 * 1e3)
t0 = time.perf_counter()
for _ in range(30): flush.zero_()
torch.cuda.synchronize(); print("flush only", (time.perf_counter() - t0) / 30 * 1e3)
x = hb[0]["visual"]; y = torch.empty_like(x, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print("H2D visual only", (time.perf_counter() - t0) / 30 * 1e3)
