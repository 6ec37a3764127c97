"""BASELINE.json config 4 (long-sequence stress: max_seq_len = 2304, all FPN levels) as a throughput figure:
    python scripts/config4_run.py [batch]          UNAV_ATTN_LONG_SIMT=1 puts the Tk > 256 attention back on the CUDA-core kernel (A/B)
videos/s of the whole path (Alignment + backbone + heads + decode + soft-NMS), device-resident inputs, CUDA events, L2 flushed
between steps; plus the per-kernel-class split of one traced pass."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K
from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.config import default_model_cfg
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch

T = 2304
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg(max_seq_len=T))
manifest = {k: list(v.shape) for k, v in model.state_dict().items()}
model.load_state_dict(synth.trained_like_state_dict(manifest), strict=True)
model = model.to(dev).eval()
b = synth.make_batch(B, T, first_index=3, len_lo=1200, len_hi=2304, with_gt=False)
vis, aud, msk = b["visual"].to(dev), b["audio"].to(dev), b["mask"].to(dev)
meta = torch.tensor([[float(b[k][i]) for k in ("feat_stride", "feat_num_frames", "fps", "duration")] for i in range(B)],
                    dtype=torch.float32, device=dev)
eng = model.engine
for _ in range(2):
    eng.run(vis, aud, msk, meta)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n = 8
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    flush.zero_()
    eng.run(vis, aud, msk, meta)
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / n
K.start_trace()
model.use_cuda_graph = False
eng2 = model.engine
eng2.run(vis, aud, msk, meta)
eng2.run(vis, aud, msk, meta)
K.stop_trace()
K.start_trace()
eng2.run(vis, aud, msk, meta)
tr = K.stop_trace()
agg = {}
for name, t, fl, by, note in tr:
    x = agg.setdefault(name, [0.0, 0.0, 0])
    x[0] += t; x[1] += fl; x[2] += 1
tot = sum(v[0] for v in agg.values())
print(json.dumps({"workload": f"configs[3]: max_seq_len {T}, batch {B}, all 6 FPN levels, valid lengths 1200..2304", "videos_per_s": B / (ms / 1e3),
                  "ms_per_batch": ms, "attention_long": "CUDA cores" if os.environ.get("UNAV_ATTN_LONG_SIMT") == "1" else "tcgen05 (256-key chunks + merge)",
                  "eager_traced_ms": tot,
                  "per_kernel_ms": {k: {"ms": round(v[0], 3), "launches": v[2], "tflops": round(v[1] / (v[0] / 1e3) / 1e12, 1) if v[1] else None}
                                    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]}}))
