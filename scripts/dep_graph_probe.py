"""use_dependency=True variant: eager vs CUDA-graph replay of the whole forward (UNAV_DEP_GRAPH=1), same bits? time per batch?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unav_yolyolva_b200 import synth
from unav_yolyolva_b200.config import default_model_cfg
from unav_yolyolva_b200.modeling import make_multimodal_meta_arch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = default_model_cfg(); cfg["use_dependency"] = True
model = make_multimodal_meta_arch("LocPointTransformer", **cfg)
sd = synth.trained_like_state_dict()
for k, v in model.state_dict().items():
    if k.startswith("dependency_block."):
        sd[k] = synth.trained_like_tensor(k, list(v.shape))
model.load_state_dict(sd, strict=True)
model = model.cuda().eval()
b = synth.make_batch(B, 224, with_gt=False)
outs = {}
for graph in ("0", "1"):
    os.environ["UNAV_DEP_GRAPH"] = graph
    model.invalidate_engine()
    for _ in range(3):
        plan = model.run_hot_path(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        plan = model.run_hot_path(b)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    outs[graph] = (plan["logits"].clone(), plan["out_scores"].clone())
    print(f"UNAV_DEP_GRAPH={graph}: {dt * 1e3:.2f} ms per batch of {B} ({B / dt:.0f} videos/s), graph={model.engine.use_graph}", flush=True)
print("same logits bits:", torch.equal(outs["0"][0], outs["1"][0]), " same scores:", torch.equal(outs["0"][1], outs["1"][1]))
