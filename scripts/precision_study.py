"""CPU precision study with the oracle's operand-rounding hook: which stages of the path tolerate 16-bit tensor-core
operands (single MMA pass) and which need the 3-pass split, measured as max|err| / max|ref| on logits and offsets.
    python scripts/precision_study.py"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model_ref as R
from unav_yolyolva_b200 import synth

torch.set_num_threads(min(16, os.cpu_count() or 1))
sd = synth.trained_like_state_dict()
b = synth.make_batch(2, 224)

STAGE = ["other"]
def staged(name, fn):
    def w(*a, **k):
        STAGE.append(name)
        try:
            return fn(*a, **k)
        finally:
            STAGE.pop()
    return w
for nm in ("alignment", "fusion_module", "heads", "backbone"):
    setattr(R, nm, staged(nm, getattr(R, nm)))

def bf16(x): return x.bfloat16().float()
def fp16(x): return x.half().float()
def tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)

def run(rounder, low):
    def hook(x):
        return rounder(x) if STAGE[-1] in low else x
    R.OPERAND_ROUND = hook if rounder else None
    with torch.no_grad():
        lg, of, _ = R.forward_logits(sd, b["visual"], b["audio"], b["mask"])
    R.OPERAND_ROUND = None
    return torch.cat(lg, 1), torch.cat(of, 1)

ref_l, ref_o = run(None, ())
rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
ALL = ("alignment", "backbone", "fusion_module", "heads")
print("stage names: backbone = embed/stem/pyramid blocks (fusion_module is nested inside it and counted separately)")
for name, rd in (("bf16", bf16), ("fp16", fp16), ("tf32", tf32)):
    l, o = run(rd, ALL)
    print(f"{name:5s} everywhere             logits {rel(l, ref_l):.2e}  offsets {rel(o, ref_o):.2e}")
for name, rd in (("bf16", bf16), ("fp16", fp16)):
    for st in ALL:
        l, o = run(rd, (st,))
        print(f"{name:5s} only in {st:14s}  logits {rel(l, ref_l):.2e}  offsets {rel(o, ref_o):.2e}")
