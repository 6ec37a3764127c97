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

def split16(x):       # hi + lo, both fp16: what a 3-pass (or the A side of a 2-pass) GEMM sees
    hi = x.half().float()
    return hi + (x - hi).half().float()

def run(cfg):
    """cfg: stage -> passes (1: A_hi.B_hi, 2: (A_hi+A_lo).B_hi, 3: split on both sides, 4: A_hi.(B_hi+B_lo) = two passes with
    only the WEIGHT side split); missing stage = exact."""
    def ha(x):
        n = cfg.get(STAGE[-1])
        return x if n is None else (fp16(x) if n in (1, 4) else split16(x))
    def hb(x):
        n = cfg.get(STAGE[-1])
        return x if n is None else (fp16(x) if n <= 2 else split16(x))
    R.OPERAND_ROUND, R.OPERAND_ROUND_B = (ha, hb) if cfg else (None, None)
    with torch.no_grad():
        lg, of, _ = R.forward_logits(sd, b["visual"], b["audio"], b["mask"])
    R.OPERAND_ROUND = R.OPERAND_ROUND_B = None
    return torch.cat(lg, 1), torch.cat(of, 1)

ref_l, ref_o = run({})
rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
ALL = ("alignment", "backbone", "fusion_module", "heads")
print("fp16-split operands, MMA passes per stage (alignment, backbone, fusion_module, heads):")
import itertools
for combo in [(3,3,3,3),(1,1,1,1),(2,2,2,2),(4,4,4,4),(1,1,2,2),(1,1,4,4),(1,1,3,3),(1,1,2,3),(1,1,4,3),(4,4,3,3),(2,1,2,2),(1,1,1,2),(1,1,2,1),(2,2,3,3),(1,2,2,2),(1,1,3,2)]:
    l, o = run(dict(zip(ALL, combo)))
    print(f"  {combo}   logits {rel(l, ref_l):.2e}  offsets {rel(o, ref_o):.2e}")
