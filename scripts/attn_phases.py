"""Per-CTA phase timeline of the tcgen05 attention kernel (unav_set_phase_trace).
    python scripts/attn_phases.py <nb> <T> <nh> <hs>"""
import sys, os, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K, _cabi

nb, T, nh, hs = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda", 0)
op = K.BF16X2
C = nh * hs
q, k, v = (torch.randn(nb, T, C) for _ in range(3))
qo = K.pack_operand(q.reshape(-1, C).to(dev), op)
ko = K.pack_operand(k.reshape(-1, C).to(dev), op)
vt = K.new_operand(nb * C, T, op, dev)
K.transpose_cast(v.reshape(-1, C).to(dev).contiguous(), C, vt, nb, T, C, op)
kmask = torch.ones(nb, T, dtype=torch.uint8, device=dev)
out = K.new_operand(nb * T, C, op, dev)
grp = [{"q": qo, "k": ko, "vt": vt, "kmask": kmask, "out": out}]
for _ in range(3):
    K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op)
torch.cuda.synchronize()
# warm graph timing
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(20):
        K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr.replay(); torch.cuda.synchronize(); a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
print(f"[{nb},{nh},{T},{T},{hs}] warm {a.elapsed_time(b) / 20 * 1e3:.1f} us per launch")
cap = 4096
buf = torch.zeros(cap, 8, dtype=torch.int64, device=dev)
lib = _cabi.load()
lib.unav_set_phase_trace(ctypes.c_void_p(buf.data_ptr()), cap)
K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op)
torch.cuda.synchronize()
lib.unav_set_phase_trace(None, 0)
t = buf.cpu(); t = t[t[:, 1] != 0]
d = lambda i, j: (t[:, j] - t[:, i]).double()
for nm, i, j in [("setup incl. TMEM alloc", 1, 2), ("Q,K load + S MMA", 2, 3), ("softmax + P store", 3, 4), ("V load wait + PV MMA", 4, 5),
                 ("epilogue (warp 2)", 5, 6), ("tail", 6, 7), ("CTA total", 1, 7)]:
    x = d(i, j)
    print(f"  {nm:28s} mean {x.mean():9.0f} clk  min {x.min():8.0f}  max {x.max():8.0f}")
sm = t[:, 0]; spans = []
for s_ in sm.unique():
    r = t[sm == s_]; spans.append((int(r[:, 7].max() - r[:, 1].min()), r.shape[0]))
spans.sort()
print("  ctas", t.shape[0], "per-SM span (clk, ctas): min", spans[0], "median", spans[len(spans) // 2], "max", spans[-1], " SMs used", len(spans))
