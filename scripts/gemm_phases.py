"""Per-CTA phase timeline of the tcgen05 GEMM (unav_set_phase_trace): where a tile's life goes.
    python scripts/gemm_phases.py <G> <M> <N> <K> [act]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from unav_yolyolva_b200 import kernels as K, _cabi

G, M, N, Kd = (int(x) for x in sys.argv[1:5])
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda", 0)
op = K.BF16X2
groups = []
for g in range(G):
    A = K.new_operand(M, Kd, op, dev); A.normal_()
    W = K.new_operand(N, Kd, op, dev); W.normal_()
    groups.append({"A": A, "W": W, "bias": torch.zeros(N, device=dev), "out_op": K.new_operand(M, N, op, dev)})
for _ in range(3):
    K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05)
torch.cuda.synchronize()
cap = 4096
buf = torch.zeros(cap, 8, dtype=torch.int64, device=dev)
lib = _cabi.load()
lib.unav_set_phase_trace(ctypes.c_void_p(buf.data_ptr()), cap)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05)
b.record()
torch.cuda.synchronize()
lib.unav_set_phase_trace(None, 0)
t = buf.cpu()
t = t[t[:, 1] != 0]
n = t.shape[0]
d = lambda i, j: (t[:, j] - t[:, i]).double()
names = [("setup (alloc, barriers)", 1, 2), ("first operands land", 2, 3), ("k-loop issue (first land -> last MMA issued)", 3, 4),
         ("last issue -> accumulator ready", 4, 5), ("epilogue (first warp)", 5, 6), ("epilogue tail + sync", 6, 7), ("CTA total", 1, 7)]
print(f"{G}x[{M},{N},{Kd}] ctas={n} kernel {a.elapsed_time(b) * 1e3:.1f} us (eager, includes launch)")
for nm, i, j in names:
    x = d(i, j)
    print(f"  {nm:48s} mean {x.mean():9.0f} clk  min {x.min():8.0f}  max {x.max():8.0f}")
# per-SM occupancy timeline: span of all CTAs on the busiest SM
sm = t[:, 0]
spans = []
for s_ in sm.unique():
    r = t[sm == s_]
    spans.append((int(r[:, 7].max() - r[:, 1].min()), r.shape[0]))
spans.sort()
print("  per-SM span (clk, ctas): min", spans[0], "median", spans[len(spans) // 2], "max", spans[-1], " SMs used", len(spans))
