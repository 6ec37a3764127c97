"""Per-CTA phase timeline of the persistent CTA-pair GEMM (unav_set_phase_trace): where a tile's life goes.
    UNAV_TC_PPAIR=1 python scripts/gemm_phases_pp.py <G> <M> <N> <K> [act] [out: both|f32|op]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from unav_yolyolva_b200 import kernels as K, _cabi

os.environ["UNAV_TC_PPAIR"] = "1"
G, M, N, Kd = (int(x) for x in sys.argv[1:5])
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
outs = sys.argv[6] if len(sys.argv) > 6 else "both"
dev = torch.device("cuda", 0)
op = K.BF16X2
groups = []
for g in range(G):
    A = K.new_operand(M, Kd, op, dev); A.normal_()
    W = K.new_operand(N, Kd, op, dev); W.normal_()
    d = {"A": A, "W": W, "bias": torch.zeros(N, device=dev)}
    if outs in ("both", "f32"):
        d["out_f32"] = torch.empty(M, N, device=dev)
    if outs in ("both", "op"):
        d["out_op"] = K.new_operand(M, N, op, dev)
    groups.append(d)
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
t = buf.cpu().view(-1, 32)
t = t[t[:, 1] != 0]
print(f"{G}x[{M},{N},{Kd}] act={act} outs={outs} ctas={t.shape[0]} kernel {a.elapsed_time(b) * 1e3:.1f} us (eager, includes launch)")
st = lambda x: f"mean {x.double().mean():8.0f} min {int(x.min()):7d} max {int(x.max()):7d}"
print("  setup (alloc, barriers, cluster sync)      ", st(t[:, 2] - t[:, 1]))
print("  CTA total                                   ", st(t[:, 3] - t[:, 1]))
even = t[0::2] if t.shape[0] > 1 else t      # rank 0 of each pair carries the MMA stamps
for i in range(3):
    o = 8 * (i + 1)
    r = t[t[:, o + 4] != 0]
    if r.shape[0] == 0:
        break
    e = even[even[:, o + 3] != 0]
    print(f"  tile {i}: ({r.shape[0]} CTAs)")
    print("    start -> producer's first load of the tile ", st(r[:, o + 0] - r[:, 1]))
    if e.shape[0]:
        print("    start -> accumulator free (MMA)            ", st(e[:, o + 1] - e[:, 1]))
        print("    accumulator free -> first operands landed  ", st(e[:, o + 2] - e[:, o + 1]))
        print("    k-loop issue (first landed -> last issued) ", st(e[:, o + 3] - e[:, o + 2]))
    print("    start -> accumulator ready (epilogue)      ", st(r[:, o + 4] - r[:, 1]))
    print("    first 64-column pass                       ", st(r[:, o + 5] - r[:, o + 4]))
    print("    all four passes                            ", st(r[:, o + 6] - r[:, o + 4]))
    print("    release (bar.sync + arrive)                ", st(r[:, o + 7] - r[:, o + 6]))

if os.environ.get("UNAV_PP_FINE"):
    for o, what in ((24, "first pass of the first tile"), (28, "second pass of the second tile")):
        f = t[t[:, o] != 0]
        if f.shape[0]:
            print(f"  fine trace, {what} (warp 2):")
            print("    bias loads + tcgen05.ld + wait             ", st(f[:, o + 1] - f[:, o]))
            print("    staging stores + syncwarp                  ", st(f[:, o + 2] - f[:, o + 1]))
            print("    transposed store (if any) + row loop       ", st(f[:, o + 3] - f[:, o + 2]))
