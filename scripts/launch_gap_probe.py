"""Per-launch cost of back-to-back kernels inside one CUDA graph (same stream, dependent):
    tiny rowcopy, LayerNorm, attention with 1 / 64 / 256 CTAs, small GEMM.  Separates CTA time from launch-to-launch overhead."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K

dev = torch.device("cuda", 0)
op = K.BF16X2
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def per_launch(fn, n=50):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


x = torch.randn(64, 256, device=dev)
dst = K.new_operand(64, 256, op, dev)
print("rowcopy 64 rows      :", round(per_launch(lambda: K.rowcopy([{"src": x, "dst": dst, "nseg": 1, "seg_len_in": 64, "seg_len_out": 64, "C": 256}], op)), 2), "us")
w = torch.ones(256, device=dev); bb = torch.zeros(256, device=dev)
print("layernorm 64 rows    :", round(per_launch(lambda: K.layernorm_rows([{"x": x, "w": w, "b": bb, "out_op": dst}], 64, 256, op)), 2), "us")
for nb, T, nh, hs in ((1, 56, 1, 64), (16, 56, 4, 64), (32, 56, 4, 64), (8, 224, 4, 64), (16, 224, 4, 64), (32, 224, 4, 64)):
    C = nh * hs
    qo = K.new_operand(nb * T, C, op, dev); qo.normal_()
    ko = K.new_operand(nb * T, C, op, dev); ko.normal_()
    vt = K.new_operand(nb * C, T, op, dev); vt.normal_()
    kmask = torch.ones(nb, T, dtype=torch.uint8, device=dev)
    out = K.new_operand(nb * T, C, op, dev)
    grp = [{"q": qo, "k": ko, "vt": vt, "kmask": kmask, "out": out}]
    ctas = nb * nh * ((T + 127) // 128)
    print(f"attention_tc nb={nb} T={T} nh={nh} ({ctas} CTAs):", round(per_launch(lambda: K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op)), 2), "us")
for (M, N, Kd) in ((128, 64, 256), (1792, 256, 256), (7168, 256, 256), (7168, 512, 512)):
    A = K.new_operand(M, Kd, op, dev); A.normal_()
    W = K.new_operand(N, Kd, op, dev); W.normal_()
    o32 = torch.empty(M, N, device=dev)
    g = [{"A": A, "W": W, "bias": torch.zeros(N, device=dev), "out_f32": o32}]
    print(f"gemm [{M},{N},{Kd}]:", round(per_launch(lambda: K.gemm(g, M, N, Kd, op, 0, False, K.GEMM_TCGEN05)), 2), "us")
