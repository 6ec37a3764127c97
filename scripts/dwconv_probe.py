"""Time the dwconv_ln launches of the batch-16 path: streaming kernel vs the tiled one (UNAV_DWCONV_TILED=1), CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K
dev = torch.device("cuda", 0)
op = K.BF16X2
# (groups, nseg, T, stride, C, n_pre, n_out, f32_out)
SHAPES = [(2, 16, 224, 1, 512, 2, 3, False), (1, 32, 224, 1, 256, 0, 3, False), (2, 16, 224, 1, 512, 0, 2, False), (1, 32, 224, 1, 512, 0, 1, False),
          (1, 32, 224, 2, 512, 0, 1, True), (1, 32, 112, 1, 256, 0, 3, False), (1, 32, 112, 2, 512, 0, 1, True), (1, 32, 56, 1, 256, 0, 3, False),
          (1, 32, 14, 1, 256, 0, 3, False), (1, 32, 7, 1, 256, 0, 3, False)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (G, nseg, T, stride, C, n_pre, n_out, f32) in SHAPES:
    To = T // stride
    groups = []
    for g in range(G):
        x = torch.randn(nseg * T, C, device=dev)
        outs = []
        for j in range(n_out):
            d = {"dw": torch.randn(3 * C, device=dev), "ln_w": torch.rand(C, device=dev) + 0.5, "ln_b": torch.randn(C, device=dev), "src": (j % 2 if n_pre else -1),
                 "out_op": K.new_operand(nseg * To, C, op, dev)}
            if f32:
                d["out_f32"] = torch.empty(nseg * To, C, device=dev)
            outs.append(d)
        groups.append({"x": x, "mask_out": torch.ones(nseg * To, dtype=torch.uint8, device=dev),
                       "pre": [(torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)) for _ in range(n_pre)], "outs": outs})
    line = f"{G}x[{nseg}x{T},{C}] s{stride} pre{n_pre} out{n_out}:"
    for name in ("tiled", "stream"):
        os.environ.pop("UNAV_DWCONV_TILED", None); os.environ.pop("UNAV_DWCONV_STREAM", None)
        os.environ["UNAV_DWCONV_TILED" if name == "tiled" else "UNAV_DWCONV_STREAM"] = "1"
        for _ in range(2):
            K.dwconv_ln(groups, nseg, T, stride, C, op)
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            K.dwconv_ln(groups, nseg, T, stride, C, op)
        cold = []
        for _ in range(5):
            flush.zero_(); a.record(); g1.replay(); b.record(); torch.cuda.synchronize(); cold.append(a.elapsed_time(b) * 1e3)
        nbytes = G * nseg * To * C * (4 * stride + (4 + (4 if f32 else 0)) * n_out)
        line += f"  {name} {min(cold):6.1f} us ({nbytes / min(cold) / 1e3:6.0f} GB/s)"
    print(line, flush=True)
