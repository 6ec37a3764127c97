"""A/B of GEMM tile policies on the hot path's shapes, one process, CUDA-graph replay (device time):
    python scripts/gemm_ab.py [x3|x1]      columns: default policy vs UNAV_TC_PPAIR=1 (persistent CTA pairs), warm / cold (L2 flushed)
Also checks that both give the same bits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import _cabi
from unav_yolyolva_b200 import kernels as K

SHAPES = [(2, 3600, 2048, 512, K.ACT_GELU), (2, 3600, 512, 2048, 0), (1, 7200, 1536, 512, 0), (1, 7200, 512, 512, 0),
          (6, 3584, 512, 512, 0), (2, 3584, 512, 512, 0), (2, 3584, 2048, 512, K.ACT_GELU), (2, 3584, 512, 2048, 0),
          (2, 3584, 512, 1536, 0), (1, 3584, 512, 2048, 0), (1, 16384, 1280, 224, 0), (1, 7168, 512, 1024, 0),
          (3, 7168, 256, 256, 0), (1, 7168, 256, 256, 0), (1, 7168, 256, 768, 0), (1, 7168, 512, 1536, 0), (3, 7168, 512, 512, 0),
          (1, 7168, 512, 512, 0), (1, 3584, 512, 1024, 0), (3, 3584, 256, 256, 0), (1, 3584, 256, 256, 0), (1, 3584, 256, 768, 0),
          (1, 3584, 512, 1536, 0), (1, 1792, 512, 1024, 0), (3, 1792, 256, 256, 0), (1, 1792, 512, 1536, 0),
          (1, 7056, 1024, 3072, 0), (2, 7056, 512, 1536, 0)]
dev = torch.device("cuda", 0)
op = K.BF16X2 if (len(sys.argv) < 2 or sys.argv[1] == "x3") else K.BF16
passes = int(os.environ.get("AB_PASSES", "0"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# default = the library's policy; ppair = persistent CTA pairs forced (16 epilogue warps); pp8 column = the same with 16 epilogue warps (UNAV_PP_EW=16)
VARIANTS = (("default", {}), ("ppair", {"UNAV_TC_PPAIR": "1"}), ("pp8", {"UNAV_TC_PPAIR": "1", "UNAV_PP_EW": "16"}))
tot = {"default": 0.0, "ppair": 0.0, "pp8": 0.0, "best": 0.0}
for (G, M, N, Kd, act) in SHAPES:
    groups = []
    for g in range(G):
        A = K.new_operand(M, Kd, op, dev); A.normal_()
        W = K.new_operand(N, Kd, op, dev); W.normal_(); W.mul_(Kd ** -0.5)
        groups.append({"A": A, "W": W, "bias": torch.zeros(N, device=dev), "out_f32": torch.empty(M, N, device=dev),
                       "out_op": K.new_operand(M, N, op, dev)})
    res, outs = {}, {}
    for name, env in VARIANTS:
        for k in ("UNAV_TC_PPAIR", "UNAV_PP_EW"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(2):
            K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05, passes=passes)
        var = K.GEMM_KERNELS[_cabi.load(op).unav_gemm_last_variant()]
        torch.cuda.synchronize()
        outs[name] = [g["out_f32"].clone() for g in groups]
        R = 10
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(R):
                K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05, passes=passes)
        gr.replay(); torch.cuda.synchronize()
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        warm = a.elapsed_time(b) / R * 1e3
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05, passes=passes)
        cold = []
        for _ in range(5):
            flush.zero_(); a.record(); g1.replay(); b.record(); torch.cuda.synchronize(); cold.append(a.elapsed_time(b) * 1e3)
        res[name] = (var, warm, min(cold))
    same = all(torch.equal(x, y) for x, y in zip(outs["default"], outs["ppair"])) and all(torch.equal(x, y) for x, y in zip(outs["default"], outs["pp8"]))
    fl = 2.0 * G * M * N * Kd
    d, p_ = res["default"], res["ppair"]
    tot["default"] += d[2]; tot["ppair"] += p_[2]; tot["pp8"] += res["pp8"][2]; tot["best"] += min(d[2], p_[2])
    print(f"{G}x[{M},{N},{Kd}] act={act} | {d[0][13:]:18s} warm {d[1]:6.1f} cold {d[2]:6.1f} us {fl/d[2]/1e6:6.1f} TF | "
          f"{p_[0][13:]:18s} warm {p_[1]:6.1f} cold {p_[2]:6.1f} us {fl/p_[2]/1e6:6.1f} TF | pp8 cold {res['pp8'][2]:6.1f} | x{d[2]/p_[2]:.2f} same_bits={same}", flush=True)
print("sum of cold times (us):", {k: round(v, 1) for k, v in tot.items()})
