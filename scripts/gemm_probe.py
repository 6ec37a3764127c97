"""Time hot-path GEMM shapes through the C ABI with CUDA-graph replay (device time, no host launch overhead).
    python scripts/gemm_probe.py [shape-index] [x3|x1]        (UNAV_TC_STAGES=n overrides the ring depth)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K

SHAPES = [(2, 3584, 2048, 512, K.ACT_GELU), (1, 448, 256, 256, 0), (6, 3584, 512, 512, 0), (1, 7056, 1024, 3072, 0),
          (3, 3584, 256, 256, 0), (1, 16384, 1280, 224, 0), (2, 3584, 512, 2048, 0), (1, 448, 256, 512, 0),
          (1, 448, 256, 1024, 0), (1, 448, 512, 1536, 0), (1, 3584, 512, 1536, 0), (1, 896, 256, 256, 0), (1, 448, 256, 64, 0),
          (1, 128, 64, 64, 0), (1, 7200, 1536, 512, 0), (2, 7056, 512, 1536, 0), (1, 7168, 512, 1536, 0), (3, 7168, 512, 512, 0),
          (1, 7168, 512, 1024, 0), (2, 3584, 512, 512, 0), (1, 7200, 512, 512, 0), (3, 7168, 256, 256, 0), (3, 1792, 256, 256, 0),
          (1, 1792, 512, 1536, 0), (1, 896, 512, 1536, 0), (1, 3584, 256, 768, 0), (1, 7056, 200, 1536, 0)]
dev = torch.device("cuda", 0)
only = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "all" else None
op = K.BF16X2 if (len(sys.argv) < 3 or sys.argv[2] == "x3") else K.BF16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for si, (G, M, N, Kd, act) in enumerate(SHAPES):
    if only is not None and si != only:
        continue
    groups = []
    for g in range(G):
        A = K.new_operand(M, Kd, op, dev); A.normal_()
        W = K.new_operand(N, Kd, op, dev); W.normal_()
        groups.append({"A": A, "W": W, "bias": torch.zeros(N, device=dev), "out_f32": torch.empty(M, N, device=dev),
                       "out_op": K.new_operand(M, N, op, dev)})
    for _ in range(3):
        K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05)
    torch.cuda.synchronize()
    R = 20
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(R):
            K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr.replay(); torch.cuda.synchronize()
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    warm = a.elapsed_time(b) / R * 1e3
    # cold: single launch after an L2 flush
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1):
        K.gemm(groups, M, N, Kd, op, act, False, K.GEMM_TCGEN05)
    cold = []
    for _ in range(5):
        flush.zero_(); a.record(); g1.replay(); b.record(); torch.cuda.synchronize(); cold.append(a.elapsed_time(b) * 1e3)
    print(f"{G}x[{M},{N},{Kd}] act={act} warm {warm:7.1f} us  cold {min(cold):7.1f} us  {2.0*G*M*N*Kd/warm/1e6:7.1f} TF/s alg (warm)", flush=True)
