"""Build libunav_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m unav_yolyolva_b200.csrc.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["lib.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "rowops.cu", "attention.cu", "attention_tcgen05.cu", "decode_nms.cu", "metrics.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "unav_b200.h")]
OUT = os.path.join(HERE, "libunav_b200.so")               # BF16 halves
OUT_F16 = os.path.join(HERE, "libunav_b200_f16.so")       # same sources, -DUNAV_HALF_F16 (FP16 halves)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale():
    if not (os.path.exists(OUT) and os.path.exists(OUT_F16)):
        return True
    t = min(os.path.getmtime(OUT), os.path.getmtime(OUT_F16))
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    variants = (("", [], OUT), ("_f16", ["-DUNAV_HALF_F16"], OUT_F16))
    objs = {tag: [] for tag, _, _ in variants}
    for tag, defs, _ in variants:
        for src in SOURCES:
            obj = os.path.join(HERE, "build", src.replace(".cu", f"{tag}.o"))
            objs[tag].append(obj)
            cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", os.path.join(HERE, src), "-o", obj]
            procs.append((src + tag, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see log above")
    for tag, _, out in variants:
        subprocess.check_call([nvcc, "-shared", "-o", out, *objs[tag], "-cudart", "static"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
