"""Build libunav_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m unav_yolyolva_b200.csrc.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["lib.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "rowops.cu", "attention.cu", "attention_tcgen05.cu", "decode_nms.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "unav_b200.h")]
OUT = os.path.join(HERE, "libunav_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
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
    subprocess.check_call([nvcc, "-shared", "-o", OUT, *objs, "-cudart", "static"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
