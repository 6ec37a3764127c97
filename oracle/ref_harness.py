"""Reference harness — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Two jobs, both optional and both confined to this container / to ``oracle/_ref``:

1. ``build_ref_nms()`` compiles the reference's one native file,
   ``/root/reference/libs/utils/csrc/nms_cpu.cpp`` (pybind11 module ``nms_1d_cpu``,
   built by ``/root/reference/libs/utils/setup.py:7-19``), *from where it lies* into
   ``oracle/_ref/`` (git-ignored, but it travels to the GPU box with the snapshot).
   No reference source is copied into the repo.
2. ``import_reference()`` makes the reference's Python package importable
   (``libs.modeling`` / ``libs.utils``) by putting three import stubs
   (``oracle/stubs``: mmengine / matplotlib / seaborn) and
   ``/root/reference`` on ``sys.path``.  Only usable where ``/root/reference`` exists,
   i.e. in the build container — golden fixtures are generated with it
   (``tests/golden/make_golden.py``) and committed.

On the GPU box ``/root/reference`` does not exist: only ``load_ref_nms()`` (the prebuilt
``oracle/_ref/nms_1d_cpu*.so``) is available there.
"""
from __future__ import annotations

import glob
import importlib
import importlib.util
import os
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
REF_OUT = os.path.join(ORACLE_DIR, "_ref")
STUBS = os.path.join(ORACLE_DIR, "stubs")
_NMS_SRC = os.path.join(REF_ROOT, "libs", "utils", "csrc", "nms_cpu.cpp")


def have_reference() -> bool:
    return os.path.isfile(_NMS_SRC)


def _find_ref_nms_so():
    hits = sorted(glob.glob(os.path.join(REF_OUT, "nms_1d_cpu*.so")))
    return hits[0] if hits else None


def build_ref_nms(verbose: bool = False) -> str:
    """Compile the reference nms_cpu.cpp into oracle/_ref (idempotent). Returns .so path."""
    so = _find_ref_nms_so()
    if so is not None:
        return so
    if not have_reference():
        raise FileNotFoundError("reference sources not present; oracle/_ref must be prebuilt")
    os.makedirs(REF_OUT, exist_ok=True)
    from torch.utils.cpp_extension import load

    load(
        name="nms_1d_cpu",
        sources=[_NMS_SRC],
        extra_cflags=["-O2", "-fopenmp"],  # flags of /root/reference/libs/utils/setup.py:13
        build_directory=REF_OUT,
        verbose=verbose,
        is_python_module=True,
    )
    so = _find_ref_nms_so()
    assert so is not None, "reference nms build produced no .so"
    return so


def load_ref_nms():
    """Import the prebuilt reference extension as module ``nms_1d_cpu`` (or None)."""
    if "nms_1d_cpu" in sys.modules:
        return sys.modules["nms_1d_cpu"]
    so = _find_ref_nms_so()
    if so is None:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location("nms_1d_cpu", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["nms_1d_cpu"] = mod
    return mod


def import_reference():
    """Return the reference's ``libs`` package (build container only)."""
    if not have_reference():
        raise FileNotFoundError("/root/reference is not available on this machine")
    build_ref_nms()
    assert load_ref_nms() is not None
    for p in (STUBS, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("libs")


if __name__ == "__main__":
    print(build_ref_nms(verbose=True))
