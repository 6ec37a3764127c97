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

3. ``ship_reference()`` copies the reference's Python tree (``libs/``, ``configs/``, ``eval.py``) into
   ``oracle/_ref/reference/`` — git-ignored like the ``.so`` next to it, NOT gpurun-ignored — so that the
   UNMODIFIED reference travels to the GPU box with the snapshot: ``bench.py --impl reference`` times the
   reference's own ``PtTransformer`` there and a ``-m gpu`` test runs its ``eval.py``.  Nothing under
   ``oracle/_ref`` is ever committed; the recipe is this function (called from ``__graft_entry__.build()``).

``reference_root()`` is ``/root/reference`` where it exists (this container) and the shipped copy otherwise
(the GPU box).
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
REF_SHIP = os.path.join(REF_OUT, "reference")
STUBS = os.path.join(ORACLE_DIR, "stubs")
_NMS_SRC = os.path.join(REF_ROOT, "libs", "utils", "csrc", "nms_cpu.cpp")


def have_reference() -> bool:
    """The reference SOURCE tree is on this machine (the build container)."""
    return os.path.isfile(_NMS_SRC)


def reference_root():
    """Directory holding the reference's ``libs`` package: the source tree here, the shipped copy on the GPU box, or None."""
    if have_reference():
        return REF_ROOT
    if os.path.isfile(os.path.join(REF_SHIP, "libs", "modeling", "multimodal_meta_archs.py")):
        return REF_SHIP
    return None


def ship_reference() -> str:
    """Copy the reference's Python tree to oracle/_ref/reference (build container only; idempotent)."""
    import shutil
    if not have_reference():
        raise FileNotFoundError("reference sources not present")
    if os.path.isdir(REF_SHIP):
        shutil.rmtree(REF_SHIP)
    os.makedirs(REF_SHIP)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "build", "*.so", "*.o", "*.egg-info")
    for d in ("libs", "configs"):
        shutil.copytree(os.path.join(REF_ROOT, d), os.path.join(REF_SHIP, d), ignore=ignore)
    shutil.copy2(os.path.join(REF_ROOT, "eval.py"), os.path.join(REF_SHIP, "eval.py"))
    return REF_SHIP


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


def reference_pythonpath():
    """sys.path / PYTHONPATH entries that make the unmodified reference importable (also in worker processes: the stubs
    must be real on-disk packages): stubs, the reference root, and oracle/_ref for ``nms_1d_cpu``."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("neither /root/reference nor oracle/_ref/reference is available on this machine")
    return [STUBS, root, REF_OUT]


def import_reference():
    """Return the reference's ``libs`` package (the source tree in the build container, the shipped copy on the GPU box)."""
    paths = reference_pythonpath()
    if have_reference():
        build_ref_nms()
    assert load_ref_nms() is not None, "oracle/_ref/nms_1d_cpu.so is missing (run __graft_entry__.build() in the build container)"
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("libs")


if __name__ == "__main__":
    print(build_ref_nms(verbose=True))
