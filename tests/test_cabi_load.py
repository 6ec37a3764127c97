"""The C-ABI library builds, loads without a GPU and exports every symbol include/unav_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "unav_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(unav_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from unav_yolyolva_b200.csrc import build
    so = build.build()
    assert os.path.exists(so)
    names = _declared_symbols()
    assert len(names) >= 17
    for path in (so, build.OUT_F16):            # BF16-halves and FP16-halves builds of the same sources
        assert os.path.exists(path)
        lib = ctypes.CDLL(path)
        for n in names:
            assert hasattr(lib, n), f"{n} declared in unav_b200.h but not exported by {os.path.basename(path)}"


def test_ctypes_prototypes_cover_header():
    from unav_yolyolva_b200 import _cabi
    assert set(_declared_symbols()) == set(_cabi.EXPORTS)
    lib = _cabi.load()
    assert b"sm_100a" in lib.unav_version() and b"BF16" in lib.unav_version()
    lib16 = _cabi.load(_cabi.F16X2)
    assert lib16 is not lib and b"FP16" in lib16.unav_version()


def test_no_device_is_reported_not_hidden():
    import torch
    from unav_yolyolva_b200 import _cabi
    if torch.cuda.is_available():
        return
    rc = _cabi.load().unav_check_device(0)
    assert rc != 0
    assert _cabi.load().unav_last_error()


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """cuobjdump evidence that the GEMM is tcgen05/TMA (UTCHMMA / UTMALDG / LDTM), not mma.sync (HMMA)."""
    import shutil
    import subprocess
    from unav_yolyolva_b200 import _cabi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        return
    sass = subprocess.run([exe, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert " HMMA" not in sass
