"""CPU checks of the drop-in boundary: libskagrid.so loads, exports every symbol include/skagrid.h declares,
and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "skagrid.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skagrid_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ska_sdp_accelerate_gridding_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/skagrid.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.EXPORTS) == names


def test_header_compiles_as_c():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "skagrid.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_mirror_header_compiles():
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp_smalltest.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ska_sdp_accelerate_gridding_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.skagrid_create(0, C.byref(h))
    assert rc == -4 and not h.value  # SKAGRID_ENODEV
    assert b"no CPU fallback" in lib.skagrid_last_error(None)
    from ska_sdp_accelerate_gridding_b200 import gridding
    with pytest.raises(_lib.SkagridError):
        gridding.frac_coord(16, 4, [0.1])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or link it."""
    pkg = os.path.join(ROOT, "ska_sdp_accelerate_gridding_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle\.py|oracle/)")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f"{f} references the oracle"
