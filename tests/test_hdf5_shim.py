"""The HDF5 side of the drop-in: hdf5/skagrid_hdf5.cc must export exactly the C symbols the reference's
Haskell layer binds (src/Hdf5.hs:30-67) -- plus createDatasetLLong, which hdf5/hdf5.cc:109 exports without a binding -- with the
reference's signatures.  libhdf5 is absent here, so the file is compiled in its guarded form; what is checked is the ABI."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "hdf5", "skagrid_hdf5.cc")

# foreign import ccall unsafe "<symbol>", src/Hdf5.hs:30-67, in order of appearance
BOUND_BY_HDF5_HS = [
    "createh5File", "getRankDataset", "getDimsDataset", "readDatasetInt", "readDatasetLLong", "readDatasetDouble", "readDatasetComplex",
    "readDatasetsComplex", "readDatasetsDouble", "createDatasetInt", "createDatasetDouble", "createDatasetComplex", "listGroupMembers",
]


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hdf5") / "libskagrid_hdf5.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-DSKAGRID_HDF5_STUB", SRC, "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_exports_match_the_haskell_bindings(shim):
    r = subprocess.run(["nm", "-D", "--defined-only", shim], capture_output=True, text=True)
    assert r.returncode == 0
    exported = sorted(line.split()[-1] for line in r.stdout.splitlines() if " T " in line and not line.split()[-1].startswith("_"))
    assert len(BOUND_BY_HDF5_HS) == 13
    assert exported == sorted(BOUND_BY_HDF5_HS + ["createDatasetLLong"])


def test_guarded_build_is_loud_and_safe(shim, capfd):
    lib = C.CDLL(shim)
    lib.getRankDataset.restype = C.c_int
    assert lib.getRankDataset(b"nofile", b"/vis/vis") == -1
    lib.listGroupMembers.restype = C.POINTER(C.c_char_p)
    names = lib.listGroupMembers(b"nofile", b"/wkern")
    assert names[0] is None          # an empty, NULL-terminated list the caller can walk (and free)
    buf = (C.c_double * 4)(1.0, 2.0, 3.0, 4.0)
    lib.readDatasetDouble(b"nofile", b"/vis/uvw", buf)
    assert list(buf) == [1.0, 2.0, 3.0, 4.0]   # outputs untouched on failure
    err = capfd.readouterr().err
    assert "built without libhdf5" in err


def test_source_compiles_against_the_real_api_when_present():
    """Where <hdf5.h> exists the real implementation must compile; here it is absent and the test says so."""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-"], input="#include <hdf5.h>\n", capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("libhdf5 headers are not installed in this environment")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
