"""Host-side logic that needs no GPU: the dataset / kernel selection of src/ImageDataset.hs:108-178 as mirrored in
image_dataset.py, and the command line of app/Main.hs:64-77 as mirrored in __main__.py."""
import numpy as np
import pytest

from ska_sdp_accelerate_gridding_b200 import __main__ as cli
from ska_sdp_accelerate_gridding_b200 import image_dataset as D


def test_convert_and_sort_is_numeric():
    vals, names = D.convertAndSort(["10", "9", "-3.5", "100.0", "0"])
    assert vals == [-3.5, 0.0, 9.0, 10.0, 100.0]
    assert names == ["-3.5", "0", "9", "10", "100.0"]


def test_find_closest_list_matches_the_reference_loop():
    ws = [0.0, 1.0, 2.0, 4.0, 8.0]
    # src/ImageDataset.hs:151-167: binary search with max = len - 1, ties go to the upper neighbour
    assert [D.findClosestList(ws, w) for w in (-5, 0.4, 0.5, 0.6, 2.9, 3.0, 3.1, 7.9, 100)] == [0, 0, 1, 1, 2, 3, 3, 4, 4]
    assert D.findClosestList([7.0], 3.0) == 0


def test_kernel_loaders_select_and_order():
    k = lambda v: np.full((2, 2, 3, 3), v, complex)
    wstore = {"/wkern/0.008/%s/kern" % w: k(float(w)) for w in ("10.0", "-2.5", "9.0", "0.0")}
    wstore["/wkern/0.01/5.0/kern"] = k(99.0)  # another theta: must be ignored
    kerns, wbins = D.getWKernels(wstore, 0.008)
    assert wbins.tolist() == [-2.5, 0.0, 9.0, 10.0]
    assert [kerns[i][0, 0, 0, 0].real for i in range(4)] == [-2.5, 0.0, 9.0, 10.0]
    astore = {}
    for ant in ("0", "1", "2", "10"):
        for t in ("100.0", "200.0"):
            for f in ("1.0e8", "2.0e8"):
                astore["/akern/0.008/%s/%s/%s/kern" % (ant, t, f)] = np.full((3, 3), int(ant) + float(t) / 1e3 + float(f) / 1e10, complex)
    ak = D.getAKernels(astore, 0.008, 190.0, 1.1e8)  # closest time 200, closest frequency 1e8
    assert ak.shape == (4, 3, 3)
    assert np.allclose([a[0, 0].real for a in ak], [n + 0.2 + 0.01 for n in (0, 1, 2, 10)])  # antennas in numeric order


def test_command_line_parser():
    d = cli.parser([])
    assert d == {"n": 1, "input": "data", "out": None, "old": False, "flags": []}  # defArgs, app/Main.hs:26
    d = cli.parser(["-gpu", "-n", "5000", "-i", "/tmp/x", "-o", "res.npz", "-old", "-ddump-cc", "-debug"])
    assert d["n"] == 5000 and d["input"] == "/tmp/x" and d["out"] == "res.npz" and d["old"] and d["flags"] == ["dump-cc"]
    assert cli.parser(["-all"])["n"] is None
    with pytest.raises(SystemExit):
        cli.parser(["--bogus"])
    with pytest.raises(SystemExit):
        cli.parser(["-n"])
