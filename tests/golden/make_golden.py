"""Regenerates tests/golden/smalltest_oracle.npz and random_cases.npz from the CPU oracle (oracle/).
The reference itself cannot be run here (no GHC / LLVM 7 / libhdf5), so these vectors are
ORACLE outputs on the reference's in-source inputs (test/SmallTest.hs:51-73) and on seeded random inputs;
they pin the CUDA path against regressions of the oracle as well as of the kernels."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402


def smalltest_inputs():
    j = json.load(open(os.path.join(HERE, "smalltest_inputs.json")))
    x = np.arange(15)
    wk = np.broadcast_to(x * (0.01 + 0.005j) + 0.1, (1, 1, 1, 15, 15)).copy()
    ak = np.full((3, 15, 15), 0.1 + 0j)
    uvw = np.array(j["uvw"])
    idx = np.array(j["index"], np.int64)
    vis = np.array([complex(*v) for v in j["vis"]])
    return wk, ak, uvw, idx, vis


def main():
    wk, ak, uvw, idx, vis = smalltest_inputs()
    g = orc.convgrid_aw(wk, ak, np.zeros((10, 10), complex), uvw[:, 0], uvw[:, 1], idx[:, 0], idx[:, 1], idx[:, 2], vis)
    x, xf, y, yf = orc.frac_coords(10, 10, 1, uvw[:, 0], uvw[:, 1])
    np.savez(os.path.join(HERE, "smalltest_oracle.npz"), grid=g, x=x, xf=xf, y=y, yf=yf)
    rng = np.random.default_rng(20261018)
    n, s, q, nw, cnt = 96, 7, 4, 3, 500
    gcf = rng.standard_normal((nw, q, q, s, s)) + 1j * rng.standard_normal((nw, q, q, s, s))
    u = rng.uniform(-0.55, 0.55, cnt)
    v = rng.uniform(-0.55, 0.55, cnt)
    wb = rng.integers(0, nw, cnt)
    vv = rng.standard_normal(cnt) + 1j * rng.standard_normal(cnt)
    grid = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vv, wbin=wb)
    deg = orc.convdegrid(gcf, grid, u, v, wbin=wb)
    xs = orc.frac_coords(n, n, q, u, v)
    np.savez(os.path.join(HERE, "random_cases.npz"), gcf=gcf, u=u, v=v, wbin=wb, vis=vv, grid=grid, degrid=deg,
             x=xs[0], xf=xs[1], y=xs[2], yf=xs[3], n=n)
    print("written")


if __name__ == "__main__":
    main()
