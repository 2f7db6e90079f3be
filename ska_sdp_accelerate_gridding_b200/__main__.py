"""Command line of the reference's executable (app/Main.hs:17-78), on the B200 path:

    python -m ska_sdp_accelerate_gridding_b200 [-n N | -all] [-old] [-i DIR] [-o FILE] [-g|-gpu|...] [-debug] [-d<flag>]

Reads DIR/SKA1_Low_wkern2, DIR/SKA1_Low_akern3, DIR/SKA1_Low_quick (`.npz` stand-ins whose keys are the HDF5 dataset
paths of the reference, see image_dataset.py / INTEGRATION.md section 5), runs ImageDataset.aw_gridding on the first N
visibilities (default 1, as app/Main.hs:26; -all = every one) and prints the maximum of the image, like `putStrLn (show
fourier)`.  -o writes the image to FILE (`/img`).  The backend flags (-g, -gpu, -debug, ...) and Accelerate's -d<flag>
debug switches are accepted and ignored: there is one backend here, the GPU, and no CPU fallback."""
from __future__ import annotations

import os
import sys


def parser(argv):
    """app/Main.hs:64-77, same flags, same defaults, same error on anything else."""
    args = {"n": 1, "input": "data", "out": None, "old": False, "flags": []}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-debug", "-g", "-G", "-gpu", "-GPU", "-Gpu"):
            pass
        elif a == "-n" and i + 1 < len(argv):
            i += 1
            args["n"] = int(argv[i])
        elif a == "-all":
            args["n"] = None
        elif a == "-old":
            args["old"] = True
        elif a == "-i" and i + 1 < len(argv):
            i += 1
            args["input"] = argv[i]
        elif a == "-o" and i + 1 < len(argv):
            i += 1
            args["out"] = argv[i]
        elif a.startswith("-d") and len(a) > 2:
            args["flags"].append(a[2:])
        else:
            raise SystemExit("Error while parsing" + repr(argv[i:]))
        i += 1
    return args


def _find(directory, stem):
    for ext in (".npz", ".h5.npz"):
        p = os.path.join(directory, stem + ext)
        if os.path.exists(p):
            return p
    raise SystemExit("%s: no %s.npz (the reference's .h5 inputs need libhdf5, which this build does not have; "
                     "scripts/make_standin_dataset.py writes stand-ins with the same layout)" % (directory, stem))


def main(argv=None):
    a = parser(sys.argv[1:] if argv is None else argv)
    from . import image_dataset as D
    if a["out"] is not None and os.path.exists(a["out"]):
        os.remove(a["out"])  # `remover`, app/Main.hs:57-61
    mx = D.aw_gridding(_find(a["input"], "SKA1_Low_wkern2"), _find(a["input"], "SKA1_Low_akern3"), _find(a["input"], "SKA1_Low_quick"),
                       n=a["n"], outfile=a["out"], old=a["old"])
    print(repr(float(mx)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
