#!/usr/bin/env python
"""Generates tests/golden/snap_legacy.json: bispectrum components and their derivatives computed by the REFERENCE's
own in-tree implementation SnapLegacyBS/CG/GSH (src/potential/snaplegacy/lib, compiled unmodified with -DLAMMPS into
oracle/_ref/libxsref_snap.so, see oracle/Makefile) on seeded BCC-like neighbourhoods.  Run in the build container only:

    python tests/golden/make_golden_snap.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import oracle as O  # noqa: E402


def neighbourhood(seed, a=3.3, reach=5.2, sigma=0.08):
    rng = np.random.default_rng(seed)
    pts = []
    for i in range(-2, 3):
        for j in range(-2, 3):
            for k in range(-2, 3):
                for b in ((0, 0, 0), (.5, .5, .5)):
                    p = np.array([i + b[0], j + b[1], k + b[2]]) * a
                    if 1e-6 < np.linalg.norm(p) < reach:
                        pts.append(p)
    return np.array(pts) + rng.normal(0, sigma, (len(pts), 3))


def main():
    R = O.ref_snap()
    if R is None:
        raise SystemExit("oracle/_ref/libxsref_snap.so is not built (needs /root/reference): make -C oracle ref")
    out = {"generator": "tests/golden/make_golden_snap.py",
           "source": "oracle/_ref/libxsref_snap.so = reference SnapLegacyBS/CG/GSH -DLAMMPS (rfac0 0.99363, rmin0 0, PI 3.14159265359)", "cases": []}
    for seed, twoj, rcut in ((1, 4, 4.7), (2, 6, 4.7), (3, 7, 5.0), (4, 8, 4.7), (5, 8, 4.2)):
        pts = neighbourhood(seed)
        rx, ry, rz = [np.ascontiguousarray(pts[:, k]) for k in range(3)]
        n = len(pts); nidx = R.xsref_snap_nidx(twoj / 2)
        bs = np.zeros(nidx); dbs = np.zeros((n, nidx, 3)); im = C.c_double()
        rc = R.xsref_snap_bs(twoj / 2, rcut, n, rx, ry, rz, bs, dbs.ctypes.data_as(C.c_void_p), C.byref(im))
        assert rc == 0
        keep = [0, n // 3, n - 1]     # derivative rows of three neighbours are enough to pin the recursion
        out["cases"].append({"twojmax": twoj, "rcut": rcut, "pos": [[float(v).hex() for v in p] for p in pts],
                             "bs": [float(v).hex() for v in bs],
                             "dbs_rows": {str(i): [[float(v).hex() for v in row] for row in dbs[i]] for i in keep}})
    with open(os.path.join(HERE, "snap_legacy.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "snap_legacy.json"))


if __name__ == "__main__":
    main()
