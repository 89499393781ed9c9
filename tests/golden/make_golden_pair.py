#!/usr/bin/env python
"""Generates tests/golden/ref_pair.json from the REFERENCE's own, unmodified pair-potential headers (zbl/potential.h,
exp6.h, buckingham.h compiled where they lie under /root/reference into oracle/_ref/libxsref.so).  Build container only:

    python tests/golden/make_golden_pair.py

Floats are stored as C99 hex literals.  These potentials call exp()/pow(): the checker allows 1e-14 relative (libm may pick
an FMA or non-FMA exp kernel depending on the host CPU), everything else about the restatement is exact."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import EV  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # zbl: {r1, rc, z_a, z_b}; decks: potentials/snap/monomat_zbl.msp (Ta, r1 0.1 rc 4.615858), multi_WBe.msp (r1 4.0 rc 4.8)
    "zbl_Ta": (1, [0.1, 4.615858, 73, 73]), "zbl_W_Be": (1, [4.0, 4.8, 74, 4]), "zbl_Be_Be": (1, [4.0, 4.8, 4, 4]),
    "exp6": (2, [3.0e5 * EV, 3.6, 60.0 * EV, 1.0e-6 * EV]), "buckingham": (3, [1.2e3 * EV, 0.32, 25.0 * EV]),
}


def main():
    R = O.ref()
    if R is None or not hasattr(R, "xsref_pair"):
        raise SystemExit("oracle/_ref/libxsref.so is not built with the pair headers: make -C oracle ref")
    out = {"generator": "tests/golden/make_golden_pair.py", "source": "oracle/_ref/libxsref.so (reference headers, unmodified)", "cases": {}}
    e, de = C.c_double(), C.c_double()
    for name, (pot, prm) in CASES.items():
        p = np.array(prm, dtype=np.float64)
        rows = []
        for r in np.concatenate([np.linspace(0.6, 5.2, 40), [prm[0], prm[1]] if pot == 1 else []]):
            R.xsref_pair(pot, p, float(r), C.byref(e), C.byref(de))
            rows.append([float(r).hex(), e.value.hex(), de.value.hex()])
        out["cases"][name] = {"pot": pot, "params": [float(v).hex() for v in p], "rows": rows}
    with open(os.path.join(HERE, "ref_pair.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote ref_pair.json:", {k: len(v["rows"]) for k, v in out["cases"].items()})


if __name__ == "__main__":
    main()
