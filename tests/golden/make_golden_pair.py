#!/usr/bin/env python
"""Generates tests/golden/ref_pair.json and ref_eam_analytic.json from the REFERENCE's own, unmodified headers (zbl/potential.h,
exp6.h, buckingham.h, yukawa.h, relax/potential.h, zero/potential.h; sutton_chen.h, vniitf.h, johnson.h compiled where they lie
under /root/reference into oracle/_ref/libxsref.so).  Build container only:

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

J = 1.0 / 1.602176634e-19 * EV          # one joule in internal units
from oracle import oracle as O  # noqa: E402

CASES = {
    # zbl: {r1, rc, z_a, z_b}; decks: potentials/snap/monomat_zbl.msp (Ta, r1 0.1 rc 4.615858), multi_WBe.msp (r1 4.0 rc 4.8)
    "zbl_Ta": (1, [0.1, 4.615858, 73, 73]), "zbl_W_Be": (1, [4.0, 4.8, 74, 4]), "zbl_Be_Be": (1, [4.0, 4.8, 4, 4]),
    "exp6": (2, [3.0e5 * EV, 3.6, 60.0 * EV, 1.0e-6 * EV]), "buckingham": (3, [1.2e3 * EV, 0.32, 25.0 * EV]),
    # yukawa {A, kappa}; relax {r1, rc}: the overlap-relaxation ramp, clamped below r1 and above rc; zero {}
    # yukawa as in the reference's regression deck potentials/pair/yukawa/single_specy_nosym.msp:6 (A 2.43 eV*ang, kappa 4.1 1/ang)
    "yukawa": (4, [2.43 * EV, 4.1]), "yukawa_soft": (4, [25.0 * EV, 1.3]), "relax": (5, [0.8, 3.9]), "zero": (6, []),
}
# single-species analytic EAM models (eam_potential_template): model id, parameters in the reference struct's order
EAM_CASES = {
    # sutton_chen {c, epsilon, a0, n, m}: Cu of the reference's regression deck potentials/eam/eam_sutton_chen/single_specy.msp:6-13
    "sutton_chen_Cu": (1, [3.317e1, 3.605e-21 * J, 3.27, 9.05, 5.005]),
    # vniitf {rmax, rmin, rt0, Ecoh, E0, beta, A, Z, n, alpha, D, eta, mu}: Sn of potentials/eam/eam_vniitf/single_specy.msp:6-21
    "vniitf_Sn": (2, [5.599, 1.0, 3.437, 2.956031e-19 * J, 5.15003855e-20 * J, 6.0, 1.401, 7.618, 0.724, 3.072, 0.145, 2.72, -1.87]),
    # a second set with a narrow switching window (rmin close to rmax): both branches of the S3 spline inside the sampled range
    "vniitf_narrow_switch": (2, [5.5, 4.9, 3.44, 3.1 * EV, 0.02 * EV, 5.1, 1.05, 10.0, 0.62, 3.7, 0.08, 6.0, 2.5]),
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
        for r in np.concatenate([np.linspace(0.6, 5.2, 40), [prm[0], prm[1]] if pot in (1, 5) else []]):
            R.xsref_pair(pot, p, float(r), C.byref(e), C.byref(de))
            rows.append([float(r).hex(), e.value.hex(), de.value.hex()])
        out["cases"][name] = {"pot": pot, "params": [float(v).hex() for v in p], "rows": rows}
    with open(os.path.join(HERE, "ref_pair.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote ref_pair.json:", {k: len(v["rows"]) for k, v in out["cases"].items()})
    oute = {"generator": "tests/golden/make_golden_pair.py", "source": "oracle/_ref/libxsref.so (reference headers, unmodified)", "cases": {}}
    f_, df_ = C.c_double(), C.c_double()
    for name, (model, prm) in EAM_CASES.items():
        p = np.array(prm, dtype=np.float64)
        rows = []
        for what, xs in ((0, np.linspace(1.9, 6.2, 36)), (1, np.linspace(1.9, 6.2, 36)), (2, np.concatenate([[0.0], np.geomspace(1e-3, 60.0, 30)]))):
            for x in xs:
                R.xsref_eam_analytic(model, p, what, float(x), C.byref(f_), C.byref(df_))
                rows.append([what, float(x).hex(), f_.value.hex(), df_.value.hex()])
        oute["cases"][name] = {"model": model, "params": [float(v).hex() for v in p], "rows": rows}
    with open(os.path.join(HERE, "ref_eam_analytic.json"), "w") as f:
        json.dump(oute, f, indent=0)
    print("wrote ref_eam_analytic.json:", {k: len(v["rows"]) for k, v in oute["cases"].items()})


if __name__ == "__main__":
    main()
