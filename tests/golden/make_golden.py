#!/usr/bin/env python
"""Generates tests/golden/ref_math.json from the REFERENCE's own, unmodified math headers
(oracle/_ref/libxsref.so = lennard_jones.h, johnson.h, eam_alloy.h/.cpp compiled where they lie under
/root/reference, see oracle/Makefile).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every float is stored as a C99 hex literal so the comparison in tests/test_oracle_math.py is bit-exact.
The eam/alloy vectors are taken on setfl files written by tests/helpers.write_setfl (deterministic), so the
checker can rebuild the very same file anywhere; the reference's own data files are additionally compared
live in test_oracle_math.py when /root/reference is present.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ctypes as C  # noqa: E402
from helpers import EV, JOHNSON_CU, SC_CU, SC_XX, johnson_params, write_setfl  # noqa: E402
from oracle import oracle as O  # noqa: E402

# the Ta Johnson set of data/regression_new/potentials/eam/eam_johnson/single_specy.msp:6-27, with the unit
# suffixes the deck gives (eV, eV/ang -> internal energy units; note the deck also tags eta with eV)
JOHNSON_TA_DECK = dict(re=2.860082, fe=3.08634 * EV, rhoe=33.787168 * EV, alpha=8.489528, beta=4.527748, A=0.611679 * EV, B=1.032101 * EV,
                       kappa=0.176977, **{"lambda": 0.353954}, Fn0=-5.103845 * EV, Fn1=-0.405524 * EV, Fn2=1.112997 * EV, Fn3=-3.585325 * EV,
                       F0=-5.14 * EV, F1=0.0, F2=1.640098 * EV, F3=0.221375 * EV, Fo=-5.141526 * EV, eta=0.848843 * EV)
J_ORDER = ["re", "fe", "rhoe", "alpha", "beta", "A", "B", "kappa", "lambda", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo", "eta"]


def hx(v):
    return float(v).hex()


def setfl_cases(tmp):
    return {
        "cu_1species": dict(path=write_setfl(os.path.join(tmp, "a.eam.alloy"), [SC_CU], nrho=500, drho=0.4, nr=600, rc=6.0),
                            args=dict(elements=["SC_CU"], nrho=500, drho=0.4, nr=600, rc=6.0)),
        "cu_xx_2species": dict(path=write_setfl(os.path.join(tmp, "b.eam.alloy"), [SC_CU, SC_XX], nrho=400, drho=0.5, nr=500, rc=6.5),
                               args=dict(elements=["SC_CU", "SC_XX"], nrho=400, drho=0.5, nr=500, rc=6.5)),
    }


def main():
    R = O.ref()
    if R is None:
        raise SystemExit("oracle/_ref/libxsref.so is not built (needs /root/reference): make -C oracle ref")
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libxsref.so (reference headers, unmodified)"}
    e, de = C.c_double(), C.c_double()

    lj = []
    for eps, sigma in ((0.0104 * EV, 3.4), (0.583 * EV, 2.27)):
        for r in np.linspace(0.8 * sigma, 2.6 * sigma, 48):
            R.xsref_lj(eps, sigma, float(r), C.byref(e), C.byref(de))
            lj.append([hx(eps), hx(sigma), hx(r), hx(e.value), hx(de.value)])
    out["lj"] = lj

    jo = {}
    for name, d in (("cu_zhou", None), ("ta_deck", JOHNSON_TA_DECK)):
        p = johnson_params() if d is None else np.array([d[k] for k in J_ORDER], dtype=np.float64)
        rows = []
        for what, xs in ((0, np.linspace(1.6, 6.5, 40)), (1, np.linspace(1.6, 6.5, 40)),
                         (2, np.concatenate([np.linspace(0.05, 1.4, 40) * p[2], [0.85 * p[2], 1.15 * p[2]]]))):
            for x in xs:
                R.xsref_johnson(p, what, float(x), C.byref(e), C.byref(de))
                rows.append([what, hx(x), hx(e.value), hx(de.value)])
        jo[name] = {"params19": [hx(v) for v in p], "rows": rows}
    out["johnson"] = jo

    tmp = tempfile.mkdtemp()
    ea = {}
    for name, case in setfl_cases(tmp).items():
        T = O.EamAlloy(case["path"], use_ref=True)
        nel = T.nelements
        rec = {"write_setfl": case["args"], "nelements": nel, "nr": T.nr, "nrho": T.nrho, "rdr": hx(T.rdr), "rdrho": hx(T.rdrho),
               "rc": hx(T.rc), "rhomax": hx(T.rhomax), "ev_internal": hx(R.xsref_ev_internal())}
        tabs = {}
        for which, tname in ((0, "frho"), (1, "rhor"), (2, "z2r")):
            t = T.table(which)
            rows = sorted(set([1, 2, 3, len(t) // 3, len(t) // 2, len(t) - 3, len(t) - 2, len(t) - 1]))
            tabs[tname] = {"shape": list(t.shape), "rows": {str(m): [hx(v) for v in t[m]] for m in rows},
                           "sum": hx(float(np.sum(t)))}
        rec["tables"] = tabs
        ev = []
        rs = np.concatenate([np.linspace(0.3, T.rc * 1.02, 60), [T.rc, T.rc - 1e-9]])
        for ti in range(nel):
            for tj in range(nel):
                for r in rs:
                    v, _ = T.eval(0, r, ti, tj)
                    f, phi = T.eval(2, r, ti, tj, fpi=-0.37 * EV, fpj=-0.41 * EV)
                    ev.append([ti, tj, hx(r), hx(v), hx(f), hx(phi)])
        rec["pair_eval"] = ev
        em = []
        for ti in range(nel):
            for rho in np.concatenate([np.linspace(0.0, T.rhomax * 1.1, 50), [T.rhomax, -0.5]]):
                phi, fp = T.eval(1, rho, ti)
                em.append([ti, hx(rho), hx(phi), hx(fp)])
        rec["embed_eval"] = em
        ea[name] = rec
    out["eam_alloy"] = ea

    with open(os.path.join(HERE, "ref_math.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(HERE, "ref_math.json"))


if __name__ == "__main__":
    main()
