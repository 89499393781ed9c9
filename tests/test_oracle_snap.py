"""Pins the SNAP oracle (oracle/snap_oracle.cpp, LAMMPS-SNA restatement) against the reference's own in-tree
bispectrum implementation (SnapLegacyBS/CG/GSH): committed golden vectors, the live library when present, the
reference's stored test vectors bs.ref2 when /root/reference is present, plus internal consistency (finite differences,
Newton's third law, Euler identity used by the CUDA kernel for the energy).  No GPU needed."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "snap_legacy.json")))
fh = float.fromhex
# SnapLegacy uses PI = 3.14159265359 (SnapLegacyGSH.cpp:38, SnapLegacyBS.cpp:171): agreement is limited to ~1e-11
TOL = 2e-11


@pytest.mark.parametrize("k", range(len(GOLD["cases"])))
def test_bispectrum_and_derivatives_match_reference_snaplegacy_golden(k):
    c = GOLD["cases"][k]
    pos = np.array([[fh(v) for v in p] for p in c["pos"]])
    r = np.linalg.norm(pos, axis=1)
    inr = np.nonzero(r <= c["rcut"])[0]
    S = O.Snap(c["twojmax"], c["rcut"], [0.5], [1.0])
    B, dB, _, _ = S.atom(pos[inr, 0], pos[inr, 1], pos[inr, 2], want_db=True)
    bs = np.array([fh(v) for v in c["bs"]])
    assert len(bs) == S.ncoeff
    assert np.abs(B - bs).max() <= TOL * np.abs(bs).max()
    for i, rows in c["dbs_rows"].items():
        i = int(i)
        ref = np.array([[fh(v) for v in row] for row in rows])
        if r[i] > c["rcut"]:
            assert np.all(ref == 0.0)
            continue
        mine = dB[list(inr).index(i)]
        # SnapLegacyBS stores -dB/dr_j (SnapLegacyBS.cpp:256)
        assert np.abs(mine + ref).max() <= 20 * TOL * max(1.0, np.abs(ref).max())


def test_live_reference_snaplegacy_when_present():
    R = O.ref_snap()
    if R is None:
        pytest.skip("oracle/_ref/libxsref_snap.so not built")
    import ctypes as C
    rng = np.random.default_rng(11)
    pos = rng.uniform(-4.5, 4.5, (60, 3)); pos = pos[np.linalg.norm(pos, axis=1) > 1.8]
    rx, ry, rz = [np.ascontiguousarray(pos[:, k]) for k in range(3)]
    twoj, rcut = 6, 4.4
    nidx = R.xsref_snap_nidx(twoj / 2); bs = np.zeros(nidx); im = C.c_double()
    assert R.xsref_snap_bs(twoj / 2, rcut, len(pos), rx, ry, rz, bs, None, C.byref(im)) == 0
    inr = np.linalg.norm(pos, axis=1) <= rcut
    B, _, _, _ = O.Snap(twoj, rcut, [0.5], [1.0]).atom(rx[inr], ry[inr], rz[inr])
    assert np.abs(B - bs).max() <= TOL * np.abs(bs).max()


def test_reference_stored_vectors_bs_ref2_when_present():
    """tests/snap-compute-bs/{atom_positions.ref,bs.ref2}: the Fortran-mode component set (j1 = j2) of the reference's
    passing unit test, mapped onto the LAMMPS triples through B(j1,j2,j)/(j+1) symmetry; theta0 = r (pi-0.02)/rcut."""
    d = "/root/reference/src/potential/snaplegacy/tests/snap-compute-bs"
    if not os.path.exists(os.path.join(d, "bs.ref2")):
        pytest.skip("reference tree not present")
    pos = np.loadtxt(os.path.join(d, "atom_positions.ref"))[:, :3]
    ref = np.loadtxt(os.path.join(d, "bs.ref2"))
    PI = 3.14159265359
    twoj, rcut = 7, 5.0
    r = np.linalg.norm(pos, axis=1)
    inr = (r <= rcut) & (r >= 1e-12)
    S = O.Snap(twoj, rcut, [0.5], [1.0], rfac0=(PI - 0.020) / PI)
    B, _, _, _ = S.atom(pos[inr, 0], pos[inr, 1], pos[inr, 2])
    idx = {tuple(t): k for k, t in enumerate(S.idxb())}
    comp = []
    for j1 in range(twoj + 1):                       # SnapLegacyBS::n_idx_bs, non-LAMMPS branch
        for j in range(0, min(twoj, 2 * j1) + 1):
            if (j + 2 * j1) % 2 == 1:
                continue
            comp.append(B[idx[(j1, j1, j)]] if j >= j1 else B[idx[(j1, j, j1)]] * (j + 1.0) / (j1 + 1.0))
    comp = np.array(comp)
    assert len(comp) == len(ref) == 26
    assert np.abs((comp - ref) / ref).max() < 1e-7     # the reference test's own criterion (snap-compute-bs.cpp:104)


def _case(seed=3, twoj=8, nel=1):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-4.2, 4.2, (70, 3)); pos = pos[(np.linalg.norm(pos, axis=1) > 1.9) & (np.linalg.norm(pos, axis=1) < 4.6)]
    S0 = O.Snap(twoj, 4.7, [0.5] * nel, [1.0] * nel)
    beta = rng.normal(0, 1, (nel, S0.ncoeff + 1))
    rad = [0.5, 0.46][:nel]; wj = [1.0, 0.7][:nel]
    return pos, O.Snap(twoj, 4.7, rad, wj, beta), rng


@pytest.mark.parametrize("twoj,nel", [(8, 1), (6, 2), (3, 1)])
def test_force_is_energy_gradient_and_euler_identity(twoj, nel):
    pos, S, rng = _case(twoj=twoj, nel=nel)
    ej = rng.integers(0, nel, len(pos)).astype(np.int32)
    ei = nel - 1
    rc = (S.radelem[ei] + S.radelem[ej]) * S.rcutfac
    keep = np.linalg.norm(pos, axis=1) < rc
    pos, ej = pos[keep], ej[keep]
    B, _, e, dedr = S.atom(pos[:, 0], pos[:, 1], pos[:, 2], ej, ei, want_force=True)
    assert abs(e - (S.beta[ei, 0] + S.beta[ei, 1:] @ B)) < 1e-12 * max(1.0, abs(e))
    h = 1e-6
    for i in (0, len(pos) // 2, len(pos) - 1):
        for k in range(3):
            p = pos.copy(); p[i, k] += h; ep = S.atom(p[:, 0], p[:, 1], p[:, 2], ej, ei)[2]
            p[i, k] -= 2 * h; em = S.atom(p[:, 0], p[:, 1], p[:, 2], ej, ei)[2]
            assert abs((ep - em) / (2 * h) - dedr[i, k]) < 2e-8 * np.abs(dedr).max()


def test_snap_force_grid_newton3_and_translation():
    from helpers import GridSystem, lattice
    pos, typ, box = lattice("BCC", 5, 3.3, 0.06, seed=5)
    gs = GridSystem(pos, typ, box, box[0] / 3, 1)
    g = gs.oracle_grid()
    rng = np.random.default_rng(2)
    S0 = O.Snap(4, 4.7, [0.5], [1.0])
    S = O.Snap(4, 4.7, [0.5], [1.0], rng.normal(0, 1, (1, S0.ncoeff + 1)))
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 5.2, 1, True)
    fx, fy, fz, ep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, fx, fy, fz, ep)
    # fold ghost contributions back onto their owners (update_force_energy_from_ghost): the total force vanishes
    F = np.zeros((len(pos), 3))
    np.add.at(F, gs.src_index, np.stack([fx, fy, fz], axis=1))
    assert np.abs(F.sum(axis=0)).max() < 1e-9 * np.abs(F).max()
    assert np.all(ep[~gs.is_ghost] != 0.0)
