"""Test/bench helpers: synthetic lattices and a numpy statement of the cell grid with materialised periodic
ghost particles (what exaNBody's grid + ghost_update_* give the force operators)."""
import numpy as np


def lattice(structure, ncells, a, noise_sigma=0.0, seed=1, types=None):
    """perfect lattice + Gaussian noise (decks: lattice + gaussian_noise_r).  Returns pos[N,3] in [0,L), type[N], box[3]."""
    basis = {"FCC": [[0.25, 0.25, 0.25], [0.25, 0.75, 0.75], [0.75, 0.25, 0.75], [0.75, 0.75, 0.25]],
             "BCC": [[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]],
             "SC": [[0.5, 0.5, 0.5]]}[structure.upper()]
    nc = np.array([ncells] * 3 if np.isscalar(ncells) else ncells, dtype=np.int64)
    i, j, k = np.meshgrid(np.arange(nc[0]), np.arange(nc[1]), np.arange(nc[2]), indexing="ij")
    cells = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    pos = (cells[:, None, :] + np.asarray(basis)[None, :, :]).reshape(-1, 3) * a
    nb = len(basis)
    if types is None:
        typ = np.zeros(len(pos), dtype=np.uint8)
    else:
        typ = np.tile(np.asarray(types, dtype=np.uint8), len(pos) // nb)
    box = nc.astype(np.float64) * a
    if noise_sigma > 0:
        rng = np.random.default_rng(seed)
        pos = pos + rng.normal(0.0, noise_sigma, pos.shape)
        pos = np.mod(pos, box)
        pos = np.where(pos >= box, 0.0, pos)
    return np.ascontiguousarray(pos), typ, box


class GridSystem:
    """flat SoA sorted by cell (IJK, i fastest) including ghost cells filled with periodic images."""

    def __init__(self, pos, typ, box, cell_size, ghost_layers, xform=None):
        pos = np.asarray(pos, dtype=np.float64)
        box = np.asarray(box, dtype=np.float64)
        n_own = np.maximum(1, np.floor(box / cell_size + 1e-9).astype(np.int64))
        self.cell_size = float(cell_size)
        # cells must tile the box exactly for periodic wrap; callers pick cell_size = box / integer
        assert np.allclose(n_own * cell_size, box), "cell_size must divide the box"
        gl = int(ghost_layers)
        self.gl, self.n_own, self.box = gl, n_own, box
        self.dims = n_own + 2 * gl
        self.origin = -gl * self.cell_size * np.ones(3)
        self.xform = np.eye(3) if xform is None else np.asarray(xform, dtype=np.float64)
        nx, ny, nz = [int(d) for d in self.dims]
        ncells = nx * ny * nz
        ijk = np.clip(np.floor(pos / cell_size).astype(np.int64), 0, n_own - 1) + gl
        cid = ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2])
        order = np.argsort(cid, kind="stable")
        own_count = np.bincount(cid, minlength=ncells)
        own_start = np.concatenate([[0], np.cumsum(own_count)])
        # every cell (own or ghost) mirrors an own cell: wrap per axis
        ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        # flat order must be i fastest
        ci, cj, ck = [np.transpose(a, (2, 1, 0)).ravel() for a in (ci, cj, ck)]
        cc = np.stack([ci, cj, ck], axis=1)
        wrapn = np.floor_divide(cc - gl, n_own)
        src = (cc - gl) - wrapn * n_own + gl
        src_cid = src[:, 0] + nx * (src[:, 1] + ny * src[:, 2])
        counts = own_count[src_cid]
        self.cell_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        ntot = int(self.cell_off[-1])
        # for each output particle: index into the sorted-own array
        cell_of = np.repeat(np.arange(ncells), counts)
        within = np.arange(ntot) - np.repeat(self.cell_off[:-1].astype(np.int64), counts)
        src_sorted = own_start[src_cid[cell_of]] + within
        self.src_index = order[src_sorted]              # index into the caller's original arrays
        shift = (wrapn[cell_of] * box[None, :]).astype(np.float64)
        p = pos[self.src_index] + shift
        self.rx = np.ascontiguousarray(p[:, 0]); self.ry = np.ascontiguousarray(p[:, 1]); self.rz = np.ascontiguousarray(p[:, 2])
        self.type = np.ascontiguousarray(np.asarray(typ, dtype=np.uint8)[self.src_index])
        self.is_ghost = np.any(wrapn[cell_of] != 0, axis=1)
        self.ghost_cell = np.any(wrapn != 0, axis=1)
        self.n = ntot
        self.n_owned = int((~self.is_ghost).sum())
        self.ncells = ncells

    def oracle_grid(self):
        from oracle import oracle as O
        return O.make_grid(self.dims, self.gl, self.cell_size, self.origin, self.xform)

    def zeros(self):
        return np.zeros(self.n, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------
# synthetic potential parameter sets / files
# ---------------------------------------------------------------------------------------------------
EV = 1.602176634e-19 / (1.66053906660e-27 * 1.0e4)   # 1 eV in internal units (ang, Da, ps)

# Johnson-form EAM parameter set: the only one shipped by the reference is Ta
# (data/regression_new/potentials/eam/eam_johnson/single_specy.msp:6-27); a Cu set of the same functional
# form (Zhou/Johnson/Wadley 2004 Cu values, eV / ang) is used for the FCC Cu benchmark.  Energies in eV
# are converted to internal units by the caller (johnson_params()).
JOHNSON_CU = dict(re=2.556162, fe=1.554485, rhoe=21.175871, alpha=8.127620, beta=4.334731, A=0.396620, B=0.548085,
                  kappa=0.308782, **{"lambda": 0.756515}, Fn0=-2.170269, Fn1=-0.263788, Fn2=1.088878, Fn3=-0.817603,
                  F0=-2.19, F1=0.0, F2=0.561830, F3=-2.100595, Fo=-2.186568, eta=0.310490)
_J_ORDER = ["re", "fe", "rhoe", "alpha", "beta", "A", "B", "kappa", "lambda", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo", "eta"]
_J_ENERGY = {"A", "B", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo"}


def johnson_params(d=JOHNSON_CU):
    """19 scalars in the reference's order (johnson.h:29-50), energies converted eV -> internal."""
    return np.array([d[k] * (EV if k in _J_ENERGY else 1.0) for k in _J_ORDER], dtype=np.float64)


def write_setfl(path, elements, nrho=2000, drho=0.1, nr=2000, rc=7.29, rmin=0.6):
    """write a setfl (eam/alloy) file with Sutton-Chen-form tables, the same functional forms as the reference's
    generator scripts/python/pytab-eam-alloy/pytab-eam-alloy-sutton-chen.py (F=-c*eps*sqrt(rho), rho=(a/r)^m,
    r*phi = r*eps*(a/r)^n), evaluated with r clamped below at `rmin` so every knot is finite.
    elements: list of dict(name, z, mass, a0, c, eps, a, n, m).  Cross pairs use arithmetic/geometric mixing."""
    dr = rc / nr
    r = np.maximum(np.arange(nr) * dr, rmin)
    rho_x = np.arange(nrho) * drho
    with open(path, "w") as f:
        f.write("synthetic Sutton-Chen setfl generated by tests/helpers.py\nxsb200 test fixture\nXX\n")
        f.write("%d %s\n" % (len(elements), " ".join(e["name"] for e in elements)))
        f.write("%d %.12e %d %.12e %.12e\n" % (nrho, drho, nr, dr, rc))
        for e in elements:
            f.write("%d %.4f %.4f fcc\n" % (e["z"], e["mass"], e["a0"]))
            F = -e["c"] * e["eps"] * np.sqrt(rho_x)
            rho = (e["a"] / r) ** e["m"]
            f.write("\n".join("%.16e" % v for v in F) + "\n")
            f.write("\n".join("%.16e" % v for v in rho) + "\n")
        for i in range(len(elements)):
            for j in range(i + 1):
                ei, ej = elements[i], elements[j]
                eps = np.sqrt(ei["eps"] * ej["eps"]); a = 0.5 * (ei["a"] + ej["a"]); n = 0.5 * (ei["n"] + ej["n"])
                rphi = r * eps * (a / r) ** n
                f.write("\n".join("%.16e" % v for v in rphi) + "\n")
    return path


# Sutton-Chen Cu of the reference's generator script (c, eps[J]->eV, a, n, m : lines 91-96) and a second,
# made-up element so that multi-species paths can be exercised where /root/reference is absent.
SC_CU = dict(name="Cu", z=29, mass=63.546, a0=3.27, c=33.17, eps=3.605e-21 / 1.6021892e-19, a=3.27, n=9.050, m=5.005)
SC_XX = dict(name="Xx", z=13, mass=26.982, a0=3.50, c=30.00, eps=0.0200, a=3.40, n=8.5, m=5.5)


# ---------------------------------------------------------------------------------------------------
# parameter files the reference ships, committed as fixtures (tests/golden/make_potential_fixtures.py)
# ---------------------------------------------------------------------------------------------------
def potential_file(name):
    """path of an uncompressed copy of tests/golden/potentials/<name>.gz (checked against SHA256SUMS)"""
    import gzip
    import hashlib
    import os
    import tempfile
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "potentials")
    want = {l.split()[1]: l.split()[0] for l in open(os.path.join(gold, "SHA256SUMS")) if l.strip()}
    cache = os.path.join(tempfile.gettempdir(), "xsb200_potentials_%d" % os.getuid())
    os.makedirs(cache, exist_ok=True)
    out = os.path.join(cache, name)
    if not os.path.exists(out) or hashlib.sha256(open(out, "rb").read()).hexdigest() != want[name]:
        raw = gzip.open(os.path.join(gold, name + ".gz"), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == want[name], "fixture %s does not match its recorded checksum" % name
        tmp = out + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(raw)
        os.replace(tmp, out)
    return out


def read_snap_files(param_path, coeff_path):
    """LAMMPS .snapparam / .snapcoeff (reference reader: src/potential/snaplegacy/lib/snap_read_lammps.cpp:25-93; the C++
    twin used by the deck layer is host/xsbh_readers.cpp).  Coefficients stay in eV: callers scale by EV."""
    prm = {}
    for line in open(param_path):
        t = line.split("#")[0].split()
        if len(t) >= 2:
            prm[t[0]] = t[1]
    tok = []
    for line in open(coeff_path):
        tok += line.split("#")[0].split()
    nel, ncoef = int(tok[0]), int(tok[1])
    k = 2
    els = []
    for _ in range(nel):
        name, rad, wj = tok[k], float(tok[k + 1]), float(tok[k + 2]); k += 3
        beta = np.array([float(v) for v in tok[k:k + ncoef]]); k += ncoef
        els.append(dict(name=name, radius=rad, weight=wj, beta=beta))
    return dict(twojmax=int(prm["twojmax"]), rcutfac=float(prm["rcutfac"]), rfac0=float(prm.get("rfac0", 0.99363)),
                rmin0=float(prm.get("rmin0", 0.0)), bzeroflag=int(prm.get("bzeroflag", 1)), quadraticflag=int(prm.get("quadraticflag", 0)),
                switchflag=int(prm.get("switchflag", 1)), elements=els, ncoeff_with_beta0=ncoef)
