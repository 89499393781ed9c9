"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/liboracle.so) and of the
reference-math library (oracle/_ref/libxsref.so, built from the reference's own headers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module.  The product package exastamp_b200 never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class GridDesc(C.Structure):
    """orc_grid_t (same field order as xsb_grid_desc's leading part)."""
    _fields_ = [("dims", C.c_int32 * 3), ("ghost_layers", C.c_int32), ("cell_size", C.c_double),
                ("origin", C.c_double * 3), ("xform", C.c_double * 9), ("xform_is_identity", C.c_int32),
                ("pad_", C.c_int32)]


def build(force=False):
    """make liboracle.so (+ _ref/libxsref.so when /root/reference exists)."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    if not need:
        so_t = os.path.getmtime(os.path.join(HERE, "liboracle.so"))
        for f in ("xs_oracle.cpp", "snap_oracle.cpp", "orc_math.h"):  # noqa
            p = os.path.join(HERE, f)
            if os.path.exists(p) and os.path.getmtime(p) > so_t:
                need = True
    if need:
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/potential") and (force or not os.path.exists(os.path.join(HERE, "_ref", "libxsref.so"))
                                                             or not os.path.exists(os.path.join(HERE, "_ref", "libxsref_snap.so"))):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None
_SIGS = []      # entry points declared by lib()


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        gp = C.POINTER(GridDesc)
        L.orc_nbh_build.restype = C.c_void_p
        L.orc_nbh_build.argtypes = [gp, _u64p, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int]
        L.orc_nbh_free.argtypes = [C.c_void_p]
        L.orc_nbh_total_size.restype = C.c_uint64
        L.orc_nbh_total_size.argtypes = [C.c_void_p]
        L.orc_nbh_export.argtypes = [C.c_void_p, _u64p, C.c_void_p]
        L.orc_nbh_import.restype = C.c_void_p
        L.orc_nbh_import.argtypes = [C.c_uint64, _u64p, _u16p, C.c_int, C.c_int]
        L.orc_nbh_decode.restype = C.c_uint64
        L.orc_nbh_decode.argtypes = [gp, _u64p, C.c_void_p, _u32p, _u64p, C.c_void_p]
        L.orc_nbh_bruteforce_counts.argtypes = [gp, C.c_uint64, _dp, _dp, _dp, C.c_double, _u32p]
        vp = C.c_void_p
        L.orc_pair_force.argtypes = [gp, _u64p, _dp, _dp, _dp, vp, C.c_int, _dp, C.c_double, C.c_int, _dp, _dp, _dp, vp, vp]
        L.orc_pair_multi_force.argtypes = [gp, _u64p, _dp, _dp, _dp, _u8p, vp, C.c_int, C.c_int, _dp, C.c_double, C.c_int,
                                           _dp, _dp, _dp, vp, vp]
        L.orc_eam_johnson.argtypes = [gp, _u64p, _dp, _dp, _dp, vp, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, vp, _dp]
        L.orc_eam_analytic.argtypes = [gp, _u64p, _dp, _dp, _dp, vp, C.c_int, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, vp, _dp]
        L.orc_eam_alloy_load.restype = vp
        L.orc_eam_alloy_load.argtypes = [C.c_char_p]
        L.orc_eam_alloy_free.argtypes = [vp]
        ip, dpp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.orc_eam_alloy_info.argtypes = [vp, ip, ip, ip, dpp, dpp, dpp, dpp]
        L.orc_eam_alloy_table.restype = C.POINTER(C.c_double)
        L.orc_eam_alloy_table.argtypes = [vp, C.c_int]
        L.orc_eam_alloy_eval.restype = C.c_double
        L.orc_eam_alloy_eval.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, dpp]
        L.orc_eam_alloy.argtypes = [gp, _u64p, _dp, _dp, _dp, _u8p, vp, vp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, vp, _dp]
        L.orc_snap_create.restype = vp
        L.orc_snap_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, vp]
        L.orc_snap_free.argtypes = [vp]
        L.orc_snap_ncoeff.argtypes = [vp]
        L.orc_snap_idxb.argtypes = [vp, vp]
        L.orc_snap_sizes.argtypes = [vp, ip, ip, ip, ip]
        L.orc_snap_atom.argtypes = [vp, C.c_int, _dp, _dp, _dp, vp, C.c_int, vp, vp, vp, vp]
        L.orc_snap_force.argtypes = [gp, _u64p, _dp, _dp, _dp, vp, vp, vp, C.c_double, C.c_int, _dp, _dp, _dp, vp, vp]
        L.orc_num_threads.restype = C.c_int
        L.orc_lj_eval.argtypes = [C.c_double, C.c_double, C.c_double, dpp, dpp]
        L.orc_johnson_eval.argtypes = [_dp, C.c_int, C.c_double, dpp, dpp]
        L.orc_eam_analytic_eval.argtypes = [C.c_int, _dp, C.c_int, C.c_double, dpp, dpp]
        L.orc_ev_internal.restype = C.c_double
        L.orc_pair_eval.argtypes = [C.c_int, _dp, C.c_double, dpp, dpp]
        L.orc_pair_ecut.restype = C.c_double
        L.orc_pair_ecut.argtypes = [C.c_int, _dp, C.c_double]
        L.orc_set_num_threads.argtypes = [C.c_int]
        _SIGS[:] = [n for n in dir(L) if n.startswith("orc_")]
        _lib = L
    return _lib


def lib_timed():
    """bench.py's CPU arm only: switch this module to a build of the same sources with -O3 -march=native (made on the
    machine that runs the timing; falls back to the pinned -march=x86-64-v3 build when g++ is unavailable).
    Returns (library, compiler flags)."""
    global _lib
    flags = "-O3 -march=x86-64-v3 -fopenmp -ffp-contract=off (pinned build; native build unavailable)"
    base = lib()
    import tempfile
    d = os.path.join(tempfile.gettempdir(), "xsb200_oracle_native_%d" % os.getuid())
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "liboracle_native.so")      # outside the tree: -march=native code must not travel to another CPU
    try:
        subprocess.check_call(["make", "-s", "-C", HERE, "native", "NATIVE_OUT=" + so], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(so)
        for name in _SIGS:
            f = getattr(L, name); b = getattr(base, name)
            f.argtypes = b.argtypes; f.restype = b.restype
        _lib = L
        flags = "-O3 -march=native -fopenmp"
    except Exception:      # noqa: BLE001
        L = base
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_num_threads.restype = C.c_int
    return L, flags


def lib_pinned():
    """back to the pinned build (bit-exact against the golden vectors) after lib_timed()"""
    global _lib
    _lib = None
    return lib()


def ref():
    """reference-math library; None when it was never built (no /root/reference and no prebuilt .so)."""
    global _ref
    if _ref is None:
        p = os.path.join(HERE, "_ref", "libxsref.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        dpp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        R.xsref_lj.argtypes = [C.c_double, C.c_double, C.c_double, dpp, dpp]
        R.xsref_johnson.argtypes = [_dp, C.c_int, C.c_double, dpp, dpp]
        R.xsref_eam_alloy_load.restype = vp
        R.xsref_eam_alloy_load.argtypes = [C.c_char_p, C.c_int]
        R.xsref_eam_alloy_free.argtypes = [vp]
        R.xsref_eam_alloy_info.argtypes = [vp, ip, ip, ip, dpp, dpp, dpp, dpp]
        R.xsref_eam_alloy_table.restype = C.POINTER(C.c_double)
        R.xsref_eam_alloy_table.argtypes = [vp, C.c_int]
        R.xsref_eam_alloy_eval.restype = C.c_double
        R.xsref_eam_alloy_eval.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, dpp]
        R.xsref_ev_internal.restype = C.c_double
        if hasattr(R, "xsref_pair"):
            R.xsref_pair.argtypes = [C.c_int, _dp, C.c_double, dpp, dpp]
        if hasattr(R, "xsref_eam_analytic"):
            R.xsref_eam_analytic.argtypes = [C.c_int, _dp, C.c_int, C.c_double, dpp, dpp]
        _ref = R
    return _ref


def make_grid(dims, ghost_layers, cell_size, origin, xform=None):
    g = GridDesc()
    g.dims[:] = [int(d) for d in dims]
    g.ghost_layers = int(ghost_layers)
    g.cell_size = float(cell_size)
    g.origin[:] = [float(o) for o in origin]
    X = np.eye(3) if xform is None else np.asarray(xform, dtype=np.float64).reshape(3, 3)
    g.xform[:] = [float(v) for v in X.ravel()]
    g.xform_is_identity = int(np.array_equal(X, np.eye(3)))
    return g


class Neighbors:
    """chunk_neighbors list held by the oracle (reference uint16 stream per cell)."""

    def __init__(self, handle, grid, cell_off, chunk_size, has_offsets):
        self.h, self.grid, self.cell_off = handle, grid, cell_off
        self.chunk_size, self.has_offsets = chunk_size, has_offsets

    @classmethod
    def build(cls, grid, cell_off, rx, ry, rz, nbh_dist_lab, chunk_size=1, build_particle_offset=True):
        h = lib().orc_nbh_build(C.byref(grid), cell_off, rx, ry, rz, float(nbh_dist_lab), int(chunk_size), int(build_particle_offset))
        return cls(h, grid, cell_off, chunk_size, build_particle_offset)

    @classmethod
    def from_streams(cls, grid, cell_off, stream_off, data, chunk_size=1, has_offsets=True):
        h = lib().orc_nbh_import(len(stream_off) - 1, np.ascontiguousarray(stream_off, dtype=np.uint64),
                                 np.ascontiguousarray(data, dtype=np.uint16), int(chunk_size), int(has_offsets))
        return cls(h, grid, cell_off, chunk_size, has_offsets)

    def export(self):
        ncells = len(self.cell_off) - 1
        off = np.zeros(ncells + 1, dtype=np.uint64)
        lib().orc_nbh_export(self.h, off, None)
        data = np.zeros(int(off[-1]), dtype=np.uint16)
        lib().orc_nbh_export(self.h, off, data.ctypes.data_as(C.c_void_p))
        return off, data

    def decode(self):
        """(counts[N], idx_off[N+1], idx[total]) flat CSR in stream-traversal order."""
        n = int(self.cell_off[-1])
        counts = np.zeros(n, dtype=np.uint32)
        idx_off = np.zeros(n + 1, dtype=np.uint64)
        tot = lib().orc_nbh_decode(C.byref(self.grid), self.cell_off, self.h, counts, idx_off, None)
        idx = np.zeros(int(tot), dtype=np.uint32)
        lib().orc_nbh_decode(C.byref(self.grid), self.cell_off, self.h, counts, idx_off, idx.ctypes.data_as(C.c_void_p))
        return counts, idx_off, idx

    def __del__(self):
        try:
            if self.h:
                lib().orc_nbh_free(self.h)
                self.h = None
        except Exception:
            pass


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


POT_LJ, POT_ZBL, POT_EXP6, POT_BUCKINGHAM, POT_YUKAWA, POT_RELAX, POT_ZERO = 0, 1, 2, 3, 4, 5, 6
EAM_JOHNSON, EAM_SUTTON_CHEN, EAM_VNIITF = 0, 1, 2


def pair_eval(pot, params, r):
    """(e, de) of one pair evaluation of the restated potential"""
    e, de = C.c_double(), C.c_double()
    lib().orc_pair_eval(int(pot), np.ascontiguousarray(params, dtype=np.float64), float(r), C.byref(e), C.byref(de))
    return e.value, de.value


def pair_ecut(pot, params, rcut):
    return lib().orc_pair_ecut(int(pot), np.ascontiguousarray(params, dtype=np.float64), float(rcut))


def pair_force(grid, cell_off, rx, ry, rz, nbh, params, rcut, ghost, fx, fy, fz, ep=None, vir=None, pot=0):
    lib().orc_pair_force(C.byref(grid), cell_off, rx, ry, rz, nbh.h, pot, np.ascontiguousarray(params, dtype=np.float64),
                         float(rcut), int(ghost), fx, fy, fz, _opt(ep), _opt(vir))


def pair_multi_force(grid, cell_off, rx, ry, rz, typ, nbh, pair_params, rcut_max, ghost, fx, fy, fz, ep=None, vir=None, pot=0):
    pp = np.ascontiguousarray(pair_params, dtype=np.float64)
    lib().orc_pair_multi_force(C.byref(grid), cell_off, rx, ry, rz, typ, nbh.h, pot, pp.shape[0], pp, float(rcut_max), int(ghost),
                               fx, fy, fz, _opt(ep), _opt(vir))


def eam_johnson(grid, cell_off, rx, ry, rz, nbh, params19, rcut, flags, fx, fy, fz, ep, vir, rho_dEmb):
    lib().orc_eam_johnson(C.byref(grid), cell_off, rx, ry, rz, nbh.h, np.ascontiguousarray(params19, dtype=np.float64), float(rcut),
                          int(flags), fx, fy, fz, ep, _opt(vir), rho_dEmb)


def eam_analytic(grid, cell_off, rx, ry, rz, nbh, model, params, rcut, flags, fx, fy, fz, ep, vir, rho_dEmb):
    """single-species analytic EAM (eam_potential_template): model 0 johnson, 1 sutton_chen, 2 vniitf; flags as eam_johnson"""
    lib().orc_eam_analytic(C.byref(grid), cell_off, rx, ry, rz, nbh.h, int(model), np.ascontiguousarray(params, dtype=np.float64), float(rcut),
                           int(flags), fx, fy, fz, ep, _opt(vir), rho_dEmb)


def eam_analytic_eval(model, params, what, x):
    """(f, df) of phi (what 0), rho (1) or fEmbed (2) of the restated model"""
    f, df = C.c_double(), C.c_double()
    lib().orc_eam_analytic_eval(int(model), np.ascontiguousarray(params, dtype=np.float64), int(what), float(x), C.byref(f), C.byref(df))
    return f.value, df.value


class EamAlloy:
    def __init__(self, path, use_ref=False):
        self.use_ref = use_ref
        self.L = ref() if use_ref else lib()
        self.pre = "xsref_" if use_ref else "orc_"
        load = getattr(self.L, self.pre + "eam_alloy_load")
        self.h = load(path.encode(), 0) if use_ref else load(path.encode())
        if not self.h:
            raise IOError("cannot read setfl file %s" % path)
        ne, nr, nrho = C.c_int(), C.c_int(), C.c_int()
        rdr, rdrho, rc, rhomax = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        getattr(self.L, self.pre + "eam_alloy_info")(self.h, ne, nr, nrho, rdr, rdrho, rc, rhomax)
        self.nelements, self.nr, self.nrho = ne.value, nr.value, nrho.value
        self.rdr, self.rdrho, self.rc, self.rhomax = rdr.value, rdrho.value, rc.value, rhomax.value

    def table(self, which):
        n = {0: self.nelements * (self.nrho + 1), 1: self.nelements * (self.nr + 1),
             2: self.nelements * (self.nelements + 1) // 2 * (self.nr + 1)}[which]
        p = getattr(self.L, self.pre + "eam_alloy_table")(self.h, which)
        return np.ctypeslib.as_array(p, shape=(n, 8)).copy()

    def eval(self, what, x, ti=0, tj=0, fpi=0.0, fpj=0.0):
        o2 = C.c_double()
        v = getattr(self.L, self.pre + "eam_alloy_eval")(self.h, what, float(x), ti, tj, float(fpi), float(fpj), o2)
        return v, o2.value


def eam_alloy(grid, cell_off, rx, ry, rz, typ, nbh, eam, rcut, flags, fx, fy, fz, ep, vir, rho_dEmb):
    assert not eam.use_ref
    lib().orc_eam_alloy(C.byref(grid), cell_off, rx, ry, rz, typ, nbh.h, eam.h, float(rcut), int(flags), fx, fy, fz, ep, _opt(vir), rho_dEmb)


class Snap:
    """SNAP parameters held by the oracle (LAMMPS SNA conventions; beta[nelements][ncoeff+1] in output energy units)."""

    def __init__(self, twojmax, rcutfac, radelem, wjelem, beta=None, rfac0=0.99363, rmin0=0.0, switchflag=1, bzeroflag=0):
        rad = np.ascontiguousarray(radelem, dtype=np.float64); wj = np.ascontiguousarray(wjelem, dtype=np.float64)
        self.nelements = len(rad)
        b = None if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
        self.h = lib().orc_snap_create(int(twojmax), float(rfac0), float(rmin0), int(switchflag), int(bzeroflag), self.nelements, rad, wj,
                                       float(rcutfac), None if b is None else b.ctypes.data_as(C.c_void_p))
        self.ncoeff = lib().orc_snap_ncoeff(self.h)
        self.twojmax, self.rcutfac, self.radelem, self.wjelem, self.beta = twojmax, rcutfac, rad, wj, b
        self.rfac0, self.rmin0, self.switchflag, self.bzeroflag = rfac0, rmin0, switchflag, bzeroflag
        if b is not None:
            assert b.shape == (self.nelements, self.ncoeff + 1)

    def idxb(self):
        t = np.zeros((self.ncoeff, 3), dtype=np.int32)
        lib().orc_snap_idxb(self.h, t.ctypes.data_as(C.c_void_p))
        return t

    def rcut_max(self):
        return 2.0 * float(self.radelem.max()) * self.rcutfac

    def atom(self, dx, dy, dz, elem_j=None, elem_i=0, want_db=False, want_force=False):
        """(B[ncoeff], dB[n][ncoeff][3] or None, energy, dE/dr_j [n][3] or None) for one neighbourhood"""
        dx, dy, dz = [np.ascontiguousarray(v, dtype=np.float64) for v in (dx, dy, dz)]
        n = len(dx)
        ej = None if elem_j is None else np.ascontiguousarray(elem_j, dtype=np.int32)
        B = np.zeros(self.ncoeff); dB = np.zeros((n, self.ncoeff, 3)) if want_db else None
        e = C.c_double(); dedr = np.zeros((n, 3)) if want_force else None
        lib().orc_snap_atom(self.h, n, dx, dy, dz, _opt(ej), int(elem_i), B.ctypes.data_as(C.c_void_p), _opt(dB), C.cast(C.byref(e), C.c_void_p), _opt(dedr))
        return B, dB, e.value, dedr

    def __del__(self):
        try:
            if self.h:
                lib().orc_snap_free(self.h); self.h = None
        except Exception:
            pass


def snap_force(grid, cell_off, rx, ry, rz, typ, nbh, snap, flags, fx, fy, fz, ep=None, vir=None):
    """flags: bit0 ghost, bit1 energy, bit2 virial"""
    lib().orc_snap_force(C.byref(grid), cell_off, rx, ry, rz, _opt(typ), nbh.h, snap.h, snap.rcut_max(), int(flags), fx, fy, fz, _opt(ep), _opt(vir))


def ref_snap():
    """the reference's in-tree SnapLegacyBS (LAMMPS mode); None when oracle/_ref/libxsref_snap.so was never built"""
    p = os.path.join(HERE, "_ref", "libxsref_snap.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    R.xsref_snap_nidx.argtypes = [C.c_double]
    R.xsref_snap_bs.argtypes = [C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, _dp, C.c_void_p, C.POINTER(C.c_double)]
    return R
