"""exastamp_b200 -- thin ctypes binding of libxsb200.so (the C ABI declared in include/xsb200.h).

The package holds no numerics of its own: every operator call goes through the C ABI into hand-written
sm_100a CUDA kernels.  There is no CPU fallback -- importing works without a GPU (so the build check and the
symbol tests can run), but creating a Context or calling any operator without the library / a B200 raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxsb200.so")

# xsb_field
F_RX, F_RY, F_RZ, F_FX, F_FY, F_FZ, F_EP, F_VX, F_VY, F_VZ, F_VIRIAL, F_RHO_DEMB, F_TYPE, F_ID = range(14)
FLAG_GHOST, FLAG_ENERGY, FLAG_VIRIAL, FLAG_MIXED = 1, 2, 4, 8
EAM_RHO, EAM_RHO2EMB, EAM_GHOST, EAM_FORCE, EAM_EFLAG = 1, 2, 4, 8, 16
POT_LJ, POT_ZBL, POT_EXP6, POT_BUCKINGHAM, POT_YUKAWA, POT_RELAX, POT_ZERO = 0, 1, 2, 3, 4, 5, 6
EAM_JOHNSON, EAM_SUTTON_CHEN, EAM_VNIITF = 0, 1, 2
_FIELD_DTYPE = {F_TYPE: np.uint8, F_ID: np.uint64}

# every symbol include/xsb200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "xsb_create", "xsb_destroy", "xsb_last_error", "xsb_sync", "xsb_version", "xsb_kernel_launch_count",
    "xsb_profile_enable", "xsb_profile_read", "xsb_timer_record", "xsb_timer_elapsed_ms", "xsb_measure_peaks",
    "xsb_grid_set", "xsb_grid_set_xform", "xsb_particles_set_cells", "xsb_num_particles", "xsb_num_cells", "xsb_field_upload",
    "xsb_field_download", "xsb_field_device_ptr", "xsb_zero_force_energy",
    "xsb_chunk_neighbors_build", "xsb_chunk_neighbors_stats", "xsb_chunk_neighbors_export_size",
    "xsb_chunk_neighbors_export", "xsb_chunk_neighbors_download_flat",
    "xsb_pair_force", "xsb_pair_multi_force", "xsb_eam_johnson_force", "xsb_eam_analytic_force",
    "xsb_snap_ncoeff", "xsb_snap_set", "xsb_snap_rcut_max", "xsb_snap_force", "xsb_snap_overflow",
    "xsb_eam_alloy_read", "xsb_eam_alloy_free", "xsb_eam_alloy_set", "xsb_eam_alloy_force",
    "xsb_particles_assign", "xsb_particles_rebin", "xsb_push_f_v_r", "xsb_push_f_v", "xsb_force_to_accel", "xsb_backup_r",
    "xsb_particle_displ_over", "xsb_verlet_boundary", "xsb_comm_unique_id", "xsb_comm_init", "xsb_comm_allreduce_max", "xsb_num_own_particles", "xsb_cell_offsets_download", "xsb_ghost_comm_scheme", "xsb_ghost_update", "xsb_ghost_reduce_add",
    "xsb_thermo_state", "xsb_ghost_plan", "xsb_migration_stats",
    "xsb_verlet_boundary_async", "xsb_displ_poll", "xsb_ghost_transport", "xsb_eam_inner_skin", "xsb_eam_sublist_stats", "xsb_chain_stats",
    "xsb_fields_upload_async", "xsb_fields_download_async", "xsb_copy_wait", "xsb_out_of_domain_count",
    "xsb_step_capture_begin", "xsb_step_capture_end", "xsb_step_replay", "xsb_step_release",
]


class XsbError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("ghost_layers", C.c_int32), ("cell_size", C.c_double),
                ("origin", C.c_double * 3), ("xform", C.c_double * 9), ("xform_is_identity", C.c_int32),
                ("pad_", C.c_int32)]


class ChunkNeighborsConfig(C.Structure):
    _fields_ = [("chunk_size", C.c_int32), ("build_particle_offset", C.c_int32), ("subcell_compaction", C.c_int32),
                ("free_scratch_memory", C.c_int32), ("stream_prealloc_factor", C.c_double)]


class EamAlloyTables(C.Structure):
    _fields_ = [("nelements", C.c_int32), ("nr", C.c_int32), ("nrho", C.c_int32), ("pad_", C.c_int32),
                ("rdr", C.c_double), ("rdrho", C.c_double), ("rc", C.c_double), ("rhomax", C.c_double),
                ("conversion_z2r", C.c_double), ("conversion_frho", C.c_double),
                ("frho", C.POINTER(C.c_double)), ("rhor", C.POINTER(C.c_double)), ("z2r", C.POINTER(C.c_double))]


class SnapParams(C.Structure):
    _fields_ = [("twojmax", C.c_int32), ("switchflag", C.c_int32), ("bzeroflag", C.c_int32), ("nelements", C.c_int32),
                ("quadraticflag", C.c_int32), ("chemflag", C.c_int32), ("switchinnerflag", C.c_int32), ("pad_", C.c_int32),
                ("rfac0", C.c_double), ("rmin0", C.c_double), ("rcutfac", C.c_double),
                ("radelem", C.POINTER(C.c_double)), ("wjelem", C.POINTER(C.c_double)), ("beta", C.POINTER(C.c_double))]


class DomainDesc(C.Structure):
    _fields_ = [("global_cells", C.c_int32 * 3), ("periodic", C.c_int32 * 3), ("rank_dims", C.c_int32 * 3),
                ("rank_coord", C.c_int32 * 3), ("box", C.c_double * 3)]


_lib = None


def build():
    from . import buildlib as _b
    return _b.build()


def load_library():
    """dlopen libxsb200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XsbError("libxsb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "exastamp_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, dbl = C.c_void_p, C.c_int, C.c_uint64, C.c_double
    L.xsb_create.argtypes = [i32, C.POINTER(vp)]
    L.xsb_destroy.argtypes = [vp]
    L.xsb_last_error.restype = C.c_char_p
    L.xsb_last_error.argtypes = [vp]
    L.xsb_version.restype = C.c_char_p
    L.xsb_sync.argtypes = [vp]
    L.xsb_kernel_launch_count.restype = u64
    L.xsb_kernel_launch_count.argtypes = [vp]
    L.xsb_timer_record.argtypes = [vp, i32]
    L.xsb_timer_elapsed_ms.argtypes = [vp, C.POINTER(dbl)]
    L.xsb_grid_set_xform.argtypes = [vp, C.POINTER(dbl)]
    L.xsb_measure_peaks.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]
    L.xsb_profile_enable.argtypes = [vp, i32]
    L.xsb_step_capture_begin.argtypes = [vp]; L.xsb_step_capture_end.argtypes = [vp, C.POINTER(C.c_int)]
    L.xsb_step_replay.argtypes = [vp, i32]; L.xsb_step_release.argtypes = [vp, i32]
    L.xsb_profile_read.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(u64)]
    L.xsb_grid_set.argtypes = [vp, C.POINTER(GridDesc)]
    L.xsb_particles_set_cells.argtypes = [vp, vp]
    L.xsb_num_particles.restype = u64
    L.xsb_num_particles.argtypes = [vp]
    L.xsb_num_cells.restype = u64
    L.xsb_num_cells.argtypes = [vp]
    L.xsb_field_upload.argtypes = [vp, i32, vp]
    L.xsb_field_download.argtypes = [vp, i32, vp]
    L.xsb_field_device_ptr.restype = vp
    L.xsb_field_device_ptr.argtypes = [vp, i32]
    L.xsb_zero_force_energy.argtypes = [vp, i32]
    L.xsb_chunk_neighbors_build.argtypes = [vp, dbl, C.POINTER(ChunkNeighborsConfig)]
    L.xsb_chunk_neighbors_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(C.c_uint32)]
    L.xsb_chunk_neighbors_export_size.argtypes = [vp, C.POINTER(u64)]
    L.xsb_chunk_neighbors_export.argtypes = [vp, vp, vp]
    L.xsb_chunk_neighbors_download_flat.argtypes = [vp, vp, vp, vp]
    L.xsb_pair_force.argtypes = [vp, i32, vp, i32, dbl, i32]
    L.xsb_pair_multi_force.argtypes = [vp, i32, i32, vp, i32, dbl, i32]
    L.xsb_eam_johnson_force.argtypes = [vp, vp, dbl, i32, i32]
    L.xsb_eam_analytic_force.argtypes = [vp, i32, vp, i32, dbl, i32, i32]
    L.xsb_eam_alloy_read.argtypes = [C.c_char_p, C.POINTER(EamAlloyTables), C.c_char_p, C.c_size_t]
    L.xsb_eam_alloy_free.argtypes = [C.POINTER(EamAlloyTables)]
    L.xsb_eam_alloy_set.argtypes = [vp, C.POINTER(EamAlloyTables)]
    L.xsb_eam_alloy_force.argtypes = [vp, dbl, i32, i32]
    L.xsb_snap_ncoeff.argtypes = [i32]
    L.xsb_snap_set.argtypes = [vp, C.POINTER(SnapParams)]
    L.xsb_snap_rcut_max.restype = dbl
    L.xsb_snap_rcut_max.argtypes = [vp]
    L.xsb_snap_force.argtypes = [vp, i32]
    L.xsb_snap_overflow.argtypes = [vp, C.POINTER(i32)]
    L.xsb_particles_assign.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.xsb_particles_rebin.argtypes = [vp, C.POINTER(DomainDesc)]
    L.xsb_push_f_v_r.argtypes = [vp, dbl]
    L.xsb_push_f_v.argtypes = [vp, dbl]
    L.xsb_force_to_accel.argtypes = [vp, i32, vp]
    L.xsb_backup_r.argtypes = [vp]
    L.xsb_particle_displ_over.argtypes = [vp, dbl, C.POINTER(i32), C.POINTER(dbl)]
    L.xsb_verlet_boundary.argtypes = [vp, i32, vp, dbl, dbl, C.POINTER(i32), C.POINTER(dbl)]
    L.xsb_thermo_state.argtypes = [vp, i32, vp, vp]
    L.xsb_ghost_plan.argtypes = [vp, i32, vp, vp, u64, vp]
    L.xsb_migration_stats.argtypes = [vp, vp, vp]
    L.xsb_comm_unique_id.argtypes = [vp]
    L.xsb_comm_init.argtypes = [vp, i32, i32, vp]
    L.xsb_ghost_comm_scheme.argtypes = [vp, C.POINTER(DomainDesc)]
    L.xsb_comm_allreduce_max.argtypes = [vp, C.POINTER(C.c_double)]
    L.xsb_num_own_particles.restype = u64
    L.xsb_num_own_particles.argtypes = [vp]
    L.xsb_cell_offsets_download.argtypes = [vp, vp]
    L.xsb_ghost_update.argtypes = [vp, C.c_uint32]
    L.xsb_eam_inner_skin.argtypes = [vp, dbl]
    L.xsb_eam_sublist_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.xsb_chain_stats.argtypes = [vp, C.POINTER(u64)]
    L.xsb_ghost_transport.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.xsb_verlet_boundary_async.argtypes = [vp, i32, vp, dbl]
    L.xsb_displ_poll.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.xsb_fields_upload_async.argtypes = [vp, i32, vp, vp, i32]
    L.xsb_fields_download_async.argtypes = [vp, i32, vp, vp, i32]
    L.xsb_copy_wait.argtypes = [vp]
    L.xsb_out_of_domain_count.argtypes = [vp, C.POINTER(u64)]
    L.xsb_ghost_reduce_add.argtypes = [vp, C.c_uint32]
    _lib = L
    return L


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


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """one xsb_ctx (one GPU).  Methods mirror the C ABI one to one; host numpy buffers in, host numpy buffers out."""

    def __init__(self, device=0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.xsb_create(int(device), C.byref(self.h))
        if rc != 0:
            msg = self.L.xsb_last_error(self.h).decode() if self.h else "xsb_create failed"
            if self.h:
                self.L.xsb_destroy(self.h)
                self.h = None
            raise XsbError("xsb_create(%d): %s" % (device, msg))

    def close(self):
        if getattr(self, "h", None):
            self.L.xsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise XsbError("%s failed (%d): %s" % (what, rc, self.L.xsb_last_error(self.h).decode()))

    # ---- a1
    def grid_set(self, grid):
        self.grid = grid
        self._ck(self.L.xsb_grid_set(self.h, C.byref(grid)), "xsb_grid_set")

    def grid_set_xform(self, xform):
        X = np.ascontiguousarray(xform, dtype=np.float64).reshape(9)
        self._ck(self.L.xsb_grid_set_xform(self.h, X.ctypes.data_as(C.POINTER(C.c_double))), "xsb_grid_set_xform")
        self.grid.xform[:] = [float(v) for v in X]
        self.grid.xform_is_identity = int(np.array_equal(X.reshape(3, 3), np.eye(3)))

    def particles_set_cells(self, cell_off):
        off = np.ascontiguousarray(cell_off, dtype=np.uint64)
        self._ck(self.L.xsb_particles_set_cells(self.h, _ptr(off)), "xsb_particles_set_cells")

    @property
    def n(self):
        return int(self.L.xsb_num_particles(self.h))

    def upload(self, field, arr):
        dt = _FIELD_DTYPE.get(field, np.float64)
        a = np.ascontiguousarray(arr, dtype=dt)
        want = self.n * (9 if field == F_VIRIAL else 1)
        if a.size != want:
            raise XsbError("upload(field %d): %d elements given, %d expected" % (field, a.size, want))
        self._ck(self.L.xsb_field_upload(self.h, field, _ptr(a)), "xsb_field_upload")
        self._ck(self.L.xsb_sync(self.h), "xsb_sync")   # `a` may be a temporary

    def download(self, field, out=None):
        dt = _FIELD_DTYPE.get(field, np.float64)
        shape = (self.n, 9) if field == F_VIRIAL else (self.n,)
        if out is None:
            out = np.empty(shape, dtype=dt)
        self._ck(self.L.xsb_field_download(self.h, field, _ptr(out)), "xsb_field_download")
        return out

    def device_ptr(self, field):
        return self.L.xsb_field_device_ptr(self.h, field)

    def zero_force_energy(self, ghost=False):
        self._ck(self.L.xsb_zero_force_energy(self.h, int(ghost)), "xsb_zero_force_energy")

    def sync(self):
        self._ck(self.L.xsb_sync(self.h), "xsb_sync")

    @property
    def launches(self):
        return int(self.L.xsb_kernel_launch_count(self.h))

    # ---- a2
    def chunk_neighbors(self, nbh_dist_lab, chunk_size=1, build_particle_offset=True, stream_prealloc_factor=1.05):
        cfg = ChunkNeighborsConfig(int(chunk_size), int(build_particle_offset), 1, 0, float(stream_prealloc_factor))
        self._ck(self.L.xsb_chunk_neighbors_build(self.h, float(nbh_dist_lab), C.byref(cfg)), "xsb_chunk_neighbors_build")

    def chunk_neighbors_stats(self):
        t, m = C.c_uint64(), C.c_uint32()
        self._ck(self.L.xsb_chunk_neighbors_stats(self.h, C.byref(t), C.byref(m)), "xsb_chunk_neighbors_stats")
        return t.value, m.value

    def chunk_neighbors_export(self):
        tot = C.c_uint64()
        self._ck(self.L.xsb_chunk_neighbors_export_size(self.h, C.byref(tot)), "xsb_chunk_neighbors_export_size")
        off = np.zeros(int(self.L.xsb_num_cells(self.h)) + 1, dtype=np.uint64)
        data = np.zeros(max(1, tot.value), dtype=np.uint16)
        self._ck(self.L.xsb_chunk_neighbors_export(self.h, _ptr(off), _ptr(data)), "xsb_chunk_neighbors_export")
        return off, data[:tot.value]

    def chunk_neighbors_flat(self):
        total, _ = self.chunk_neighbors_stats()
        counts = np.zeros(self.n, dtype=np.uint32)
        offs = np.zeros(self.n + 1, dtype=np.uint64)
        idx = np.zeros(max(1, total), dtype=np.uint32)
        self._ck(self.L.xsb_chunk_neighbors_download_flat(self.h, _ptr(counts), _ptr(offs), _ptr(idx)), "xsb_chunk_neighbors_download_flat")
        return counts, offs, idx[:total]

    # ---- a3-a6
    def pair_force(self, params, rcut, flags=FLAG_ENERGY, pot=POT_LJ):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.xsb_pair_force(self.h, pot, _ptr(p), p.size, float(rcut), int(flags)), "xsb_pair_force")

    def pair_multi_force(self, n_types, pair_params, rcut_max, flags=FLAG_ENERGY, pot=POT_LJ):
        p = np.ascontiguousarray(pair_params, dtype=np.float64)
        self._ck(self.L.xsb_pair_multi_force(self.h, pot, int(n_types), _ptr(p), p.shape[1] - 1, float(rcut_max), int(flags)), "xsb_pair_multi_force")

    # ---- a7-a8
    def eam_johnson_force(self, params19, rcut, phases=7, flags=0):
        p = np.ascontiguousarray(params19, dtype=np.float64)
        assert p.size == 19
        self._ck(self.L.xsb_eam_johnson_force(self.h, _ptr(p), float(rcut), int(phases), int(flags)), "xsb_eam_johnson_force")

    def eam_analytic_force(self, model, params, rcut, phases=7, flags=0):
        """single-species analytic EAM of eam_potential_template: EAM_JOHNSON (19 scalars), EAM_SUTTON_CHEN (5), EAM_VNIITF (13)"""
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.xsb_eam_analytic_force(self.h, int(model), _ptr(p), p.size, float(rcut), int(phases), int(flags)), "xsb_eam_analytic_force")

    def eam_alloy_load(self, path):
        t = EamAlloyTables()
        names = C.create_string_buffer(256)
        rc = self.L.xsb_eam_alloy_read(path.encode(), C.byref(t), names, 256)
        if rc != 0:
            raise XsbError("xsb_eam_alloy_read(%s) failed (%d)" % (path, rc))
        try:
            self._ck(self.L.xsb_eam_alloy_set(self.h, C.byref(t)), "xsb_eam_alloy_set")
            info = dict(nelements=t.nelements, nr=t.nr, nrho=t.nrho, rc=t.rc, names=names.value.decode().split())
        finally:
            self.L.xsb_eam_alloy_free(C.byref(t))
        return info

    def eam_inner_skin(self, skin):
        self._ck(self.L.xsb_eam_inner_skin(self.h, float(skin)), "xsb_eam_inner_skin")

    def eam_sublist_stats(self):
        """(rho phases that re-filtered the neighbour list, rho phases that only re-evaluated the sub-list)"""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.xsb_eam_sublist_stats(self.h, C.byref(a), C.byref(b)), "xsb_eam_sublist_stats")
        return a.value, b.value

    def chain_stats(self):
        """pair operators evaluated inside the force pass of the eam_alloy_force operator in front of them"""
        a = C.c_uint64()
        self._ck(self.L.xsb_chain_stats(self.h, C.byref(a)), "xsb_chain_stats")
        return a.value

    def eam_alloy_force(self, rcut, phases=EAM_RHO | EAM_RHO2EMB | EAM_GHOST | EAM_FORCE, flags=0):
        self._ck(self.L.xsb_eam_alloy_force(self.h, float(rcut), int(phases), int(flags)), "xsb_eam_alloy_force")

    # ---- a9
    def snap_set(self, twojmax, rcutfac, radelem, wjelem, beta, rfac0=0.99363, rmin0=0.0, switchflag=1, bzeroflag=0):
        rad = np.ascontiguousarray(radelem, dtype=np.float64); wj = np.ascontiguousarray(wjelem, dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        nc = self.L.xsb_snap_ncoeff(int(twojmax))
        if b.shape != (len(rad), nc + 1):
            raise XsbError("snap_set: beta must be [nelements][ncoeff+1] = [%d][%d]" % (len(rad), nc + 1))
        p = SnapParams(int(twojmax), int(switchflag), int(bzeroflag), len(rad), 0, 0, 0, 0, float(rfac0), float(rmin0), float(rcutfac),
                       rad.ctypes.data_as(C.POINTER(C.c_double)), wj.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)))
        self._ck(self.L.xsb_snap_set(self.h, C.byref(p)), "xsb_snap_set")
        return self.L.xsb_snap_rcut_max(self.h)

    def snap_force(self, flags=FLAG_ENERGY):
        self._ck(self.L.xsb_snap_force(self.h, int(flags)), "xsb_snap_force")

    def snap_overflow(self):
        f = C.c_int()
        self._ck(self.L.xsb_snap_overflow(self.h, C.byref(f)), "xsb_snap_overflow")
        return bool(f.value)

    # ---- recorded steps (one CUDA graph launch per step)
    def step_capture_begin(self):
        self._ck(self.L.xsb_step_capture_begin(self.h), "xsb_step_capture_begin")

    def step_capture_end(self):
        sid = C.c_int(-1)
        self._ck(self.L.xsb_step_capture_end(self.h, C.byref(sid)), "xsb_step_capture_end")
        return sid.value

    def step_replay(self, sid):
        self._ck(self.L.xsb_step_replay(self.h, int(sid)), "xsb_step_replay")

    def step_release(self, sid):
        self._ck(self.L.xsb_step_release(self.h, int(sid)), "xsb_step_release")

    # ---- profiling (CUDA events on the context's stream)
    PROF_TAGS = ["nbr_build", "pair", "eam_rho", "eam_rho2emb", "eam_force", "ghost", "integrate", "snap", "move"]

    def profile_enable(self, on=True):
        self._ck(self.L.xsb_profile_enable(self.h, int(on)), "xsb_profile_enable")

    def profile_read(self):
        out = {}
        for t, name in enumerate(self.PROF_TAGS):
            ms, cnt = C.c_double(), C.c_uint64()
            self._ck(self.L.xsb_profile_read(self.h, t, C.byref(ms), C.byref(cnt)), "xsb_profile_read")
            out[name] = (ms.value, cnt.value)
        return out

    def measure_peaks(self):
        """(FP64 TFLOP/s of a DFMA loop, FP32 TFLOP/s of an FFMA loop, HBM GB/s of a 1 GiB copy) measured on this device"""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.xsb_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(c)), "xsb_measure_peaks")
        return a.value, b.value, c.value

    def timer_start(self):
        self._ck(self.L.xsb_timer_record(self.h, 0), "xsb_timer_record")

    def timer_stop_ms(self):
        self._ck(self.L.xsb_timer_record(self.h, 1), "xsb_timer_record")
        ms = C.c_double()
        self._ck(self.L.xsb_timer_elapsed_ms(self.h, C.byref(ms)), "xsb_timer_elapsed_ms")
        return ms.value

    # raw-pointer variants for callers that manage their own (pinned) host buffers
    def upload_ptr(self, field, ptr):
        self._ck(self.L.xsb_field_upload(self.h, field, ptr), "xsb_field_upload")

    def download_ptr(self, field, ptr):
        self._ck(self.L.xsb_field_download(self.h, field, ptr), "xsb_field_download")

    # asynchronous own-atom / whole-array transfers of scalar double fields from / to pinned host memory (raw pointers)
    def fields_upload_async(self, fields, ptrs, own_only=True):
        f = (C.c_int * len(fields))(*[int(x) for x in fields]); p = (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])
        self._ck(self.L.xsb_fields_upload_async(self.h, len(fields), f, p, int(own_only)), "xsb_fields_upload_async")

    def fields_download_async(self, fields, ptrs, own_only=True):
        f = (C.c_int * len(fields))(*[int(x) for x in fields]); p = (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])
        self._ck(self.L.xsb_fields_download_async(self.h, len(fields), f, p, int(own_only)), "xsb_fields_download_async")

    def copy_wait(self):
        self._ck(self.L.xsb_copy_wait(self.h), "xsb_copy_wait")

    def out_of_domain_count(self):
        c = C.c_uint64()
        self._ck(self.L.xsb_out_of_domain_count(self.h, C.byref(c)), "xsb_out_of_domain_count")
        return c.value

    # ---- a10
    def comm_init(self, nranks, rank, unique_id=None):
        self._ck(self.L.xsb_comm_init(self.h, int(nranks), int(rank), unique_id), "xsb_comm_init")

    def set_domain(self, global_cells, periodic=(1, 1, 1), rank_dims=(1, 1, 1), rank_coord=(0, 0, 0), box=None):
        d = DomainDesc()
        d.global_cells[:] = [int(v) for v in global_cells]
        d.periodic[:] = [int(v) for v in periodic]
        d.rank_dims[:] = [int(v) for v in rank_dims]
        d.rank_coord[:] = [int(v) for v in rank_coord]
        if box is None:
            box = [g * self.grid.cell_size for g in global_cells]
        d.box[:] = [float(v) for v in box]
        self.domain = d
        return d

    def ghost_comm_scheme(self):
        self._ck(self.L.xsb_ghost_comm_scheme(self.h, C.byref(self.domain)), "xsb_ghost_comm_scheme")

    def particles_assign(self, rx, ry, rz, vx=None, vy=None, vz=None, typ=None, ids=None):
        f = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
        arrs = [f(rx, np.float64), f(ry, np.float64), f(rz, np.float64), f(vx, np.float64), f(vy, np.float64), f(vz, np.float64),
                f(typ, np.uint8), f(ids, np.uint64)]
        ptrs = [None if a is None else _ptr(a) for a in arrs]
        self._ck(self.L.xsb_particles_assign(self.h, len(arrs[0]), *ptrs), "xsb_particles_assign")

    def particles_rebin(self):
        self._ck(self.L.xsb_particles_rebin(self.h, C.byref(self.domain)), "xsb_particles_rebin")

    def migration_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.xsb_migration_stats(self.h, C.byref(a), C.byref(b)), "xsb_migration_stats")
        return a.value, b.value

    def push_f_v_r(self, dt):
        self._ck(self.L.xsb_push_f_v_r(self.h, float(dt)), "xsb_push_f_v_r")

    def push_f_v(self, dt):
        self._ck(self.L.xsb_push_f_v(self.h, float(dt)), "xsb_push_f_v")

    def force_to_accel(self, masses):
        m = np.ascontiguousarray(masses, dtype=np.float64)
        self._ck(self.L.xsb_force_to_accel(self.h, m.size, _ptr(m)), "xsb_force_to_accel")

    def backup_r(self):
        self._ck(self.L.xsb_backup_r(self.h), "xsb_backup_r")

    def particle_displ_over(self, threshold):
        r, d = C.c_int(), C.c_double()
        self._ck(self.L.xsb_particle_displ_over(self.h, float(threshold), C.byref(r), C.byref(d)), "xsb_particle_displ_over")
        return bool(r.value), d.value

    def verlet_boundary(self, masses, dt, threshold):
        """force_to_accel, push_f_v(dt/2) | push_f_v_r(dt), push_f_v(dt/2), particle_displ_over(threshold) in one pass"""
        m = np.ascontiguousarray(masses, dtype=np.float64)
        r, d = C.c_int(), C.c_double()
        self._ck(self.L.xsb_verlet_boundary(self.h, m.size, _ptr(m), float(dt), float(threshold), C.byref(r), C.byref(d)), "xsb_verlet_boundary")
        return bool(r.value), d.value

    def verlet_boundary_async(self, masses, dt):
        m = np.ascontiguousarray(masses, dtype=np.float64)
        self._ck(self.L.xsb_verlet_boundary_async(self.h, m.size, _ptr(m), float(dt)), "xsb_verlet_boundary_async")

    def displ_poll(self, lag=1):
        """(max displacement since backup_r, max displacement of that step) recorded `lag` verlet_boundary_async calls ago"""
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.xsb_displ_poll(self.h, int(lag), C.byref(a), C.byref(b)), "xsb_displ_poll")
        return a.value, b.value

    def thermo_state(self, masses):
        """simulation_thermodynamic_state: dict of the reference's 27 sums (own cells, all ranks)"""
        m = np.ascontiguousarray(masses, dtype=np.float64)
        out = np.zeros(27)
        self._ck(self.L.xsb_thermo_state(self.h, m.size, _ptr(m), _ptr(out)), "xsb_thermo_state")
        return dict(virial=out[0:9].reshape(3, 3), ke_tensor=out[9:18].reshape(3, 3), momentum=out[18:21], kinetic_energy=out[21:24],
                    potential_energy=out[24], mass=out[25], particle_count=int(out[26]))

    def ghost_transport(self):
        buf = C.create_string_buffer(256)
        self._ck(self.L.xsb_ghost_transport(self.h, buf, 256), "xsb_ghost_transport")
        return buf.value.decode()

    def ghost_update(self, fields):
        self._ck(self.L.xsb_ghost_update(self.h, sum(1 << f for f in fields)), "xsb_ghost_update")

    def ghost_reduce_add(self, fields):
        self._ck(self.L.xsb_ghost_reduce_add(self.h, sum(1 << f for f in fields)), "xsb_ghost_reduce_add")

    def cell_offsets(self):
        off = np.zeros(int(self.L.xsb_num_cells(self.h)) + 1, dtype=np.uint64)
        self._ck(self.L.xsb_cell_offsets_download(self.h, _ptr(off)), "xsb_cell_offsets_download")
        return off

    @property
    def n_own(self):
        return int(self.L.xsb_num_own_particles(self.h))


def ghost_plan(global_cells, periodic, rank_dims, rank_coord, ghost_layers):
    """receive list of one brick (host only, no GPU): int32 array [n, 6] = ghost_cell, owner_rank, owner_cell, wrap xyz"""
    L = load_library()
    d = DomainDesc()
    d.global_cells[:] = [int(v) for v in global_cells]; d.periodic[:] = [int(v) for v in periodic]
    d.rank_dims[:] = [int(v) for v in rank_dims]; d.rank_coord[:] = [int(v) for v in rank_coord]
    rc = np.ascontiguousarray(rank_coord, dtype=np.int32)
    n = C.c_uint64()
    r = L.xsb_ghost_plan(C.byref(d), int(ghost_layers), _ptr(rc), None, 0, C.byref(n))
    if r != 0:
        raise XsbError("xsb_ghost_plan failed (%d)" % r)
    out = np.zeros((max(1, n.value), 6), dtype=np.int32)
    r = L.xsb_ghost_plan(C.byref(d), int(ghost_layers), _ptr(rc), _ptr(out), n.value, C.byref(n))
    if r != 0:
        raise XsbError("xsb_ghost_plan failed (%d)" % r)
    return out[:n.value]


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = load_library().xsb_comm_unique_id(buf)
    if rc != 0:
        raise XsbError("xsb_comm_unique_id failed (%d): NCCL not loadable" % rc)
    return buf.raw
