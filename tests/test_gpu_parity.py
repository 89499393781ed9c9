"""GPU parity tests (run with -m gpu on a B200): every operator is called through the C ABI (ctypes ->
libxsb200.so) and compared with the CPU oracle on the same seeded input.
Bars: neighbour lists bit-exact (exported uint16 stream == oracle stream, byte for byte);
forces / energies / virial within 1e-10 (FP64) or 1e-5 (mixed) relative to the largest magnitude of the field."""
import os

import numpy as np
import pytest

import exastamp_b200 as xsb
from helpers import EV, GridSystem, SC_CU, SC_XX, johnson_params, lattice, write_setfl

pytestmark = pytest.mark.gpu
TOL64, TOLMIX = 1e-10, 1e-5


def oracle():
    from oracle import oracle as O
    return O


def rel_err(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def make_ctx(gs):
    ctx = xsb.Context(0)
    ctx.grid_set(xsb.make_grid(gs.dims, gs.gl, gs.cell_size, gs.origin, gs.xform))
    ctx.particles_set_cells(gs.cell_off)
    ctx.upload(xsb.F_RX, gs.rx); ctx.upload(xsb.F_RY, gs.ry); ctx.upload(xsb.F_RZ, gs.rz)
    ctx.upload(xsb.F_TYPE, gs.type)
    return ctx


def system(structure="FCC", ncells=6, a=5.0, sigma=0.1, cell=None, gl=2, seed=1, types=None, xform=None):
    pos, typ, box = lattice(structure, ncells, a, sigma, seed=seed, types=types)
    return GridSystem(pos, typ, box, cell if cell else a, gl, xform=xform)


# ------------------------------------------------------------------------------------------------ a2
@pytest.mark.parametrize("chunk_size,offsets", [(1, True), (1, False), (4, True), (8, True)])
def test_chunk_neighbors_stream_bit_exact(chunk_size, offsets):
    O = oracle()
    gs = system(ncells=6, a=5.0, gl=2)
    nb = O.Neighbors.build(gs.oracle_grid(), gs.cell_off, gs.rx, gs.ry, gs.rz, 9.0, chunk_size, offsets)
    ooff, odata = nb.export()
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0, chunk_size, offsets)
    goff, gdata = ctx.chunk_neighbors_export()
    assert np.array_equal(ooff, goff)
    assert odata.tobytes() == gdata.tobytes()


def test_chunk_neighbors_flat_list_and_stats():
    O = oracle()
    gs = system(ncells=5, a=5.0, gl=2, seed=3)
    nb = O.Neighbors.build(gs.oracle_grid(), gs.cell_off, gs.rx, gs.ry, gs.rz, 9.0, 1, True)
    cnt, off, idx = nb.decode()
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    gcnt, goff, gidx = ctx.chunk_neighbors_flat()
    assert np.array_equal(cnt, gcnt) and np.array_equal(off, goff) and np.array_equal(idx, gidx)
    total, mx = ctx.chunk_neighbors_stats()
    assert total == int(cnt.sum()) and mx == int(cnt.max())


def test_chunk_neighbors_triclinic_xform_bit_exact():
    O = oracle()
    X = np.array([[1.02, 0.03, 0.01], [0.0, 0.98, 0.02], [0.0, 0.0, 1.01]])
    gs = system(ncells=6, a=5.0, gl=2, xform=X)
    nb = O.Neighbors.build(gs.oracle_grid(), gs.cell_off, gs.rx, gs.ry, gs.rz, 9.0, 1, True)
    ooff, odata = nb.export()
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    goff, gdata = ctx.chunk_neighbors_export()
    assert np.array_equal(ooff, goff) and odata.tobytes() == gdata.tobytes()


def test_chunk_neighbors_empty_and_ragged_cells():
    O = oracle()
    rng = np.random.default_rng(5)
    box = np.array([30.0, 30.0, 30.0])
    pos = rng.uniform(0, 30.0, (400, 3))
    pos = pos[(pos[:, 0] < 12) | (pos[:, 0] > 22)]          # a slab of empty cells
    gs = GridSystem(pos, np.zeros(len(pos), np.uint8), box, 5.0, 2)
    nb = O.Neighbors.build(gs.oracle_grid(), gs.cell_off, gs.rx, gs.ry, gs.rz, 8.5, 4, True)
    ooff, odata = nb.export()
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(8.5, 4, True)
    goff, gdata = ctx.chunk_neighbors_export()
    assert np.array_equal(ooff, goff) and odata.tobytes() == gdata.tobytes()


def test_empty_grid():
    ctx = xsb.Context(0)
    ctx.grid_set(xsb.make_grid([4, 4, 4], 1, 5.0, [-5.0] * 3))
    ctx.particles_set_cells(np.zeros(65, dtype=np.uint64))
    ctx.chunk_neighbors(6.0)
    assert ctx.chunk_neighbors_stats() == (0, 0)
    ctx.zero_force_energy()
    ctx.pair_force([1.0, 3.0], 5.0)
    off, data = ctx.chunk_neighbors_export()
    assert data.size == 0 and np.all(off == 0)


# ------------------------------------------------------------------------------------------------ a4/a5
def lj_reference(gs, rcut, ghost, want_vir, nbh_dist=9.0, params=(0.0104 * EV, 3.4)):
    O = oracle()
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh_dist, 1, True)
    fx, fy, fz, ep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if want_vir else None
    O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, params, rcut, ghost, fx, fy, fz, ep, vir)
    return fx, fy, fz, ep, vir


@pytest.mark.parametrize("ghost,virial", [(False, False), (False, True), (True, True)])
def test_lj_compute_force_parity(ghost, virial):
    gs = system(ncells=6, a=5.0, gl=2)
    fx, fy, fz, ep, vir = lj_reference(gs, 8.0, ghost, virial)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    ctx.zero_force_energy(ghost=True)
    flags = xsb.FLAG_ENERGY | (xsb.FLAG_GHOST if ghost else 0) | (xsb.FLAG_VIRIAL if virial else 0)
    ctx.pair_force([0.0104 * EV, 3.4], 8.0, flags)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64
    if not ghost:
        assert np.all(ctx.download(xsb.F_FX)[gs.is_ghost] == 0.0)


def test_lj_accumulates_and_zero_force_energy():
    gs = system(ncells=5, a=5.0, gl=2)
    fx, *_ = lj_reference(gs, 8.0, False, False)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    ctx.zero_force_energy()
    ctx.pair_force([0.0104 * EV, 3.4], 8.0)
    ctx.pair_force([0.0104 * EV, 3.4], 8.0)          # chained operators add up (compute_force: [a, b])
    assert rel_err(ctx.download(xsb.F_FX), 2 * fx) < TOL64
    ctx.zero_force_energy()
    assert np.all(ctx.download(xsb.F_FX) == 0.0) and np.all(ctx.download(xsb.F_EP) == 0.0)


def test_lj_mixed_precision_parity():
    gs = system(ncells=6, a=5.0, gl=2)
    fx, fy, fz, ep, _ = lj_reference(gs, 8.0, False, False)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    ctx.zero_force_energy()
    ctx.pair_force([0.0104 * EV, 3.4], 8.0, xsb.FLAG_ENERGY | xsb.FLAG_MIXED)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOLMIX


def test_lj_triclinic_xform_parity():
    X = np.array([[1.02, 0.03, 0.01], [0.0, 0.98, 0.02], [0.0, 0.0, 1.01]])
    gs = system(ncells=6, a=5.0, gl=2, xform=X)
    fx, fy, fz, ep, vir = lj_reference(gs, 8.0, False, True)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    ctx.zero_force_energy()
    ctx.pair_force([0.0104 * EV, 3.4], 8.0, xsb.FLAG_ENERGY | xsb.FLAG_VIRIAL)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep), (xsb.F_VIRIAL, vir)):
        assert rel_err(ctx.download(f), ref) < TOL64


# ------------------------------------------------------------------------------------------------ a6
@pytest.mark.parametrize("virial", [False, True])
def test_lj_multi_force_parity(virial):
    O = oracle()
    gs = system(ncells=6, a=5.0, gl=2, types=[0, 1, 0, 1])
    # rows by unique_pair_id: (0,0), (0,1), (1,1) : eps, sigma, rcut   (values of potentials/pair/lj/multi_species_nosym.msp style)
    rows = np.array([[0.0104 * EV, 3.4, 8.0], [0.0150 * EV, 3.2, 7.0], [0.0200 * EV, 3.0, 6.5]])
    ecut = []
    for e, s, rc in rows:
        q = (s / rc) ** 2; q6 = q * q * q
        ecut.append(4 * e * (q6 * q6 - q6))
    pp = np.column_stack([rows, np.array(ecut)])
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 9.0, 1, True)
    fx, fy, fz, ep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, pp, 8.0, 0, fx, fy, fz, ep, vir)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    ctx.zero_force_energy()
    ctx.pair_multi_force(2, rows, 8.0, xsb.FLAG_ENERGY | (xsb.FLAG_VIRIAL if virial else 0))
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64


# ------------------------------------------------------------------------------------------------ a7
@pytest.mark.parametrize("virial", [False, True])
def test_johnson_force_parity(virial):
    O = oracle()
    gs = system(ncells=5, a=3.615, sigma=0.05, cell=3.615, gl=4)     # ghost thickness >= 2*rcut + skin
    rcut, nbh = 5.5, 6.5
    p = johnson_params()
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    fx, fy, fz, ep, emb = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    O.eam_johnson(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, p, rcut, 7, fx, fy, fz, ep, vir, emb)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_johnson_force(p, rcut, 7, xsb.FLAG_VIRIAL if virial else 0)
    assert rel_err(ctx.download(xsb.F_RHO_DEMB), emb) < TOL64
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64


# sutton_chen / vniitf: the other analytic single-species models of eam_potential_template (xsb_eam_analytic_force)
JOULE = 1.0 / 1.602176634e-19 * EV
EAM1_CASES = {
    # parameter sets of the reference's regression decks potentials/eam/eam_sutton_chen/single_specy.msp, eam_vniitf/single_specy.msp
    "sutton_chen": (xsb.EAM_SUTTON_CHEN, [3.317e1, 3.605e-21 * JOULE, 3.27, 9.05, 5.005], 5.5, 6.5),
    "vniitf": (xsb.EAM_VNIITF, [5.599, 1.0, 3.437, 2.956031e-19 * JOULE, 5.15003855e-20 * JOULE, 6.0, 1.401, 7.618, 0.724, 3.072, 0.145, 2.72, -1.87], 5.599, 6.5),
    "vniitf_narrow_switch": (xsb.EAM_VNIITF, [5.5, 4.9, 3.44, 3.1 * EV, 0.02 * EV, 5.1, 1.05, 10.0, 0.62, 3.7, 0.08, 6.0, 2.5], 5.5, 6.5),
}


@pytest.mark.parametrize("name", sorted(EAM1_CASES))
@pytest.mark.parametrize("virial,two_step", [(False, False), (True, True)])
def test_eam_analytic_models_parity(name, virial, two_step):
    """<name>_force, and <name>_emb followed by <name>_force_reuse_emb (the force pass then consumes the cached rho'(r))"""
    model, p, rcut, nbh = EAM1_CASES[name]
    O = oracle()
    gs = system(ncells=5, a=3.615, sigma=0.05, cell=3.615, gl=4)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    fx, fy, fz, ep, emb = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    O.eam_analytic(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, model, p, rcut, 7, fx, fy, fz, ep, vir, emb)
    assert np.abs(fx).max() > 0 and np.all(np.isfinite(emb))
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    fl = xsb.FLAG_VIRIAL if virial else 0
    if two_step:
        ctx.eam_analytic_force(model, p, rcut, 3, fl)
        ctx.eam_analytic_force(model, p, rcut, 4, fl)
    else:
        ctx.eam_analytic_force(model, p, rcut, 7, fl)
    assert rel_err(ctx.download(xsb.F_RHO_DEMB), emb) < TOL64
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64, (name, f)
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64
    # a johnson call with other parameters must not reuse this model's cached pair values
    with pytest.raises(xsb.XsbError):
        ctx.eam_analytic_force(model, p[:-1], rcut, 7, 0)


# ------------------------------------------------------------------------------------------------ a8
def eam_alloy_case(tmp_path, elements, types, eflag, virial, two_step):
    O = oracle()
    path = str(tmp_path / "synthetic.eam.alloy")
    write_setfl(path, elements, nrho=2000, drho=0.1, nr=2000, rc=6.0)
    gs = system(ncells=5, a=3.615, sigma=0.05, cell=3.615, gl=4, types=types)
    rcut, nbh = 6.0, 7.0
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    eam = O.EamAlloy(path)
    fx, fy, fz, ep, emb = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    flags = 1 | 2 | 4 | 8 | (16 if eflag else 0) | (32 if virial else 0)
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, rcut, flags, fx, fy, fz, ep, vir, emb)
    ctx = make_ctx(gs)
    info = ctx.eam_alloy_load(path)
    assert info["nelements"] == len(elements) and info["nr"] == 2000
    ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    ef = xsb.EAM_EFLAG if eflag else 0
    fl = xsb.FLAG_VIRIAL if virial else 0
    if two_step:
        ctx.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | ef, fl)
        ctx.eam_alloy_force(rcut, xsb.EAM_FORCE | ef, fl)
    else:
        ctx.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_FORCE | ef, fl)
    assert rel_err(ctx.download(xsb.F_RHO_DEMB), emb) < TOL64
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64, "field %d" % f
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64


@pytest.mark.parametrize("eflag,virial,two_step", [(False, False, False), (True, False, True), (True, True, False)])
def test_eam_alloy_single_species_parity(tmp_path, eflag, virial, two_step):
    eam_alloy_case(tmp_path, [SC_CU], None, eflag, virial, two_step)


@pytest.mark.parametrize("eflag,virial", [(False, False), (True, True)])
def test_eam_alloy_two_species_parity(tmp_path, eflag, virial):
    eam_alloy_case(tmp_path, [SC_CU, SC_XX], [0, 1, 1, 0], eflag, virial, False)


# ------------------------------------------------------------------------------------------------ properties at scale
def test_c1_full_size_properties():
    """BASELINE config C1 (131 072 atoms): momentum conservation and symmetry of the list at full size."""
    pos, typ, box = lattice("FCC", 32, 5.0, 0.1, seed=1)
    gs = GridSystem(pos, typ, box, 160.0 / 17, 1)                   # 17^3 cells of 9.41 ang >= rc + skin
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(9.0)
    cnt, off, idx = ctx.chunk_neighbors_flat()
    own = ~gs.is_ghost
    assert 80 < cnt[own].mean() < 92
    ctx.zero_force_energy()
    ctx.pair_force([0.0104 * EV, 3.4], 8.0)
    fx, fy, fz = ctx.download(xsb.F_FX), ctx.download(xsb.F_FY), ctx.download(xsb.F_FZ)
    fmax = np.abs(np.concatenate([fx, fy, fz])).max()
    for f in (fx, fy, fz):
        assert abs(f[own].sum()) < 1e-9 * fmax * np.sqrt(own.sum())   # Newton's third law holds globally
    # every owned pair (a,b) in the list has its mirror image listed as well: pair count is even per distance
    ep = ctx.download(xsb.F_EP)
    assert -0.09 < ep[own].mean() / EV < -0.03


# ------------------------------------------------------------------------------------------------ a10 + grid assembly
def assigned_ctx(pos, typ, box, cell, gl, vel=None):
    n_own = np.round(box / cell).astype(int)
    ctx = xsb.Context(0)
    ctx.grid_set(xsb.make_grid(n_own + 2 * gl, gl, cell, [-gl * cell] * 3))
    v = (None, None, None) if vel is None else (vel[:, 0], vel[:, 1], vel[:, 2])
    ctx.particles_assign(pos[:, 0], pos[:, 1], pos[:, 2], *v, typ=typ)
    ctx.set_domain(n_own)
    ctx.ghost_comm_scheme()
    return ctx


@pytest.mark.parametrize("ncells,cell,gl", [(6, 5.0, 2), (4, 5.0, 3), (5, 2.5, 4)])
def test_assign_and_ghost_scheme_equal_host_statement(ncells, cell, gl):
    pos, typ, box = lattice("FCC", ncells, 5.0, 0.1, seed=7, types=[0, 1, 0, 1])
    gs = GridSystem(pos, typ, box, cell, gl)
    ctx = assigned_ctx(pos, typ, box, cell, gl)
    assert ctx.n == gs.n and ctx.n_own == gs.n_owned
    assert np.array_equal(ctx.cell_offsets(), gs.cell_off)
    assert ctx.download(xsb.F_RX).tobytes() == gs.rx.tobytes()
    assert ctx.download(xsb.F_RY).tobytes() == gs.ry.tobytes()
    assert ctx.download(xsb.F_RZ).tobytes() == gs.rz.tobytes()
    assert np.array_equal(ctx.download(xsb.F_TYPE), gs.type)
    assert np.array_equal(ctx.download(xsb.F_ID), gs.src_index.astype(np.uint64))


def test_ghost_update_and_reduce_add():
    pos, typ, box = lattice("FCC", 5, 5.0, 0.1, seed=9)
    gs = GridSystem(pos, typ, box, 5.0, 2)
    ctx = assigned_ctx(pos, typ, box, 5.0, 2)
    rng = np.random.default_rng(11)
    # move owners, ghosts must follow with their periodic shift (ghost_update_r)
    d = rng.normal(0, 0.05, (len(pos), 3))
    newpos = pos + d
    rx = gs.rx.copy(); own = ~gs.is_ghost
    rx[own] = newpos[gs.src_index[own], 0]
    ctx.upload(xsb.F_RX, rx)
    ctx.ghost_update([xsb.F_RX])
    expect = newpos[gs.src_index, 0] + (gs.rx - pos[gs.src_index, 0])       # same shift as before
    got = ctx.download(xsb.F_RX)
    assert np.max(np.abs(got - expect)) < 1e-12
    # ghost_update_opt on an optional field
    emb = rng.normal(0, 1, gs.n); emb[~own] = 0
    ctx.upload(xsb.F_RHO_DEMB, emb)
    ctx.ghost_update([xsb.F_RHO_DEMB])
    per_atom = np.zeros(len(pos)); per_atom[gs.src_index[own]] = emb[own]
    assert np.array_equal(ctx.download(xsb.F_RHO_DEMB), per_atom[gs.src_index])
    # update_force_energy_from_ghost: owners receive the sum of their images
    f = rng.normal(0, 1, gs.n)
    ctx.upload(xsb.F_FX, f); ctx.upload(xsb.F_EP, f)
    ctx.ghost_reduce_add([xsb.F_FX, xsb.F_EP])
    tot = np.zeros(len(pos)); np.add.at(tot, gs.src_index, f)
    got = ctx.download(xsb.F_FX)
    assert np.max(np.abs(got[own] - tot[gs.src_index[own]])) < 1e-12
    assert np.max(np.abs(ctx.download(xsb.F_EP)[own] - tot[gs.src_index[own]])) < 1e-12
    # update_virial_force_energy_from_ghost (src/mpi/update_from_ghosts.cu:43): virial (9 per atom) + f + ep in one exchange
    v = rng.normal(0, 1, (gs.n, 9)); g3 = rng.normal(0, 1, (3, gs.n))
    ctx.upload(xsb.F_VIRIAL, v)
    for k, fld in enumerate((xsb.F_FX, xsb.F_FY, xsb.F_FZ)):
        ctx.upload(fld, g3[k])
    ctx.upload(xsb.F_EP, f)
    ctx.ghost_reduce_add([xsb.F_VIRIAL, xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP])
    vt = np.zeros((len(pos), 9)); np.add.at(vt, gs.src_index, v)
    assert np.max(np.abs(ctx.download(xsb.F_VIRIAL)[own] - vt[gs.src_index[own]])) < 1e-12
    ft = np.zeros(len(pos)); np.add.at(ft, gs.src_index, g3[1])
    assert np.max(np.abs(ctx.download(xsb.F_FY)[own] - ft[gs.src_index[own]])) < 1e-12


def test_nve_loop_pieces_and_rebin():
    """push_f_v_r / push_f_v / force_to_accel / backup_r / particle_displ_over / rebin against numpy."""
    pos, typ, box = lattice("FCC", 5, 5.0, 0.1, seed=13)
    rng = np.random.default_rng(17)
    vel = rng.normal(0, 2.0, pos.shape)
    gs = GridSystem(pos, typ, box, 5.0, 2)
    ctx = assigned_ctx(pos, typ, box, 5.0, 2, vel)
    own = ~gs.is_ghost
    f = rng.normal(0, 50.0, (gs.n, 3))
    for k, fld in enumerate((xsb.F_FX, xsb.F_FY, xsb.F_FZ)):
        ctx.upload(fld, f[:, k])
    mass = 39.948
    ctx.force_to_accel([mass])
    ax = ctx.download(xsb.F_FX)
    assert np.allclose(ax[own], f[own, 0] / mass, rtol=1e-15) and np.array_equal(ax[~own], f[~own, 0])
    ctx.backup_r()
    dt = 1e-3
    ctx.push_f_v_r(dt)
    x = ctx.download(xsb.F_RX)
    v0 = vel[gs.src_index, 0]
    assert np.allclose(x[own], gs.rx[own] + v0[own] * dt + 0.5 * ax[own] * dt * dt, rtol=1e-15, atol=1e-15)
    ctx.push_f_v(0.5 * dt)
    assert np.allclose(ctx.download(xsb.F_VX)[own], v0[own] + ax[own] * 0.5 * dt, rtol=1e-15)
    over, dmax = ctx.particle_displ_over(0.5)
    disp = np.sqrt(sum((ctx.download(fl)[own] - g[own]) ** 2 for fl, g in ((xsb.F_RX, gs.rx), (xsb.F_RY, gs.ry), (xsb.F_RZ, gs.rz))))
    assert not over and abs(dmax - disp.max()) < 1e-12
    over, _ = ctx.particle_displ_over(disp.max() * 0.5)
    assert over
    # big move, then move_particles: wrap + re-bin + ghosts equals binning the moved positions from scratch
    ctx.push_f_v_r(1.0)
    newp = np.stack([ctx.download(fl)[own] for fl in (xsb.F_RX, xsb.F_RY, xsb.F_RZ)], axis=1)
    ids = ctx.download(xsb.F_ID)[own]
    ctx.particles_rebin(); ctx.ghost_comm_scheme()
    wrapped = newp - np.floor(newp / box) * box
    wrapped = np.where(wrapped >= box, 0.0, wrapped)
    gs2 = GridSystem(wrapped, typ, box, 5.0, 2)
    assert np.array_equal(ctx.cell_offsets(), gs2.cell_off)
    assert np.max(np.abs(ctx.download(xsb.F_RX) - gs2.rx)) < 1e-9
    assert np.array_equal(ctx.download(xsb.F_ID), ids[gs2.src_index])


# ---- a9 snap_force ---------------------------------------------------------------------------------------------------
def snap_case(twoj, nel, seed=4, structure="BCC", ncells=5, a=3.3, gl=1):
    O = oracle()
    rng = np.random.default_rng(seed)
    types = None if nel == 1 else [0, 1]
    pos, typ, box = lattice(structure, ncells, a, 0.07, seed=seed, types=types)
    gs = GridSystem(pos, typ, box, box[0] / 3, gl)
    rad = [0.5, 0.46][:nel]; wj = [1.0, 0.8][:nel]
    nc = xsb.load_library().xsb_snap_ncoeff(twoj)
    beta = rng.normal(0.0, 1.0, (nel, nc + 1)) * EV * 1e-3
    return gs, O.Snap(twoj, 4.7, rad, wj, beta), (twoj, 4.7, rad, wj, beta)


@pytest.mark.parametrize("twoj,nel,virial", [(8, 1, True), (6, 2, False), (3, 1, False), (4, 1, True)])
def test_snap_force_parity(twoj, nel, virial):
    O = oracle()
    gs, S, args = snap_case(twoj, nel)
    g = gs.oracle_grid()
    nbh_dist = S.rcut_max() + 0.5
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh_dist, 1, True)
    rfx, rfy, rfz, rep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    rvir = np.zeros((gs.n, 9)) if virial else None
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2 | (4 if virial else 0), rfx, rfy, rfz, rep, rvir)
    ctx = make_ctx(gs)
    rc = ctx.snap_set(*args)
    assert abs(rc - S.rcut_max()) < 1e-12
    ctx.chunk_neighbors(nbh_dist)
    ctx.zero_force_energy(ghost=True)
    ctx.snap_force(xsb.FLAG_ENERGY | (xsb.FLAG_VIRIAL if virial else 0))
    assert not ctx.snap_overflow()
    fx, fy, fz, ep = [ctx.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]
    fmax = max(np.abs(rfx).max(), np.abs(rfy).max(), np.abs(rfz).max())
    # FP64 mode tolerance of the north star: 1e-10 relative (forces on ghost copies included: Newton-on scatter)
    for a, b in ((fx, rfx), (fy, rfy), (fz, rfz)):
        assert np.abs(a - b).max() <= 1e-10 * fmax
    own = ~gs.is_ghost
    assert np.abs(ep[own] - rep[own]).max() <= 1e-10 * np.abs(rep[own]).max()
    if virial:
        vir = ctx.download(xsb.F_VIRIAL)
        assert np.abs(vir[own] - rvir[own]).max() <= 1e-10 * np.abs(rvir[own]).max()


def test_eam_force_after_positions_changed_does_not_reuse_stale_sublist(tmp_path):
    """the rho pass leaves an in-range sub-list for the force pass of the SAME positions; if the caller moves atoms in
    between (here: uploads new positions, keeps rho_dEmb), the force pass must fall back to the full list"""
    O = oracle()
    gs = system(ncells=5, a=3.615, sigma=0.05, cell=3.615 * 5 / 2, gl=1, seed=3)
    path = write_setfl(str(tmp_path / "cu.eam.alloy"), [SC_CU], nrho=800, drho=0.25, nr=900, rc=5.6)
    g = gs.oracle_grid(); eam = O.EamAlloy(path)
    ctx = make_ctx(gs); ctx.eam_alloy_load(path); ctx.chunk_neighbors(8.0)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(5.6, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST)
    emb = ctx.download(xsb.F_RHO_DEMB)
    rng = np.random.default_rng(9)
    own_shift = rng.normal(0, 0.12, (int(gs.src_index.max()) + 1, 3))
    rx2, ry2, rz2 = gs.rx + own_shift[gs.src_index, 0], gs.ry + own_shift[gs.src_index, 1], gs.rz + own_shift[gs.src_index, 2]
    ctx.upload(xsb.F_RX, rx2); ctx.upload(xsb.F_RY, ry2); ctx.upload(xsb.F_RZ, rz2)
    ctx.eam_alloy_force(5.6, xsb.EAM_FORCE)
    fx = ctx.download(xsb.F_FX)
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 8.0, 1, True)     # the list is still the old one
    rfx, rfy, rfz, rep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.eam_alloy(g, gs.cell_off, rx2, ry2, rz2, gs.type, nb, eam, 5.6, 8, rfx, rfy, rfz, rep, None, emb.copy())
    assert rel_err(fx, rfx) < TOL64


# ------------------------------------------------------------------------------------------------ f2 thermodynamic state
@pytest.mark.parametrize("virial", [False, True])
def test_thermo_state_matches_the_reference_sums(virial):
    """simulation_thermodynamic_state (src/thermo_state/simulation_thermodynamic_state.cpp:81-230): the 27 sums over own
    cells, restated with numpy on the same arrays; reproducible bit for bit from call to call."""
    pos, typ, box = lattice("FCC", 6, 5.0, 0.1, seed=21, types=[0, 1, 0, 1])
    rng = np.random.default_rng(23)
    vel = rng.normal(0, 1.0, pos.shape)
    gs = GridSystem(pos, typ, box, 10.0, 1)
    ctx = assigned_ctx(pos, typ, box, 10.0, 1, vel=vel)
    own = ~gs.is_ghost
    ep = rng.normal(0, 1, gs.n)
    ctx.upload(xsb.F_EP, ep)
    vir = None
    if virial:
        ctx.zero_force_energy(ghost=True)
        ctx.upload(xsb.F_EP, ep)
        ctx.chunk_neighbors(9.0)
        ctx.pair_force([0.0104 * EV, 3.4], 8.0, xsb.FLAG_VIRIAL)          # allocates and fills the virial field
        vir = ctx.download(xsb.F_VIRIAL)
    masses = np.array([39.948, 63.546])
    t = ctx.thermo_state(masses)
    m = masses[gs.type[own]]
    v = vel[gs.src_index[own]]
    assert t["particle_count"] == own.sum() == len(pos)
    assert t["mass"] == pytest.approx(m.sum(), rel=1e-13)
    assert np.allclose(t["momentum"], (v * m[:, None]).sum(axis=0), rtol=0, atol=1e-9 * np.abs(v * m[:, None]).sum())
    assert np.allclose(t["kinetic_energy"], 0.5 * (v * v * m[:, None]).sum(axis=0), rtol=1e-12)
    ket = 0.5 * np.einsum("n,ni,nj->ij", m, v, v)
    assert np.allclose(t["ke_tensor"], ket, rtol=0, atol=1e-12 * np.abs(ket).max())
    assert t["potential_energy"] == pytest.approx(ep[own].sum(), abs=1e-10 * np.abs(ep).sum())
    if virial:
        ref = vir[own].sum(axis=0).reshape(3, 3)
        assert np.allclose(t["virial"], ref, rtol=0, atol=1e-12 * np.abs(vir[own]).sum())
    else:
        assert not t["virial"].any()
    t2 = ctx.thermo_state(masses)
    assert all(np.array_equal(np.asarray(t[k]), np.asarray(t2[k])) for k in t)


# ------------------------------------------------------------------------------------------------ f3: zbl / exp6 / buckingham
PAIR_CASES = {
    # pot id, raw parameters (C-ABI order), operator rcut, nbh_dist
    "zbl_Ta": (1, [0.1, 4.615858, 73, 73], 4.615858, 5.6),          # potentials/snap/monomat_zbl.msp
    "zbl_rcut_inside_rc": (1, [2.0, 4.8, 74, 4], 4.2, 5.2),          # operator cutoff below the potential's own rc: ecut != 0
    "zbl_rcut_beyond_rc": (1, [1.0, 3.6, 29, 29], 4.4, 5.4),         # pairs between rc and rcut contribute -ecut = 0 and no force
    "exp6": (2, [3.0e5 * EV, 3.6, 60.0 * EV, 1.0e-6 * EV], 5.5, 6.5),
    "buckingham": (3, [1.2e3 * EV, 0.32, 25.0 * EV], 5.5, 6.5),
    "yukawa": (4, [2.43 * EV, 4.1], 5.5, 6.5),                        # potentials/pair/yukawa/single_specy_nosym.msp:6 (`de` as the reference writes it)
    "yukawa_soft": (4, [25.0 * EV, 1.3], 5.5, 6.5),
    "relax": (5, [2.2, 3.9], 4.4, 5.4),                               # relax/potential.h:43-51: pairs below r1 and beyond rc are clamped
}


@pytest.mark.parametrize("name", sorted(PAIR_CASES))
@pytest.mark.parametrize("virial", [False, True])
def test_pair_potentials_parity(name, virial):
    pot, prm, rcut, nbh = PAIR_CASES[name]
    O = oracle()
    gs = system(ncells=6, a=3.3, sigma=0.08, cell=3.3, gl=2)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    fx, fy, fz, ep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, prm, rcut, 0, fx, fy, fz, ep, vir, pot=pot)
    assert np.abs(fx).max() > 0
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    ctx.pair_force(prm, rcut, xsb.FLAG_ENERGY | (xsb.FLAG_VIRIAL if virial else 0), pot=pot)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64, (name, f)
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOL64


def test_zero_potential_adds_nothing():
    """zero_compute_force (zero/potential.h:49-54): e = de = 0 for every pair"""
    gs = system(ncells=6, a=3.3, sigma=0.08, cell=3.3, gl=2)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(5.0)
    ctx.zero_force_energy(ghost=True)
    ctx.pair_force([], 4.0, xsb.FLAG_ENERGY, pot=xsb.POT_ZERO)
    for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP):
        assert not ctx.download(f).any()


@pytest.mark.parametrize("twoj", [8, 4])
def test_snap_more_than_64_neighbours_is_batched(twoj):
    """the reference has no cap on the neighbours inside the SNAP cut-off; the kernels hold 64 at a time in shared memory and
    batch the rest (2J = 8: Utot kernel + per-direction force kernel through the neighbour table; 2J = 4: in-kernel re-filter)"""
    O = oracle()
    rng = np.random.default_rng(8)
    pos, typ, box = lattice("BCC", 5, 3.3, 0.07, seed=6)
    gs = GridSystem(pos, typ, box, box[0] / 3, 2)
    nc = xsb.load_library().xsb_snap_ncoeff(twoj)
    beta = rng.normal(0.0, 1.0, (1, nc + 1)) * EV * 1e-3
    rcutfac = 7.2                                        # 2 * 0.5 * 7.2 = 7.2 ang: ~85 neighbours in BCC a = 3.3
    S = O.Snap(twoj, rcutfac, [0.5], [1.0], beta)
    g = gs.oracle_grid()
    nbh_dist = S.rcut_max() + 0.3
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh_dist, 1, True)
    rfx, rfy, rfz, rep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, rfx, rfy, rfz, rep, None)
    ctx = make_ctx(gs)
    ctx.snap_set(twoj, rcutfac, [0.5], [1.0], beta)
    ctx.chunk_neighbors(nbh_dist)
    cnt, _, _ = ctx.chunk_neighbors_flat()
    assert cnt[~gs.is_ghost].min() > 70
    ctx.zero_force_energy(ghost=True)
    ctx.snap_force(xsb.FLAG_ENERGY)
    assert not ctx.snap_overflow()
    fx, fy, fz, ep = [ctx.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]
    fmax = max(np.abs(rfx).max(), np.abs(rfy).max(), np.abs(rfz).max())
    for a, b in ((fx, rfx), (fy, rfy), (fz, rfz)):
        assert np.abs(a - b).max() <= 1e-10 * fmax
    own = ~gs.is_ghost
    assert np.abs(ep[own] - rep[own]).max() <= 1e-10 * np.abs(rep[own]).max()


@pytest.mark.parametrize("twoj,nel", [(8, 1), (6, 2), (4, 1)])
def test_snap_force_mixed_precision(twoj, nel):
    """XSB_FLAG_MIXED on snap_force = the reference's SNAP_FP32_MATH build (src/potential/snap/snap_force.cu:25-29, deck
    potentials/snap/multi_WBe_fp32.msp): FP32 Wigner / Clebsch-Gordan arithmetic, FP64 positions, forces and energies;
    north-star bar for mixed mode 1e-5 of the field maximum against the FP64 oracle"""
    O = oracle()
    gs, S, args = snap_case(twoj, nel)
    g = gs.oracle_grid()
    nbh_dist = S.rcut_max() + 0.5
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh_dist, 1, True)
    rfx, rfy, rfz, rep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, rfx, rfy, rfz, rep, None)
    ctx = make_ctx(gs)
    ctx.snap_set(*args)
    ctx.chunk_neighbors(nbh_dist)
    ctx.zero_force_energy(ghost=True)
    ctx.snap_force(xsb.FLAG_ENERGY | xsb.FLAG_MIXED)
    fx, fy, fz, ep = [ctx.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]
    fmax = max(np.abs(rfx).max(), np.abs(rfy).max(), np.abs(rfz).max())
    ef = max(np.abs(a - b).max() for a, b in ((fx, rfx), (fy, rfy), (fz, rfz))) / fmax
    own = ~gs.is_ghost
    ee = np.abs(ep[own] - rep[own]).max() / np.abs(rep[own]).max()
    print("snap mixed 2J=%d nel=%d: force err %.2e energy err %.2e" % (twoj, nel, ef, ee))
    assert ef < 1e-5 and ee < 1e-5
    assert ef > 1e-12                       # FP32 really ran


def test_zbl_multi_force_parity_and_mixed_precision():
    """zbl_multi_force as in potentials/snap/multi_WBe.msp: one row per type pair, z from the species"""
    O = oracle()
    gs = system(ncells=6, a=3.3, sigma=0.08, cell=3.3, gl=2, types=[0, 1, 1, 0])
    z = [74, 4]
    rows, orows = [], []
    for hi in range(2):
        for lo in range(hi + 1):
            prm = [4.0, 4.8, z[lo], z[hi]]
            rows.append(prm + [4.8])
            orows.append(prm + [4.8, O.pair_ecut(1, prm, 4.8)])
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 5.8, 1, True)
    fx, fy, fz, ep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, np.array(orows), 4.8, 0, fx, fy, fz, ep, None, pot=1)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(5.8)
    ctx.zero_force_energy(ghost=True)
    ctx.pair_multi_force(2, np.array(rows), 4.8, xsb.FLAG_ENERGY, pot=1)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < TOL64
    ctx.zero_force_energy(ghost=True)
    ctx.pair_multi_force(2, np.array(rows), 4.8, xsb.FLAG_ENERGY | xsb.FLAG_MIXED, pot=1)
    for f, ref in ((xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep)):
        assert rel_err(ctx.download(f), ref) < 1e-5            # mixed mode tolerance of the north star
    # wrong parameter count is rejected, nothing is computed on the host
    with pytest.raises(xsb.XsbError):
        ctx.pair_force([0.1, 4.6], 4.6, pot=1)
    with pytest.raises(xsb.XsbError):
        ctx.pair_force([1.0, 2.0], 4.6, pot=7)


# ------------------------------------------------------------------------------------------------ kernel switches / caches
def _eam_forces(gs, path, rcut, nbh, two_species, env=None):
    """fx,fy,fz,ep,rho_dEmb of eam_alloy_force (rho | rho2emb | ghost, then force: the bench's call pattern) under `env`"""
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        ctx = make_ctx(gs)              # the switches are read when the context is created
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctx.eam_alloy_load(path); ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG)
    ctx.eam_alloy_force(rcut, xsb.EAM_FORCE | xsb.EAM_EFLAG)
    return [ctx.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP, xsb.F_RHO_DEMB)]


@pytest.mark.parametrize("two_species", [False, True])
def test_eam_pair_cache_and_list_order_switches_do_not_change_results(tmp_path, two_species):
    """the force pass either re-evaluates rho'(r) or takes it from the per-pair cache left by the rho pass (same
    arithmetic: identical bits), and the tile list may be in canonical or bank-dealt order (other summation order:
    rounding only); every variant matches the oracle at the FP64 bar"""
    O = oracle()
    els = [SC_CU, SC_XX] if two_species else [SC_CU]
    path = write_setfl(str(tmp_path / "t.eam.alloy"), els, nrho=2000, drho=0.1, nr=2000, rc=6.0)
    gs = system(ncells=6, a=3.615, sigma=0.08, cell=3.615 * 2, gl=1, types=[0, 1, 1, 0] if two_species else None, seed=11)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 7.0, 1, True)
    ref = [gs.zeros() for _ in range(5)]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, O.EamAlloy(path), 6.0, 1 | 2 | 4 | 8 | 16, ref[0], ref[1], ref[2], ref[3], None, ref[4])
    base = _eam_forces(gs, path, 6.0, 7.0, two_species)
    nocache = _eam_forces(gs, path, 6.0, 7.0, two_species, {"XSB_NO_PAIR_CACHE": "1"})
    dealt = _eam_forces(gs, path, 6.0, 7.0, two_species, {"XSB_TILE_DEAL": "1"})
    for k in range(5):
        assert rel_err(base[k], ref[k]) < TOL64 and rel_err(dealt[k], ref[k]) < TOL64
        assert np.array_equal(base[k], nocache[k]), "field %d: cached rho'(r) differs from the re-evaluated one" % k
        assert rel_err(dealt[k], base[k]) < 1e-13


def test_dealt_list_order_keeps_the_canonical_export_and_lj_parity():
    O = oracle()
    gs = system(ncells=6, a=5.0, sigma=0.15, cell=10.0, gl=1, seed=5)
    os.environ["XSB_TILE_DEAL"] = "1"
    try:
        ctx = make_ctx(gs)
    finally:
        os.environ.pop("XSB_TILE_DEAL", None)
    ctx.chunk_neighbors(9.0)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 9.0, 1, True)
    goff, gdata = ctx.chunk_neighbors_export(); ooff, odata = nb.export()
    assert np.array_equal(goff, ooff) and gdata.tobytes() == odata.tobytes()
    ctx.zero_force_energy(ghost=True)
    ctx.pair_force([0.0104 * EV, 3.4], 8.0)
    rfx, rfy, rfz, rep = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, [0.0104 * EV, 3.4], 8.0, 0, rfx, rfy, rfz, rep, None)
    own = ~gs.is_ghost
    for f, r in ((xsb.F_FX, rfx), (xsb.F_FY, rfy), (xsb.F_FZ, rfz), (xsb.F_EP, rep)):
        assert rel_err(ctx.download(f)[own], r[own]) < TOL64


def test_johnson_pair_cache_parity_two_phase():
    """johnson_emb then johnson_force_reuse_emb as two calls: the force pass takes rho'(r) from the emb pass's cache"""
    O = oracle()
    gs = system(ncells=6, a=3.615, sigma=0.06, cell=3.615 * 2, gl=1, seed=4)
    p = johnson_params()
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 6.5, 1, True)
    rfx, rfy, rfz, rep, emb = [gs.zeros() for _ in range(5)]
    O.eam_johnson(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, p, 5.8, 1 | 2 | 4, rfx, rfy, rfz, rep, None, emb)
    ctx = make_ctx(gs); ctx.chunk_neighbors(6.5); ctx.zero_force_energy(ghost=True)
    ctx.eam_johnson_force(p, 5.8, 1 | 2)
    ctx.eam_johnson_force(p, 5.8, 4)
    own = ~gs.is_ghost
    for f, r in ((xsb.F_FX, rfx), (xsb.F_FY, rfy), (xsb.F_FZ, rfz), (xsb.F_EP, rep)):
        assert rel_err(ctx.download(f)[own], r[own]) < TOL64


# ------------------------------------------------------------------------------------------------ C2 at full size
def test_c2_full_size_parity_and_properties(tmp_path):
    """BASELINE configs[1] at full size (FCC Cu 79^3 unit cells = 1 972 156 atoms, the bench's potential and cutoffs):
    the whole force field against the oracle (OpenMP, tens of seconds), plus size-independent properties:
    momentum conservation, the in-range sub-list equals the list filtered at rcut, cohesive energy in the Cu range."""
    O = oracle()
    from bench import A_CU, RCUT, SKIN, make_setfl, n_cells_for
    path = make_setfl(str(tmp_path))
    pos, typ, box = lattice("FCC", 79, A_CU, 0.1, seed=1)
    nc = n_cells_for(box[0])
    gs = GridSystem(pos, typ, box, box[0] / nc, 1)
    assert gs.n_owned == 1972156
    ctx = make_ctx(gs); ctx.eam_alloy_load(path); ctx.chunk_neighbors(RCUT + SKIN)
    own = ~gs.is_ghost
    owner_of = np.zeros(len(pos), dtype=np.int64); owner_of[gs.src_index[own]] = np.nonzero(own)[0]
    ctx.zero_force_energy(ghost=True)
    # one ghost layer: rho on own atoms, F'(rho) copied owner -> ghost image (ghost_update_opt in the decks), then forces
    ctx.eam_alloy_force(RCUT, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_EFLAG)
    demb = ctx.download(xsb.F_RHO_DEMB)
    ctx.upload(xsb.F_RHO_DEMB, demb[owner_of[gs.src_index]])
    ctx.eam_alloy_force(RCUT, xsb.EAM_FORCE | xsb.EAM_EFLAG)
    fx, fy, fz, ep = [ctx.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]
    fmax = max(np.abs(f[own]).max() for f in (fx, fy, fz))
    for f in (fx, fy, fz):
        assert abs(f[own].sum()) < 1e-9 * fmax * np.sqrt(own.sum())          # Newton's third law, globally
    assert -4.2 < ep[own].mean() / EV < -2.8                                   # Sutton-Chen Cu cohesive energy (eV/atom)
    total, mx = ctx.chunk_neighbors_stats()
    assert 190 < total / gs.n < 200 and mx < 260
    # the same fields from the CPU restatement on the same 2.37 M particles (ghost images included)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, RCUT + SKIN, 1, True)
    cnt, off, idx = ctx.chunk_neighbors_flat()
    ocnt, ooff, oidx = nb.decode()
    assert np.array_equal(cnt, ocnt) and np.array_equal(off, ooff) and np.array_equal(idx, oidx)      # neighbour list bit-exact at 4.6e8 entries
    del ocnt, ooff, oidx, cnt, off, idx
    rfx, rfy, rfz, rep, emb = [gs.zeros() for _ in range(5)]
    eam = O.EamAlloy(path)
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, RCUT, 1 | 2 | 16, rfx, rfy, rfz, rep, None, emb)
    emb[:] = emb[owner_of[gs.src_index]]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, RCUT, 8 | 16, rfx, rfy, rfz, rep, None, emb)
    for a, b in ((fx, rfx), (fy, rfy), (fz, rfz), (ep, rep)):
        assert rel_err(a[own], b[own]) < TOL64


# ------------------------------------------------------------------------------------------------ C5 kernel configuration
@pytest.mark.parametrize("virial", [False, True])
def test_c5_two_species_alloy_triclinic_two_phase(tmp_path, virial):
    """BASELINE configs[4] in small: random two-species FCC alloy under an upper-triangular xform (NPT cell), eam_alloy_force
    driven in two calls (rho | rho2emb | ghost, then force through the in-range sub-list + per-pair cache), energy and
    virial on, plus lj_multi_force accumulated on top of it in the same force arrays"""
    O = oracle()
    X = np.array([[1.015, 0.02, -0.01], [0.0, 0.99, 0.015], [0.0, 0.0, 1.005]])
    path = write_setfl(str(tmp_path / "ab.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    rng = np.random.default_rng(21)
    pos, typ, box = lattice("FCC", 6, 3.615, 0.06, seed=8)
    typ = (rng.random(len(pos)) < 0.4).astype(np.uint8)                      # random alloy, 40 % of species 1
    gs = GridSystem(pos, typ, box, 3.615 * 2, 2, xform=X)
    g = gs.oracle_grid()
    nbh, rcut = 6.9, 6.0
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    rfx, rfy, rfz, rep, emb = [gs.zeros() for _ in range(5)]
    rvir = np.zeros((gs.n, 9)) if virial else None
    fl = 16 | (32 if virial else 0)
    eam = O.EamAlloy(path)
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, rcut, 1 | 2 | 4 | fl, rfx, rfy, rfz, rep, rvir, emb)
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, rcut, 8 | fl, rfx, rfy, rfz, rep, rvir, emb)
    rows = np.array([[0.0104 * EV, 2.3, 6.0], [0.0150 * EV, 2.2, 5.5], [0.0200 * EV, 2.1, 5.0]])
    ecut = [4 * e * ((s / rc) ** 12 - (s / rc) ** 6) for e, s, rc in rows]
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, np.column_stack([rows, ecut]), rcut, 0, rfx, rfy, rfz, rep, rvir)
    ctx = make_ctx(gs); ctx.eam_alloy_load(path); ctx.chunk_neighbors(nbh)
    ctx.zero_force_energy(ghost=True)
    vf = xsb.FLAG_VIRIAL if virial else 0
    ctx.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG, vf)
    ctx.eam_alloy_force(rcut, xsb.EAM_FORCE | xsb.EAM_EFLAG, vf)
    ctx.pair_multi_force(2, rows, rcut, xsb.FLAG_ENERGY | vf)
    own = ~gs.is_ghost
    assert rel_err(ctx.download(xsb.F_RHO_DEMB), emb) < TOL64
    for f, r in ((xsb.F_FX, rfx), (xsb.F_FY, rfy), (xsb.F_FZ, rfz), (xsb.F_EP, rep)):
        assert rel_err(ctx.download(f)[own], r[own]) < TOL64
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL).reshape(-1, 9)[own], rvir[own]) < TOL64


def test_grid_set_xform_keeps_list_and_matches_fresh_context(tmp_path):
    """NPT: the cell matrix changes between steps, the grid / particles / neighbour list stay.  Forces after
    xsb_grid_set_xform equal the oracle evaluated with the new matrix on the OLD list, and the stale sub-list is not reused"""
    O = oracle()
    X0 = np.array([[1.0, 0.02, 0.0], [0.0, 1.0, 0.01], [0.0, 0.0, 1.0]])
    X1 = X0 * 1.004
    path = write_setfl(str(tmp_path / "cu.eam.alloy"), [SC_CU], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    gs = system(ncells=6, a=3.615, sigma=0.05, cell=3.615 * 2, gl=2, xform=X0, seed=12)
    g0 = gs.oracle_grid()
    nb = O.Neighbors.build(g0, gs.cell_off, gs.rx, gs.ry, gs.rz, 7.0, 1, True)       # list of the old cell matrix
    g1 = O.make_grid(gs.dims, gs.gl, gs.cell_size, gs.origin, X1)
    ref = [gs.zeros() for _ in range(5)]
    O.eam_alloy(g1, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, O.EamAlloy(path), 6.0, 1 | 2 | 4 | 8 | 16, ref[0], ref[1], ref[2], ref[3], None, ref[4])
    ctx = make_ctx(gs); ctx.eam_alloy_load(path); ctx.chunk_neighbors(7.0)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST)          # leaves a sub-list for X0
    ctx.grid_set_xform(X1)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG)
    ctx.eam_alloy_force(6.0, xsb.EAM_FORCE | xsb.EAM_EFLAG)
    own = ~gs.is_ghost
    for f, r in ((xsb.F_FX, ref[0]), (xsb.F_FY, ref[1]), (xsb.F_FZ, ref[2]), (xsb.F_EP, ref[3])):
        assert rel_err(ctx.download(f)[own], r[own]) < TOL64
    # a force pass right after a matrix change must not walk the sub-list of the previous matrix
    ctx.grid_set_xform(X0)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(6.0, xsb.EAM_FORCE)
    ref0 = [gs.zeros() for _ in range(4)]
    O.eam_alloy(g0, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, O.EamAlloy(path), 6.0, 8, ref0[0], ref0[1], ref0[2], ref0[3], None, ctx.download(xsb.F_RHO_DEMB))
    assert rel_err(ctx.download(xsb.F_FX)[own], ref0[0][own]) < TOL64


def test_types_uploaded_after_the_list_build_use_the_byte_copy_stage(tmp_path):
    """chunk_neighbors aligns the stage rows for TMA-copied type bytes only when it sees a multi-species system; if the
    types arrive later (list built on an all-zero type array), the multi-element passes must still stage them correctly"""
    O = oracle()
    path = write_setfl(str(tmp_path / "ab.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    gs = system(ncells=6, a=3.615, sigma=0.05, cell=3.615 * 2, gl=2, types=[0, 1, 1, 0], seed=14)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 7.0, 1, True)
    ref = [gs.zeros() for _ in range(5)]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, O.EamAlloy(path), 6.0, 1 | 2 | 4 | 8 | 16, ref[0], ref[1], ref[2], ref[3], None, ref[4])
    ctx = make_ctx(gs)
    ctx.upload(xsb.F_TYPE, np.zeros_like(gs.type))
    ctx.eam_alloy_load(path); ctx.chunk_neighbors(7.0)            # single-species as far as the build can tell
    ctx.upload(xsb.F_TYPE, gs.type)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG)
    ctx.eam_alloy_force(6.0, xsb.EAM_FORCE | xsb.EAM_EFLAG)
    own = ~gs.is_ghost
    for f, r in ((xsb.F_FX, ref[0]), (xsb.F_FY, ref[1]), (xsb.F_FZ, ref[2]), (xsb.F_EP, ref[3])):
        assert rel_err(ctx.download(f)[own], r[own]) < TOL64


@pytest.mark.parametrize("triclinic", [False, True])
def test_verlet_boundary_equals_the_five_separate_operators(triclinic):
    """xsb_verlet_boundary = force_to_accel, push_f_v(dt/2) | push_f_v_r(dt), push_f_v(dt/2), particle_displ_over in one pass"""
    pos, typ, box = lattice("FCC", 6, 5.0, 0.1, seed=3, types=[0, 1, 0, 1])
    rng = np.random.default_rng(5)
    vel = rng.normal(0.0, 2.0, pos.shape)
    X = np.array([[1.02, 0.03, 0.01], [0.0, 0.98, 0.02], [0.0, 0.0, 1.01]]) if triclinic else None
    masses, dt = [39.948, 83.798], 2.0e-3
    out = []
    for fused in (False, True):
        ctx = assigned_ctx(pos, typ, box, 10.0, 1, vel)
        if X is not None:
            ctx.grid_set_xform(X)
        ctx.backup_r()
        n = ctx.n
        f = np.random.default_rng(9).normal(0.0, 50.0, (3, n))
        for k, fld in enumerate((xsb.F_FX, xsb.F_FY, xsb.F_FZ)):
            ctx.upload(fld, f[k])
        if fused:
            over, d = ctx.verlet_boundary(masses, dt, 0.004)
        else:
            ctx.force_to_accel(masses); ctx.push_f_v(0.5 * dt)
            ctx.push_f_v_r(dt); ctx.push_f_v(0.5 * dt)
            over, d = ctx.particle_displ_over(0.004)
        out.append((over, d, [ctx.download(x) for x in (xsb.F_RX, xsb.F_RY, xsb.F_RZ, xsb.F_VX, xsb.F_VY, xsb.F_VZ, xsb.F_FX, xsb.F_FY, xsb.F_FZ)]))
    (o0, d0, a0), (o1, d1, a1) = out
    assert o0 == o1 and abs(d0 - d1) <= 1e-15 * max(d0, 1e-300) + 1e-18 and d0 > 0
    for u, v in zip(a0, a1):
        assert np.abs(u - v).max() <= 4e-16 * max(np.abs(u).max(), 1e-300)


@pytest.mark.parametrize("two_species,mixed,drift", [(False, False, False), (True, False, False), (False, True, False), (True, False, True)])
def test_eam_inner_skin_reuse_matches_plain_path_and_oracle(tmp_path, two_species, mixed, drift):
    """xsb_eam_inner_skin: the rho phase re-evaluates the sub-list of an earlier step instead of re-filtering the neighbour
    list while the device-side displacement budget holds.  An NVE run with the skin must give the plain path's forces at
    every step (same pairs; summation order differs), really reuse the list, re-filter when an atom jumps, and match the
    oracle on the final positions."""
    O = oracle()
    path = str(tmp_path / "s.eam.alloy")
    els = [SC_CU, SC_XX] if two_species else [SC_CU]
    write_setfl(path, els, nrho=2000, drho=0.1, nr=2000, rc=6.0)
    pos, typ, box = lattice("FCC", 6, 3.615, 0.05, seed=21, types=[0, 1, 1, 0] if two_species else None)
    vel = np.random.default_rng(3).normal(0.0, 3.0, pos.shape)
    masses, dt, rcut, nbh = [63.5, 27.0], 1.0e-3, 6.0, 7.0
    fl = xsb.FLAG_MIXED if mixed else 0
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    ctxs = []
    for skin in (0.0, 0.2):
        c = assigned_ctx(pos, typ, box, 3.615 * 2, 1, vel)
        c.eam_alloy_load(path); c.eam_inner_skin(skin)
        c.chunk_neighbors(nbh); c.backup_r()
        ctxs.append(c)

    def forces(c):
        c.zero_force_energy()
        c.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_EFLAG, fl)
        c.ghost_update([xsb.F_RHO_DEMB])
        c.eam_alloy_force(rcut, xsb.EAM_FORCE | xsb.EAM_EFLAG, fl)
        return [c.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]

    tol = 1e-5 if mixed else 1e-12
    X0 = np.array([[1.0, 0.01, 0.005], [0.0, 1.0, 0.01], [0.0, 0.0, 1.0]])
    for step in range(7):
        if drift:                                          # NPT-like cell matrix: changes every step, accounted in the budget
            for c in ctxs:
                c.grid_set_xform(X0 * (1.0 + 2e-6 * step))
        out = [forces(c) for c in ctxs]
        for a, b in zip(*out):
            assert rel_err(a, b) < tol, "step %d" % step
        if step == 4:
            # every atom drifts 0.3 ang per step from now on (more than half the inner skin) through the accounted path:
            # the following rho phases must re-filter (relative positions, hence forces, are unaffected by the drift)
            for c in ctxs:
                c.upload(xsb.F_VX, c.download(xsb.F_VX) + 300.0)
        for c in ctxs:
            c.verlet_boundary_async(masses, dt)
            c.ghost_update(POS)
    built, reused = ctxs[1].eam_sublist_stats()
    print("inner skin: %d re-filtered, %d reused" % (built, reused))
    assert reused >= 3 and built >= 2
    assert ctxs[0].eam_sublist_stats() == (0, 0)
    # final positions against the oracle (ghost images included: one ghost layer, rho_dEmb copied owner -> ghost)
    c = ctxs[1]
    if drift:
        c.grid_set_xform(np.eye(3))                        # a real deformation: exhausts the budget, the pass re-filters
    out = forces(c)
    off = c.cell_offsets()
    rx, ry, rz, tt = c.download(xsb.F_RX), c.download(xsb.F_RY), c.download(xsb.F_RZ), c.download(xsb.F_TYPE)
    dims = np.array(c.grid.dims[:]); cell = c.grid.cell_size
    g = O.make_grid(dims, 1, cell, [-cell] * 3)
    nb = O.Neighbors.build(g, off, rx, ry, rz, nbh, 1, True)
    n = len(rx)
    rfx, rfy, rfz, rep, emb = [np.zeros(n) for _ in range(5)]
    eam = O.EamAlloy(path)
    O.eam_alloy(g, off, rx, ry, rz, tt, nb, eam, rcut, 1 | 2 | 16, rfx, rfy, rfz, rep, None, emb)
    demb_gpu = c.download(xsb.F_RHO_DEMB)                        # already ghost-updated by forces()
    cells = np.arange(len(off) - 1); ci = cells % dims[0]; cj = (cells // dims[0]) % dims[1]; ck = cells // (dims[0] * dims[1])
    own_cell = (ci >= 1) & (ci < dims[0] - 1) & (cj >= 1) & (cj < dims[1] - 1) & (ck >= 1) & (ck < dims[2] - 1)
    own = np.repeat(own_cell, np.diff(off).astype(np.int64))
    assert rel_err(demb_gpu[own], emb[own]) < (1e-5 if mixed else TOL64)
    O.eam_alloy(g, off, rx, ry, rz, tt, nb, eam, rcut, 8 | 16, rfx, rfy, rfz, rep, None, demb_gpu.copy())
    for a, b in zip(out, (rfx, rfy, rfz, rep)):
        assert rel_err(a[own], b[own]) < (1e-5 if mixed else TOL64)


def test_verlet_boundary_async_ring_and_own_atom_transfers():
    """xsb_verlet_boundary_async + xsb_displ_poll (no host read-back) give the blocking variant's maxima, one call late on
    request; xsb_fields_upload_async / _download_async move the own atoms only and keep stream order"""
    pos, typ, box = lattice("FCC", 6, 5.0, 0.1, seed=3)
    vel = np.random.default_rng(5).normal(0.0, 2.0, pos.shape)
    masses, dt = [39.948], 2.0e-3
    a = assigned_ctx(pos, typ, box, 10.0, 1, vel); b = assigned_ctx(pos, typ, box, 10.0, 1, vel)
    for c in (a, b):
        c.backup_r()
    hist = []
    for step in range(3):
        over, d = a.verlet_boundary(masses, dt, 1.0)
        b.verlet_boundary_async(masses, dt)
        d0, s0 = b.displ_poll(0)
        assert abs(d0 - d) <= 1e-15 * d and 0 < s0 <= d0 * (1 + 1e-12)
        hist.append((d0, s0))
        if step:
            assert b.displ_poll(1) == hist[-2]
    for f in (xsb.F_RX, xsb.F_VZ):
        assert np.array_equal(a.download(f), b.download(f))
    with pytest.raises(xsb.XsbError):
        b.displ_poll(5)                     # nothing recorded that far back
    # own-atom transfers: download -> modify -> upload -> whole-array download shows the change on own atoms only
    import torch
    n_own, n = b.n_own, b.n
    pin = [torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(2)]
    before = b.download(xsb.F_RX)
    b.fields_download_async([xsb.F_RX, xsb.F_RY], [t.data_ptr() for t in pin]); b.copy_wait()
    off = b.cell_offsets(); dims = np.array(b.grid.dims[:])
    cells = np.arange(len(off) - 1); ci = cells % dims[0]; cj = (cells // dims[0]) % dims[1]; ck = cells // (dims[0] * dims[1])
    own_cell = (ci >= 1) & (ci < dims[0] - 1) & (cj >= 1) & (cj < dims[1] - 1) & (ck >= 1) & (ck < dims[2] - 1)
    own = np.repeat(own_cell, np.diff(off).astype(np.int64))
    assert own.sum() == n_own and np.array_equal(pin[0].numpy(), before[own])
    pin[0] += 0.125
    b.fields_upload_async([xsb.F_RX], [pin[0].data_ptr()])
    after = b.download(xsb.F_RX)
    assert np.array_equal(after[own], before[own] + 0.125) and np.array_equal(after[~own], before[~own])
    b.copy_wait()


@pytest.mark.parametrize("two_species,virial", [(False, False), (True, True)])
def test_eam_alloy_mixed_precision_parity(tmp_path, two_species, virial):
    """XSB_FLAG_MIXED on eam_alloy_force: FP32 spline + pair math (FP64 distances and accumulation) against the FP64 oracle
    at the mixed-mode bar (1e-5 of the field maximum); two-call pattern, so the force pass consumes the FP32 pair cache"""
    O = oracle()
    els = [SC_CU, SC_XX] if two_species else [SC_CU]
    path = write_setfl(str(tmp_path / "m.eam.alloy"), els, nrho=10000, drho=0.02, nr=5000, rc=6.0)
    gs = system(ncells=6, a=3.615, sigma=0.08, cell=3.615 * 2, gl=2, types=[0, 1, 1, 0] if two_species else None, seed=17)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 7.0, 1, True)
    ref = [gs.zeros() for _ in range(5)]
    rvir = np.zeros((gs.n, 9)) if virial else None
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, O.EamAlloy(path), 6.0, 1 | 2 | 4 | 8 | 16 | (32 if virial else 0),
                ref[0], ref[1], ref[2], ref[3], rvir, ref[4])
    ctx = make_ctx(gs); ctx.eam_alloy_load(path); ctx.chunk_neighbors(7.0)
    ctx.zero_force_energy(ghost=True)
    fl = xsb.FLAG_MIXED | (xsb.FLAG_VIRIAL if virial else 0)
    ctx.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG, fl)
    ctx.eam_alloy_force(6.0, xsb.EAM_FORCE | xsb.EAM_EFLAG, fl)
    own = ~gs.is_ghost
    assert rel_err(ctx.download(xsb.F_RHO_DEMB), ref[4]) < TOLMIX
    errs = [rel_err(ctx.download(f)[own], r[own]) for f, r in ((xsb.F_FX, ref[0]), (xsb.F_FY, ref[1]), (xsb.F_FZ, ref[2]), (xsb.F_EP, ref[3]))]
    assert max(errs) < TOLMIX, errs
    assert max(errs) > 1e-12, "mixed mode produced FP64-exact results: the FP32 path did not run"
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL).reshape(-1, 9)[own], rvir[own]) < TOLMIX


def test_recorded_step_replays_the_direct_calls_bit_for_bit(tmp_path):
    """xsb_step_capture_begin/_end + xsb_step_replay: the integrator pass, ghost update, zero and the force operators of a
    regular step recorded once and re-issued with one launch leave exactly the arrays the direct calls leave (LJ and a
    two-pass eam_alloy_force), xsb_displ_poll sees the same maxima, and a replay after a new list build is refused"""
    pos, typ, box = lattice("FCC", 6, 5.0, 0.1, seed=3)
    vel = np.random.default_rng(5).normal(0.0, 2.0, pos.shape)
    masses, dt = [39.948], 2.0e-3
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    path = write_setfl(str(tmp_path / "g.eam.alloy"), [SC_CU], nrho=2000, drho=0.05, nr=2000, rc=6.0)
    for kind in ("lj", "eam", "chain"):         # chain: an LJ operator behind the EAM one joins its (deferred) force pass
        a = assigned_ctx(pos, typ, box, 10.0, 1, vel); b = assigned_ctx(pos, typ, box, 10.0, 1, vel)

        def forces(c):
            c.zero_force_energy(ghost=True)
            if kind == "lj":
                c.pair_force([0.0104 * EV, 3.4], 8.0, xsb.FLAG_ENERGY)
            else:
                c.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_GHOST | xsb.EAM_EFLAG, 0)
                c.eam_alloy_force(6.0, xsb.EAM_FORCE | xsb.EAM_EFLAG, 0)
                if kind == "chain":
                    c.pair_force([0.0104 * EV, 3.4], 5.5, xsb.FLAG_ENERGY)

        for c in (a, b):
            if kind != "lj":
                c.eam_alloy_load(path)
            c.chunk_neighbors(9.0); c.backup_r(); forces(c)
        l0 = b.launches
        b.step_capture_begin()
        b.verlet_boundary_async(masses, dt); b.ghost_update(POS); forces(b)
        sid = b.step_capture_end()
        assert b.launches == l0                      # recorded, not executed
        for step in range(4):
            a.verlet_boundary_async(masses, dt); a.ghost_update(POS); forces(a)
            b.step_replay(sid)
            assert a.displ_poll(0) == b.displ_poll(0)
        assert b.launches > l0
        for f in (xsb.F_RX, xsb.F_RY, xsb.F_RZ, xsb.F_VX, xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP):
            assert np.array_equal(a.download(f), b.download(f)), (kind, f)
        # a blocking entry point must not be recorded
        b.step_capture_begin()
        with pytest.raises(xsb.XsbError):
            b.verlet_boundary(masses, dt, 1.0)
        with pytest.raises(xsb.XsbError):
            b.step_capture_end()
        # the stream still works, and the old recording is refused once the list was rebuilt
        b.sync(); b.chunk_neighbors(9.0)
        with pytest.raises(xsb.XsbError):
            b.step_replay(sid)
        b.step_release(sid)


def test_eam_inner_skin_with_host_driven_positions(tmp_path):
    """positions of the own atoms uploaded every step (the plugin use case: the host integrates) are charged to the inner
    skin's displacement budget by xsb_fields_upload_async itself: small moves let the rho phase reuse its sub-list, a jump
    forces a re-filter, and the forces equal those of the plain path at every step"""
    import torch
    path = str(tmp_path / "h.eam.alloy")
    write_setfl(path, [SC_CU], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    pos, typ, box = lattice("FCC", 6, 3.615, 0.05, seed=23)
    rcut, nbh = 6.0, 7.0
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    ctxs = []
    for skin in (0.0, 0.2):
        c = assigned_ctx(pos, typ, box, 3.615 * 2, 1)
        c.eam_alloy_load(path); c.eam_inner_skin(skin); c.chunk_neighbors(nbh)
        ctxs.append(c)
    n_own = ctxs[0].n_own
    pin = [torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(3)]
    ctxs[0].fields_download_async(POS, [t.data_ptr() for t in pin]); ctxs[0].copy_wait()
    rng = np.random.default_rng(9)

    def forces(c):
        c.ghost_update(POS)
        c.zero_force_energy()
        c.eam_alloy_force(rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_EFLAG, 0)
        c.ghost_update([xsb.F_RHO_DEMB])
        c.eam_alloy_force(rcut, xsb.EAM_FORCE | xsb.EAM_EFLAG, 0)
        return [c.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)]

    stats = []
    for step in range(8):
        for c in ctxs:
            c.fields_upload_async(POS, [t.data_ptr() for t in pin])
        out = [forces(c) for c in ctxs]
        for a, b in zip(*out):
            assert rel_err(a, b) < 1e-12, "step %d" % step
        stats.append(ctxs[1].eam_sublist_stats())
        for c in ctxs:
            c.copy_wait()
        for t in pin:                                   # the host's integrator: 0.004 ang of noise per step, one 0.3 ang jump
            t += torch.from_numpy(rng.normal(0.0, 0.004, n_own)) + (0.3 if step == 4 else 0.0)
    built, reused = stats[-1]
    print("host-driven inner skin: per step (re-filtered, reused) =", stats)
    assert reused >= 4 and built >= 2
    assert stats[5][0] == stats[4][0] + 1             # the jump after step 4 made step 5 re-filter
    assert ctxs[0].eam_sublist_stats() == (0, 0)


def test_pair_operator_behind_eam_walks_the_sublist(tmp_path, monkeypatch):
    """compute_force: [eam_alloy_force, lj_multi_force] (configs[4]): the pair operator reuses the in-range sub-list the EAM
    operator left when its own cut-off is not larger; XSB_PAIR_NO_SUBLIST=1 keeps the full list.  Same forces either way (only
    the summation order differs), with a shorter pair cut-off, the inner skin on, and after the atoms moved (re-evaluated list)"""
    path = write_setfl(str(tmp_path / "ab.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    rng = np.random.default_rng(5)
    pos, typ, box = lattice("FCC", 6, 3.615, 0.06, seed=9)
    typ = (rng.random(len(pos)) < 0.5).astype(np.uint8)
    vel = rng.normal(0.0, 3.0, pos.shape)
    rows = np.array([[0.0104 * EV, 2.3, 5.6], [0.0150 * EV, 2.2, 5.0], [0.0200 * EV, 2.1, 4.4]])
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    ctxs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("XSB_PAIR_NO_SUBLIST", "1")
        c = assigned_ctx(pos, typ, box, 3.615 * 2, 1, vel)
        c.eam_alloy_load(path); c.eam_inner_skin(0.2); c.chunk_neighbors(7.0); c.backup_r()
        ctxs.append(c)
    monkeypatch.delenv("XSB_PAIR_NO_SUBLIST")
    for step in range(3):
        outs = []
        for c in ctxs:
            c.zero_force_energy()
            c.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_EFLAG, 0)
            c.ghost_update([xsb.F_RHO_DEMB])
            c.eam_alloy_force(6.0, xsb.EAM_FORCE | xsb.EAM_EFLAG, 0)
            eam_only = c.download(xsb.F_FX).copy()
            c.pair_multi_force(2, rows, 5.6, xsb.FLAG_ENERGY)
            outs.append([c.download(f) for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)])
            assert np.abs(outs[-1][0] - eam_only).max() > 0            # the pair operator did add something
        for a, b in zip(*outs):
            assert rel_err(a, b) < 1e-12, step
        for c in ctxs:
            c.verlet_boundary_async([63.5, 27.0], 1.0e-3); c.ghost_update(POS)
    assert ctxs[0].eam_sublist_stats()[1] >= 1                         # a re-evaluated (not re-filtered) sub-list was walked too


@pytest.mark.parametrize("two_species,eflag,virial,mixed,triclinic", [(True, False, False, False, True), (True, True, False, False, False),
                                                                     (True, True, True, False, True), (False, True, False, False, False),
                                                                     (True, True, True, True, False), (False, False, False, True, True)])
def test_lj_operator_chained_behind_eam_joins_its_force_pass(tmp_path, monkeypatch, two_species, eflag, virial, mixed, triclinic):
    """compute_force: [eam_alloy_force, lj_multi_force] (configs[4]): the EAM force phase waits for the next entry point and
    takes a Lennard-Jones operator along (one pass over the in-range pairs).  Same forces / energies / virial as with
    XSB_NO_CHAIN_FUSION=1 (two passes) over several steps incl. re-evaluated sub-lists; anything that is not such an operator
    (another potential, a larger cut-off, a ghost flag, different flags, any other entry point) is served unfused"""
    els = [SC_CU, SC_XX] if two_species else [SC_CU]
    path = write_setfl(str(tmp_path / "ab.eam.alloy"), els, nrho=2000, drho=0.1, nr=2000, rc=6.0)
    rng = np.random.default_rng(15)
    pos, typ, box = lattice("FCC", 6, 3.615, 0.06, seed=19)
    if two_species:
        typ = (rng.random(len(pos)) < 0.5).astype(np.uint8)
    vel = rng.normal(0.0, 3.0, pos.shape)
    rows = np.array([[0.0104 * EV, 2.3, 5.6], [0.0150 * EV, 2.2, 5.0], [0.0200 * EV, 2.1, 4.4]])
    masses = [63.5, 27.0] if two_species else [63.5]
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    X = np.array([[1.015, 0.02, -0.01], [0.0, 0.99, 0.015], [0.0, 0.0, 1.005]]) if triclinic else None
    ef = xsb.EAM_EFLAG if eflag else 0
    fl = (xsb.FLAG_VIRIAL if virial else 0) | (xsb.FLAG_MIXED if mixed else 0)
    pfl = fl | (xsb.FLAG_ENERGY if eflag else 0)
    ctxs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("XSB_NO_CHAIN_FUSION", "1")
        c = assigned_ctx(pos, typ, box, 3.615 * 2, 1, vel)
        if X is not None:
            c.grid_set_xform(X)
        c.eam_alloy_load(path); c.eam_inner_skin(0.2); c.chunk_neighbors(7.0); c.backup_r()
        ctxs.append(c)
    monkeypatch.delenv("XSB_NO_CHAIN_FUSION")

    def lj(c):
        if two_species:
            c.pair_multi_force(2, rows, 5.6, pfl)
        else:
            c.pair_force([0.0104 * EV, 2.3], 5.6, pfl)

    fields = [xsb.F_FX, xsb.F_FY, xsb.F_FZ] + ([xsb.F_EP] if eflag else []) + ([xsb.F_VIRIAL] if virial else [])
    tol = 2e-6 if mixed else 1e-12
    for step in range(3):
        outs = []
        for c in ctxs:
            c.zero_force_energy()
            c.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | ef, fl)
            c.ghost_update([xsb.F_RHO_DEMB])
            c.eam_alloy_force(6.0, xsb.EAM_FORCE | ef, fl)
            lj(c)
            outs.append([c.download(f) for f in fields])
        for a, b in zip(*outs):
            assert rel_err(a, b) < tol, step
        for c in ctxs:
            c.verlet_boundary_async(masses, 1.0e-3); c.ghost_update(POS)
    assert ctxs[0].chain_stats() == 3 and ctxs[1].chain_stats() == 0
    assert ctxs[0].eam_sublist_stats()[1] >= 1
    # not absorbed: the results must still be those of the unfused context
    a, b = ctxs
    cases = [lambda c: c.pair_force([3.0e-3 * EV, 2.0, 1.0e-4 * EV], 5.0, pfl, pot=xsb.POT_BUCKINGHAM),
             lambda c: (c.pair_multi_force(2, rows + np.array([0, 0, 1.0]), 6.6, pfl) if two_species else c.pair_force([0.0104 * EV, 2.3], 6.6, pfl)),
             lambda c: (c.pair_multi_force(2, rows, 5.6, pfl ^ xsb.FLAG_VIRIAL) if eflag else c.pair_force([0.0104 * EV, 2.3], 5.6, pfl | xsb.FLAG_ENERGY))]
    n0 = a.chain_stats()
    for case in cases:
        outs = []
        for c in (a, b):
            c.zero_force_energy()
            c.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | ef, fl)
            c.ghost_update([xsb.F_RHO_DEMB])
            c.eam_alloy_force(6.0, xsb.EAM_FORCE | ef, fl)
            case(c)
            outs.append([c.download(f) for f in fields])
        for u, v in zip(*outs):
            assert rel_err(u, v) < tol
    assert a.chain_stats() == n0
    # the deferred phase alone: any other entry point launches it
    for c in (a, b):
        c.zero_force_energy()
        c.eam_alloy_force(6.0, xsb.EAM_RHO | xsb.EAM_RHO2EMB | ef, fl)
        c.ghost_update([xsb.F_RHO_DEMB])
        c.eam_alloy_force(6.0, xsb.EAM_FORCE | ef, fl)
    assert rel_err(a.download(xsb.F_FX), b.download(xsb.F_FX)) < tol


@pytest.mark.parametrize("name", ["johnson"] + sorted(EAM1_CASES))
@pytest.mark.parametrize("virial,two_step", [(False, True), (True, False)])
def test_eam_analytic_models_mixed_precision(name, virial, two_step):
    """XSB_FLAG_MIXED on johnson / sutton_chen / vniitf (FP32 rho(r), phi(r); FP64 F(rho), distances, sums): 1e-5 against the
    FP64 oracle, really computed in FP32 (differs from the FP64 pass), and an FP64 force pass never consumes the rho'(r) an
    FP32 emb pass cached"""
    model, p, rcut, nbh = (xsb.EAM_JOHNSON, johnson_params(), 5.5, 6.5) if name == "johnson" else EAM1_CASES[name]
    O = oracle()
    gs = system(ncells=5, a=3.615, sigma=0.05, cell=3.615, gl=4)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    fx, fy, fz, ep, emb = gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros(), gs.zeros()
    vir = np.zeros((gs.n, 9)) if virial else None
    O.eam_analytic(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, model, p, rcut, 7, fx, fy, fz, ep, vir, emb)
    ctx = make_ctx(gs)
    ctx.chunk_neighbors(nbh)
    fl = (xsb.FLAG_VIRIAL if virial else 0) | xsb.FLAG_MIXED
    ctx.zero_force_energy(ghost=True)
    if two_step:
        ctx.eam_analytic_force(model, p, rcut, 3, fl)
        ctx.eam_analytic_force(model, p, rcut, 4, fl)
    else:
        ctx.eam_analytic_force(model, p, rcut, 7, fl)
    got = {f: ctx.download(f).copy() for f in (xsb.F_RHO_DEMB, xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP)}
    errs = [rel_err(got[f], ref) for f, ref in ((xsb.F_RHO_DEMB, emb), (xsb.F_FX, fx), (xsb.F_FY, fy), (xsb.F_FZ, fz), (xsb.F_EP, ep))]
    print("%s mixed: max rel err rho_dEmb,fx,fy,fz,ep = %s" % (name, ["%.2e" % e for e in errs]))
    assert max(errs) < TOLMIX
    assert max(errs) > 1e-9                                            # FP32 arithmetic really ran
    if virial:
        assert rel_err(ctx.download(xsb.F_VIRIAL), vir) < TOLMIX
    # FP32 emb pass, then an FP64 force pass on the oracle's F'(rho): the pass must re-evaluate rho'(r) itself instead of
    # taking the FP32 values the emb pass cached -- then its forces are the oracle's to 1e-10
    ctx.zero_force_energy(ghost=True)
    ctx.eam_analytic_force(model, p, rcut, 3, xsb.FLAG_MIXED)
    ctx.upload(xsb.F_RHO_DEMB, emb)
    ctx.eam_analytic_force(model, p, rcut, 4, 0)
    assert rel_err(ctx.download(xsb.F_FX), fx) < TOL64
