#!/usr/bin/env python
"""Multi-GPU parity check, launched as
     python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
One rank per GPU; the periodic FCC system is cut into N bricks (xsb_domain_desc), every rank runs ghost_comm_scheme,
chunk_neighbors and the force operators on its brick with NCCL ghost exchanges, rank 0 gathers the owned atoms by id and
compares with the CPU oracle run on the WHOLE system (tolerance 1e-10 of the field maximum, FP64 mode).
Covers SURVEY.md 8(e): ghost_update_r / ghost_update_opt(rho_dEmb) / update_force_energy_from_ghost over NCCL P2P."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

import exastamp_b200 as xsb       # noqa: E402
from helpers import EV, SC_CU, SC_XX, GridSystem, lattice, write_setfl  # noqa: E402

TOL = 1e-10


def rank_dims(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def own_mask(ctx, dims, gl):
    off = ctx.cell_offsets().astype(np.int64)
    nx, ny, nz = dims
    c = np.arange(nx * ny * nz)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    ghost_cell = (i < gl) | (i >= nx - gl) | (j < gl) | (j >= ny - gl) | (k < gl) | (k >= nz - gl)
    return ~np.repeat(ghost_cell, np.diff(off))


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rd = np.array(rank_dims(world))
    coord = np.array([rank % rd[0], (rank // rd[0]) % rd[1], rank // (rd[0] * rd[1])])
    a = 3.615
    ncb = 3                                     # own cells per brick axis
    uc_brick = 7                                # FCC unit cells per brick axis -> brick 25.3 ang, cell 8.435 ang
    pos, typ, box = lattice("FCC", [int(v) for v in uc_brick * rd], a, 0.08, seed=5, types=[0, 1, 0, 0])
    brick = box / rd
    cell = brick[0] / ncb
    owner = np.minimum(np.floor(pos / brick).astype(np.int64), rd - 1)
    mine = np.all(owner == coord, axis=1)
    ids = np.arange(len(pos), dtype=np.uint64)
    rng = np.random.default_rng(3)
    vel = rng.normal(0, 1.0, pos.shape)

    ctx = xsb.Context(local)
    uid = [xsb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    dims = [ncb + 2] * 3
    ctx.grid_set(xsb.make_grid(dims, 1, cell, [(c * ncb - 1) * cell for c in coord]))
    p = pos[mine]
    ctx.particles_assign(p[:, 0], p[:, 1], p[:, 2], vel[mine, 0], vel[mine, 1], vel[mine, 2], typ[mine], ids[mine])
    ctx.set_domain([int(v) for v in ncb * rd], (1, 1, 1), [int(v) for v in rd], [int(v) for v in coord])
    ctx.ghost_comm_scheme()
    tmp = tempfile.mkdtemp()
    setfl = write_setfl(os.path.join(tmp, "cuxx.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.5)
    ctx.eam_alloy_load(setfl)
    rc_eam, rc_lj, skin = 6.5, 7.0, 0.5
    nbh = 7.5
    assert cell >= nbh
    ctx.chunk_neighbors(nbh)
    own = own_mask(ctx, dims, 1)
    assert own.sum() == ctx.n_own == mine.sum()

    results = {}

    def collect(tag, fields):
        got = {"id": ctx.download(xsb.F_ID)[own]}
        for name, f in fields.items():
            got[name] = ctx.download(f)[own]
        results[tag] = got

    # --- eam_alloy_force in three phases with the rho_dEmb ghost exchange between them (decks: compute_force_nosym)
    ctx.zero_force_energy(ghost=True)
    ctx.eam_alloy_force(rc_eam, xsb.EAM_RHO | xsb.EAM_RHO2EMB | xsb.EAM_EFLAG)
    ctx.ghost_update([xsb.F_RHO_DEMB])
    ctx.eam_alloy_force(rc_eam, xsb.EAM_FORCE | xsb.EAM_EFLAG)
    collect("eam", {"fx": xsb.F_FX, "fy": xsb.F_FY, "fz": xsb.F_FZ, "ep": xsb.F_EP})
    # --- lj_compute_force chained on top (accumulates)
    ctx.pair_force([0.0104 * EV, 3.4], rc_lj)
    collect("eam+lj", {"fx": xsb.F_FX, "fy": xsb.F_FY, "fz": xsb.F_FZ, "ep": xsb.F_EP})
    # --- snap_force: Newton-on, forces of ghost neighbours travel back with update_force_energy_from_ghost
    twoj = 4
    ncoef = xsb.load_library().xsb_snap_ncoeff(twoj)
    beta = np.random.default_rng(1).normal(0, 1, (2, ncoef + 1)) * 1e-2 * EV
    rad, wj, rcutfac = [0.5, 0.45], [1.0, 0.8], 5.0
    ctx.snap_set(twoj, rcutfac, rad, wj, beta)
    ctx.zero_force_energy(ghost=True)
    ctx.snap_force(xsb.FLAG_ENERGY)
    ctx.ghost_reduce_add([xsb.F_FX, xsb.F_FY, xsb.F_FZ])
    collect("snap", {"fx": xsb.F_FX, "fy": xsb.F_FY, "fz": xsb.F_FZ, "ep": xsb.F_EP})
    # --- ghost_update_r after a move of the owners + the displacement allreduce
    d = np.random.default_rng(100).normal(0, 0.02, pos.shape)
    newpos = pos + d
    idl = ctx.download(xsb.F_ID).astype(np.int64)
    for k, f in enumerate((xsb.F_RX, xsb.F_RY, xsb.F_RZ)):
        cur = ctx.download(f)
        cur[own] = cur[own] + d[idl[own], k]
        ctx.backup_r() if k == 0 else None
        ctx.upload(f, cur)
    over, dmax = ctx.particle_displ_over(1e-3)
    ctx.ghost_update([xsb.F_RX, xsb.F_RY, xsb.F_RZ])
    gx, gy, gz = ctx.download(xsb.F_RX), ctx.download(xsb.F_RY), ctx.download(xsb.F_RZ)
    gpos = np.stack([gx, gy, gz], axis=1)
    # every particle (own or ghost) must sit at newpos[id] modulo the box
    delta = gpos - newpos[idl]
    wrap = delta - np.round(delta / box) * box
    assert np.max(np.abs(wrap)) < 1e-9, "ghost_update_r: ghost positions do not follow their owners"
    assert over, "particle_displ_over must fire"
    ctx.zero_force_energy(ghost=True)
    ctx.pair_force([0.0104 * EV, 3.4], rc_lj)
    collect("lj_moved", {"fx": xsb.F_FX, "ep": xsb.F_EP})
    results["dmax"] = dmax
    results["transport"] = ctx.ghost_transport()
    # --- move_particles + migrate_cell_particles: a large random move sends many particles to other bricks (and across
    #     the periodic boundary); after rebin + ghost scheme + list rebuild the forces must match a fresh oracle run
    big = np.random.default_rng(200).uniform(-4.0, 4.0, pos.shape)
    big[::7] += np.array([box[0] * 0.5, 0.0, 0.0])            # some travel half the box: arbitrary-distance migration
    for k, f in enumerate((xsb.F_RX, xsb.F_RY, xsb.F_RZ)):
        cur = ctx.download(f)
        cur[own] = cur[own] + big[idl[own], k]
        ctx.upload(f, cur)
    n_before = ctx.n_own
    ctx.particles_rebin()
    sent, received = ctx.migration_stats()
    ctx.ghost_comm_scheme()
    ctx.chunk_neighbors(nbh)
    own2 = own_mask(ctx, dims, 1)
    assert own2.sum() == ctx.n_own == n_before - sent + received
    ctx.zero_force_energy(ghost=True)
    ctx.pair_force([0.0104 * EV, 5.5], rc_lj)                 # soft, wide LJ: random overlaps stay finite-ish
    results["migrated"] = {"id": ctx.download(xsb.F_ID)[own2], "fx": ctx.download(xsb.F_FX)[own2], "ep": ctx.download(xsb.F_EP)[own2],
                           "vx": ctx.download(xsb.F_VX)[own2], "type": ctx.download(xsb.F_TYPE)[own2], "rx": ctx.download(xsb.F_RX)[own2]}
    results["migration"] = (sent, received, n_before, ctx.n_own)

    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank != 0:
        return

    # ---------------- oracle on the whole periodic system
    from oracle import oracle as O
    ncg = [int(v) for v in ncb * rd]
    gs = GridSystem(pos, typ, box, cell, 1)
    assert list(gs.n_own) == ncg
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh, 1, True)
    ownm = ~gs.is_ghost
    src = gs.src_index

    def per_atom(arr):
        out = np.zeros(len(pos)); out[src[ownm]] = arr[ownm]; return out

    def merged(tag, name):
        out = np.full(len(pos), np.nan)
        for r in gathered:
            out[r[tag]["id"].astype(np.int64)] = r[tag][name]
        assert not np.isnan(out).any(), "some atoms were not owned by any rank"
        return out

    def check(tag, name, ref):
        got = merged(tag, name)
        err = np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)
        print("  %-9s %-3s rel err %.2e" % (tag, name, err))
        assert err < TOL, (tag, name, err)

    fx, fy, fz, ep, emb = [gs.zeros() for _ in range(5)]
    # same three phases as the ranks ran: rho + rho2emb on owned atoms, owner -> ghost copy of rho_dEmb, force
    eam = O.EamAlloy(setfl)
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, rc_eam, 1 | 2 | 16, fx, fy, fz, ep, None, emb)
    owner_of = np.zeros(len(pos), dtype=np.int64); owner_of[src[ownm]] = np.nonzero(ownm)[0]
    emb[:] = emb[owner_of[src]]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, rc_eam, 8 | 16, fx, fy, fz, ep, None, emb)
    for n_, a_ in (("fx", fx), ("fy", fy), ("fz", fz), ("ep", ep)):
        check("eam", n_, per_atom(a_))
    O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, [0.0104 * EV, 3.4], rc_lj, 0, fx, fy, fz, ep, None)
    for n_, a_ in (("fx", fx), ("fy", fy), ("fz", fz), ("ep", ep)):
        check("eam+lj", n_, per_atom(a_))
    S = O.Snap(twoj, rcutfac, rad, wj, beta)
    sfx, sfy, sfz, sep = [gs.zeros() for _ in range(4)]
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, sfx, sfy, sfz, sep, None)
    # Newton-on: fold the ghost images' forces back on their owners
    tot = [np.zeros(len(pos)) for _ in range(3)]
    for t, arr in zip(tot, (sfx, sfy, sfz)):
        np.add.at(t, src, arr)
    for n_, a_ in (("fx", tot[0]), ("fy", tot[1]), ("fz", tot[2]), ("ep", per_atom(sep))):
        check("snap", n_, a_)
    gs2 = GridSystem(np.mod(newpos, box), typ, box, cell, 1)
    # moved system: rank-local cells were not re-binned, so compare through a fresh oracle build on the new positions
    g2 = gs2.oracle_grid()
    nb2 = O.Neighbors.build(g2, gs2.cell_off, gs2.rx, gs2.ry, gs2.rz, nbh, 1, True)
    f2, e2 = gs2.zeros(), gs2.zeros()
    O.pair_force(g2, gs2.cell_off, gs2.rx, gs2.ry, gs2.rz, nb2, [0.0104 * EV, 3.4], rc_lj, 0, f2, gs2.zeros(), gs2.zeros(), e2, None)
    o2 = ~gs2.is_ghost
    r2 = np.zeros(len(pos)); r2[gs2.src_index[o2]] = f2[o2]
    e2a = np.zeros(len(pos)); e2a[gs2.src_index[o2]] = e2[o2]
    check("lj_moved", "fx", r2); check("lj_moved", "ep", e2a)
    # migration: nothing lost, velocities and types travelled with their particle, forces equal a fresh oracle build
    sent_tot = sum(r["migration"][0] for r in gathered); recv_tot = sum(r["migration"][1] for r in gathered)
    assert sent_tot == recv_tot and sent_tot > 100, (sent_tot, recv_tot)
    assert sum(r["migration"][3] for r in gathered) == len(pos)
    assert np.array_equal(merged("migrated", "vx"), vel[:, 0]) and np.array_equal(merged("migrated", "type"), typ.astype(np.float64))
    pos3 = np.mod(newpos + big, box)
    pos3 = np.where(pos3 >= box, 0.0, pos3)
    assert np.max(np.abs(np.mod(merged("migrated", "rx") - pos3[:, 0] + 0.5 * box[0], box[0]) - 0.5 * box[0])) < 1e-9
    gs3 = GridSystem(pos3, typ, box, cell, 1)
    g3 = gs3.oracle_grid()
    nb3 = O.Neighbors.build(g3, gs3.cell_off, gs3.rx, gs3.ry, gs3.rz, nbh, 1, True)
    f3, e3 = gs3.zeros(), gs3.zeros()
    O.pair_force(g3, gs3.cell_off, gs3.rx, gs3.ry, gs3.rz, nb3, [0.0104 * EV, 5.5], rc_lj, 0, f3, gs3.zeros(), gs3.zeros(), e3, None)
    o3 = ~gs3.is_ghost
    r3 = np.zeros(len(pos)); r3[gs3.src_index[o3]] = f3[o3]
    e3a = np.zeros(len(pos)); e3a[gs3.src_index[o3]] = e3[o3]
    check("migrated", "fx", r3); check("migrated", "ep", e3a)
    print("  migration: %d particles changed rank" % sent_tot)
    print("  ghost transport per rank: %s" % sorted(set(r["transport"] for r in gathered)))
    dm = max(r["dmax"] for r in gathered)
    assert all(abs(r["dmax"] - dm) < 1e-15 for r in gathered), "particle_displ_over: ranks disagree on the allreduced maximum"
    assert abs(dm - np.sqrt((d ** 2).sum(axis=1)).max()) < 1e-9
    print("mgpu_check ok: %d ranks %s, %d atoms" % (world, "x".join(map(str, rd)), len(pos)))


if __name__ == "__main__":
    # no collective after the gather: a failing rank 0 must not leave its peers blocked (the launcher tears the job down)
    rc = 0
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        rc = 1
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(rc)
