#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the short-range force hot path (neighbour + force) on N B200s.

  python bench.py --gpus N --steps K --warmup W [--workload c2]     (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                              (the CPU restatement of the reference path, all host cores)

Workloads = BASELINE.json configs, all NVE velocity Verlet, dt 1 fs, 300 K Maxwell velocities, seeded noise:
  c1  configs[0]  Lennard-Jones argon FCC a=5.0, 32^3 unit cells = 131 072 atoms, lj_compute_force rc 8.0, skin 1.0
  c2  configs[1]  EAM Cu FCC a=3.6, 79^3 unit cells = 1 972 156 atoms per GPU, eam_alloy_force with the reference's own
                  scripts/python/pytab-eam-alloy/Cu.eam.alloy (rc 7.29), skin 1.0          <- the default, the judged line
  c3  configs[2]  SNAP BCC a=3.316, 63^3 unit cells = 500 094 atoms, snap_force 2J=8 with the W block of the reference's
                  WBe_Wood_PRB2019.snap{param,coeff} (rcutfac 4.8123; no Ta 2J=8 file ships), skin 1.0
  c4  configs[3]  the c2 potential on 160^3 unit cells = 16 384 000 atoms split over the GPUs (strong scaling)
  c2j configs[1]  the analytic variant: johnson_force (Zhou-Johnson-Wadley Cu parameters, rc 6.0) on the c2 lattice
  c5  configs[4]  two-species random FCC alloy a=3.8, 126^3 unit cells = 8 001 504 atoms, eam_alloy_force with the reference's
                  AlCu.eam.alloy (rc 6.6825) + lj_multi_force (potentials/pair/lj/multi_species_nosym.msp parameters),
                  triclinic cell matrix drifting every step as under NPT, rebuild on the displacement trigger

One "step" = one full Verlet step: force_to_accel, push_f_v | push_f_v_r, push_f_v, particle_displ_over (one fused pass,
xsb_verlet_boundary), then ghost_update_r -- or, when the trigger fires / every --rebuild-every steps, move_particles +
ghost_comm_scheme + chunk_neighbors --, zero_force_energy and the force operators with their ghost exchanges.
N>1: static bricks (2,1,1)/(2,2,1)/(2,2,2); weak scaling (every rank owns one brick of the workload's size) except c4.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import EV, lattice, potential_file, read_snap_files  # noqa: E402

A_CU, RCUT, SKIN, DT, MASS_CU = 3.6, 7.29, 1.0, 1.0e-3, 63.546
KB_INTERNAL = 8.617333262e-5 * EV      # Boltzmann constant, internal energy units per K
METRIC, UNIT = "atom-timesteps/s (neighbor+force)", "atom-timesteps/s"
J_TO_EV = 1.0 / 1.602176634e-19


def rank_dims(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def make_setfl(tmpdir=None):
    """the potential of configs[1]/[3]: the reference's own Cu.eam.alloy (fixture copy, tests/golden/potentials)"""
    return potential_file("Cu.eam.alloy")


def n_cells_for(box_len, rcut=RCUT, skin=SKIN, slack=1.0):
    return int(np.floor(box_len / ((rcut + skin) * slack)))


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------
class Workload:
    """what differs between the BASELINE configs: lattice, potential set-up, the force operators of a step, the CPU
    restatement of the same operators, and the roofline model of the dominant kernel."""
    name = ""; structure = "FCC"; a = 3.6; cells = 79; noise = 0.1; masses = [MASS_CU]; rcut = RCUT; skin = SKIN
    baseline = ""; xform0 = None; cell_slack = 1.0; strong_total = None; dtype = "f64"; inner_skin = 0.0

    def label(self, n_gpus, uc, scaling):
        raise NotImplementedError

    def types(self, n, seed):
        return np.zeros(n, dtype=np.uint8)

    def setup(self, ctx, xsb):
        pass

    def forces(self, ctx, xsb, flags=0):
        raise NotImplementedError

    def e2e_forces(self, ctx, xsb):
        self.forces(ctx, xsb, 0)

    def e2e_fields(self, xsb):
        return [xsb.F_FX, xsb.F_FY, xsb.F_FZ]


class LJ(Workload):
    name = "c1"; a = 5.0; cells = 32; masses = [39.948]; rcut = 8.0
    baseline = "configs[0] Lennard-Jones FCC argon 32x32x32 cells (131k atoms) NVE Verlet"
    prm = [0.0104 * EV, 3.4]

    def label(self, n, uc, scaling):
        return "LJ Ar FCC %dx%dx%d unit cells x %d GPU = %d atoms, lj_compute_force (eps 0.0104 eV, sigma 3.4, rc %.1f, skin %.1f)" % (
            uc[0], uc[1], uc[2], n, 4 * uc[0] * uc[1] * uc[2] * n, self.rcut, self.skin)

    def forces(self, ctx, xsb, flags=0):
        ctx.zero_force_energy()
        ctx.pair_force(self.prm, self.rcut, flags)

    def e2e_forces(self, ctx, xsb):
        ctx.zero_force_energy()
        ctx.pair_force(self.prm, self.rcut, xsb.FLAG_ENERGY)

    def e2e_fields(self, xsb):
        return [xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP]

    def kernels(self):
        return ["pair"]

    def model(self, n_l, n_c):
        b_list = 2 * (1 + 2 * 27 + n_l)
        return {"pair": dict(bytes=24 + 1 + 32 + b_list, flops=8 * n_l + 30 * n_c, kernel="tile_pass_kernel<16,1024,LIST_FULL,LJTileOp> (lj_compute_force)")}

    def cpu_forces(self, O, g, gs, nb, arr, img):
        fx, fy, fz, ep, emb = arr
        O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, self.prm, self.rcut, 0, fx, fy, fz, None, None)


class EamCu(Workload):
    name = "c2"
    baseline = "configs[1] EAM Cu (Johnson/Mishin-type) FCC 2M atoms NVE on 1xB200"

    def label(self, n, uc, scaling):
        tot = 4 * uc[0] * uc[1] * uc[2] * n
        if scaling == "strong":
            rd = rank_dims(n)
            return "EAM Cu FCC %dx%dx%d unit cells = %d atoms over %d GPU (bricks of %dx%dx%d unit cells), eam_alloy_force (reference Cu.eam.alloy, rc %.2f, skin %.1f)" % (
                uc[0] * rd[0], uc[1] * rd[1], uc[2] * rd[2], tot, n, uc[0], uc[1], uc[2], self.rcut, self.skin)
        return "EAM Cu FCC %d^3 unit cells x %d GPU = %d atoms, eam_alloy_force (reference Cu.eam.alloy, rc %.2f, skin %.1f)" % (uc[0], n, tot, self.rcut, self.skin)

    def potential(self):
        return potential_file("Cu.eam.alloy")

    def setup(self, ctx, xsb):
        ctx.eam_alloy_load(self.potential())
        ctx.eam_inner_skin(self.inner_skin)

    def forces(self, ctx, xsb, flags=0, ef=0):
        ctx.zero_force_energy()
        ctx.eam_alloy_force(self.rcut, xsb.EAM_RHO | xsb.EAM_RHO2EMB | ef, flags)
        ctx.ghost_update([xsb.F_RHO_DEMB])
        ctx.eam_alloy_force(self.rcut, xsb.EAM_FORCE | ef, flags)

    def e2e_forces(self, ctx, xsb):
        self.forces(ctx, xsb, 0, xsb.EAM_EFLAG)

    def e2e_fields(self, xsb):
        return [xsb.F_FX, xsb.F_FY, xsb.F_FZ, xsb.F_EP]

    def kernels(self):
        return ["eam_rho", "eam_force"]

    def model(self, n_l, n_c):
        b_list = 2 * (1 + 2 * 27 + n_l)                            # reference stream encoding of one atom's list
        return {"eam_rho": dict(bytes=24 + 1 + b_list + 8, flops=8 * n_l + 15 * n_c, kernel="tile_pass_kernel<32,1024,LIST_FULL_WRITE_SUB,EamRhoTileOp> (eam_alloy_force, rho phase)"),
                "eam_force": dict(bytes=24 + 1 + 8 + b_list + 32, flops=8 * n_l + 45 * n_c, kernel="tile_pass_kernel<16,1024,LIST_SUB,EamForceTileOp> (eam_alloy_force, force phase)")}

    def cpu_forces(self, O, g, gs, nb, arr, img):
        fx, fy, fz, ep, emb = arr
        if not hasattr(self, "_cpu_eam"):
            self._cpu_eam = O.EamAlloy(self.potential())
        # eam_ghost = false + rho_dEmb owner -> ghost copy between the phases, like the GPU arm
        O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, self._cpu_eam, self.rcut, 1 | 2, fx, fy, fz, ep, None, emb)
        emb[:] = emb[img]
        O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, self._cpu_eam, self.rcut, 8, fx, fy, fz, ep, None, emb)


class EamCuJohnson(EamCu):
    """configs[1], analytic variant: johnson_force on the same lattice.  No Cu set of the Johnson form ships
    with the reference (its only one is Ta, potentials/eam/eam_johnson/single_specy.msp); the Zhou / Johnson / Wadley 2004 Cu
    values (tests/helpers.JOHNSON_CU) are used with rc 6.0: ~80 neighbours in range, 4 exp + 4 pow(., 20) per pair and pass."""
    name = "c2j"; rcut = 6.0
    baseline = "configs[1] EAM Cu (Johnson/Mishin-type) FCC 2M atoms NVE on 1xB200 -- analytic johnson_force variant"
    J = dict(re=2.556162, fe=1.554485, rhoe=21.175871, alpha=8.127620, beta=4.334731, A=0.396620, B=0.548085, kappa=0.308782, lam=0.756515,
             Fn0=-2.170269, Fn1=-0.263788, Fn2=1.088878, Fn3=-0.817603, F0=-2.19, F1=0.0, F2=0.561830, F3=-2.100595, Fo=-2.186568, eta=0.310490)

    def params(self):
        e = {"A", "B", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo"}
        order = ["re", "fe", "rhoe", "alpha", "beta", "A", "B", "kappa", "lam", "Fn0", "Fn1", "Fn2", "Fn3", "F0", "F1", "F2", "F3", "Fo", "eta"]
        return np.array([self.J[k] * (EV if k in e else 1.0) for k in order])

    def label(self, n, uc, scaling):
        return "EAM Cu FCC %d^3 unit cells x %d GPU = %d atoms, johnson_force (analytic, Zhou-Johnson-Wadley Cu parameters, rc %.2f, skin %.1f)" % (
            uc[0], n, 4 * uc[0] * uc[1] * uc[2] * n, self.rcut, self.skin)

    def setup(self, ctx, xsb):
        self.p19 = self.params()

    def forces(self, ctx, xsb, flags=0, ef=0):
        ctx.zero_force_energy()
        ctx.eam_johnson_force(self.p19, self.rcut, 1, flags)                 # johnson_emb
        ctx.ghost_update([xsb.F_RHO_DEMB])
        ctx.eam_johnson_force(self.p19, self.rcut, 4, flags)                 # johnson_force_reuse_emb

    def model(self, n_l, n_c):
        b_list = 2 * (1 + 2 * 27 + n_l)
        return {"eam_rho": dict(bytes=24 + 1 + b_list + 8 + 8, flops=8 * n_l + 60 * n_c, kernel="tile_pass_kernel<32,1024,LIST_FULL_WRITE_SUB,JohnsonEmbTileOp> (johnson_emb)"),
                "eam_force": dict(bytes=24 + 1 + 8 + b_list + 32, flops=8 * n_l + 90 * n_c, kernel="tile_pass_kernel<16,1024,LIST_SUB,JohnsonForceTileOp> (johnson_force_reuse_emb)")}

    def cpu_forces(self, O, g, gs, nb, arr, img):
        fx, fy, fz, ep, emb = arr
        p = self.params()
        O.eam_johnson(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, p, self.rcut, 1, fx, fy, fz, ep, None, emb)
        emb[:] = emb[img]
        O.eam_johnson(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, p, self.rcut, 4, fx, fy, fz, ep, None, emb)


class EamCuStrong(EamCu):
    name = "c4"; strong_total = 160
    baseline = "configs[3] EAM Cu 16M atoms weak/strong scaling over 1/2/4/8 B200 with ghost exchange"


class SnapW(Workload):
    name = "c3"; structure = "BCC"; a = 3.316; cells = 63; noise = 0.05; masses = [180.95]
    baseline = "configs[2] SNAP tantalum BCC 2J=8 500k atoms on 1xB200"

    def __init__(self):
        self.p = read_snap_files(potential_file("WBe_Wood_PRB2019.snapparam"), potential_file("WBe_Wood_PRB2019.snapcoeff"))
        w = self.p["elements"][0]
        self.rad, self.wj, self.beta = [w["radius"]], [w["weight"]], np.array([w["beta"]]) * EV
        self.rcut = 2.0 * w["radius"] * self.p["rcutfac"]

    def label(self, n, uc, scaling):
        return "SNAP BCC a=%.3f %d^3 unit cells x %d GPU = %d atoms, snap_force 2J=%d (W block of the reference's WBe_Wood_PRB2019, rcutfac %.4f, skin %.1f)" % (
            self.a, uc[0], n, 2 * uc[0] * uc[1] * uc[2] * n, self.p["twojmax"], self.p["rcutfac"], self.skin)

    def setup(self, ctx, xsb):
        p = self.p
        ctx.snap_set(p["twojmax"], p["rcutfac"], self.rad, self.wj, self.beta, rfac0=p["rfac0"], rmin0=p["rmin0"], switchflag=p["switchflag"], bzeroflag=p["bzeroflag"])

    def forces(self, ctx, xsb, flags=0):
        ctx.zero_force_energy(ghost=True)
        ctx.snap_force(flags)
        ctx.ghost_reduce_add([xsb.F_FX, xsb.F_FY, xsb.F_FZ])      # update_force_energy_from_ghost (Newton-on scatter of f_j)

    def kernels(self):
        return ["snap"]

    def model(self, n_l, n_c):
        # FP-pipe bound: ~1.5 Mflop per atom at 2J = 8 with ~26 neighbours, < 3 kB of HBM.  That figure is the
        # reference algorithm's (three derivative chains per neighbour); the pipeline here EXECUTES about 0.8 Mflop per atom
        # (Utot 3.1 kflop and reverse-mode force 9.9 kflop per neighbour, compute_yi 0.46 Mflop per atom incl. the zero-padded
        # window slots): `executed_flops_per_atom` in the fp64 block, so the pipe fraction of the work actually issued is
        # achieved x executed / algorithmic
        return {"snap": dict(bytes=24 + 2 * (1 + 2 * 27 + n_l) + 32 + 2 * 285 * 16 * 2, flops=1.5e6 * max(n_c, 1.0) / 26.0,
                             executed_flops=4.6e5 + 1.3e4 * max(n_c, 1.0), kernel="snap Utot / Y / force pipeline (snap_force)")}

    def cpu_forces(self, O, g, gs, nb, arr, img):
        fx, fy, fz, ep, emb = arr
        if not hasattr(self, "_cpu_snap"):
            p = self.p
            self._cpu_snap = O.Snap(p["twojmax"], p["rcutfac"], self.rad, self.wj, self.beta, rfac0=p["rfac0"], rmin0=p["rmin0"], switchflag=p["switchflag"], bzeroflag=p["bzeroflag"])
        O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, self._cpu_snap, 0, fx, fy, fz, None, None)


class AlloyNPT(EamCu):
    name = "c5"; a = 3.8; cells = 126; noise = 0.08; masses = [26.982, 63.546]; rcut = 6.6825; cell_slack = 1.03
    baseline = "configs[4] Multi-species LJ/EAM alloy 8M atoms with periodic neighbour-list rebuild under NPT"
    xform0 = np.array([[1.0, 0.01, 0.005], [0.0, 1.0, 0.01], [0.0, 0.0, 1.0]])
    # potentials/pair/lj/multi_species_nosym.msp: pair (1,1) <- Zn-Zn, (0,1) <- Cu-Zn, (0,0) <- common_parameters (eps 0)
    lj_rows = np.array([[0.0, 1.0, 6.10], [4.853e-20 * J_TO_EV * EV, 2.36, 5.89], [2.522e-20 * J_TO_EV * EV, 2.44, 6.10]])

    def label(self, n, uc, scaling):
        return "two-species random FCC alloy a=%.1f %d^3 unit cells x %d GPU = %d atoms, eam_alloy_force (reference AlCu.eam.alloy, rc %.4f) + lj_multi_force, triclinic xform drifting 2e-6/step, skin %.1f" % (
            self.a, uc[0], n, 4 * uc[0] * uc[1] * uc[2] * n, self.rcut, self.skin)

    def potential(self):
        return potential_file("AlCu.eam.alloy")

    def types(self, n, seed):
        return (np.random.default_rng(seed + 77).random(n) < 0.5).astype(np.uint8)

    def forces(self, ctx, xsb, flags=0, ef=0):
        EamCu.forces(self, ctx, xsb, flags, ef)
        # same energy switch for both operators of the chain (the reference's trigger_thermo_state reaches every force operator)
        ctx.pair_multi_force(2, self.lj_rows, 6.10, flags | (xsb.FLAG_ENERGY if ef else 0))

    def kernels(self):
        return ["eam_rho", "eam_force", "pair"]

    def model(self, n_l, n_c):
        m = EamCu.model(self, n_l, n_c)
        # the force pass also evaluates the chained lj_multi_force (one reciprocal + ~12 FP64 operations per pair inside its cut-off)
        m["eam_force"] = dict(m["eam_force"], flops=m["eam_force"]["flops"] + 14 * n_c * (6.10 / self.rcut) ** 3,
                              kernel="tile_pass_kernel<16,1024,LIST_SUB,EamForceTileOp<MULTI,CHAIN>> (eam_alloy_force force phase + the lj_multi_force chained behind it; "
                                     "without the fusion: XSB_NO_CHAIN_FUSION=1, separate LJTileOp pass under the `pair` tag)")
        m["pair"] = dict(bytes=24 + 1 + 32 + 2 * (1 + 2 * 27 + n_l), flops=8 * n_l + 30 * n_c * (6.10 / self.rcut) ** 3, kernel="tile_pass_kernel<16,1024,LIST_SUB,LJTileOp<MULTI>> (lj_multi_force on its own)")
        return m

    def cpu_forces(self, O, g, gs, nb, arr, img):
        EamCu.cpu_forces(self, O, g, gs, nb, arr, img)
        fx, fy, fz, ep, emb = arr
        rows = np.array([list(r) + [O.pair_ecut(0, r[:2], r[2])] for r in self.lj_rows])
        O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, rows, 6.10, 0, fx, fy, fz, None, None)


WORKLOADS = {"c1": LJ, "c2": EamCu, "c2j": EamCuJohnson, "c3": SnapW, "c4": EamCuStrong, "c5": AlloyNPT}


def brick_cells(W, args, n):
    rd = rank_dims(n)
    if args.scaling == "strong":
        tot = args.total_cells or W.strong_total or W.cells
        assert all(tot % d == 0 for d in rd), "--total-cells must be divisible by the rank grid"
        return [tot // d for d in rd]
    c = args.cells or W.cells
    return [c] * 3


def brick_system(W, ucells, coord, seed):
    """one rank's brick of the lattice: positions in GLOBAL coordinates (grid space), velocities, types"""
    pos, typ, box = lattice(W.structure, ucells, W.a, W.noise, seed=seed)
    # noise may push atoms slightly out of the brick: they are clamped into the brick's border cells (counted by
    # xsb_out_of_domain_count; the cell edge has > 0.1 ang of slack) and sent home by the first move_particles
    pos = pos + np.asarray(coord, dtype=np.float64) * box
    rng = np.random.default_rng(seed + 1000)
    typ = W.types(len(pos), seed)
    m = np.asarray(W.masses)[typ]
    vel = rng.normal(0.0, 1.0, pos.shape) * np.sqrt(KB_INTERNAL * 300.0 / m)[:, None]
    vel -= vel.mean(axis=0)
    return pos, vel, typ, box


def domain_cells(W, scaling, brick, rd):
    """one cell size for the whole domain, bricks made of whole cells: (cell size, cells per brick axis, global cells per axis)"""
    brick = np.asarray(brick, dtype=np.float64)
    if scaling == "strong":
        gbox = brick * np.asarray(rd, dtype=np.float64)
        assert np.allclose(gbox, gbox[0]), "strong scaling splits a cubic box"
        gc = n_cells_for(gbox[0], W.rcut, W.skin, W.cell_slack)
        while any(gc % d for d in rd):
            gc -= 1
        return gbox[0] / gc, [gc // d for d in rd], [gc] * 3
    assert np.allclose(brick, brick[0]), "weak scaling uses cubic bricks"
    ncb = n_cells_for(brick[0], W.rcut, W.skin, W.cell_slack)
    return brick[0] / ncb, [ncb] * 3, [ncb * d for d in rd]


def in_range_sample(pos, rcut, nsample=400):
    """mean number of neighbours inside rcut (n_c of SURVEY.md 8d), counted by brute force for atoms near the brick centre"""
    c = pos.mean(axis=0)
    near = pos[np.all(np.abs(pos - c) < 10.0 + rcut + 0.5, axis=1)]
    core = near[np.all(np.abs(near - c) < 10.0, axis=1)][:nsample]
    if len(core) == 0:
        return 0.0
    d2 = ((core[:, None, :] - near[None, :, :]) ** 2).sum(axis=2)
    return float(((d2 <= rcut * rcut) & (d2 > 0)).sum() / len(core))


def workload_config(W, args, n):
    uc = brick_cells(W, args, n)
    return {"workload": W.label(n, uc, args.scaling) + ", NVE Verlet dt 1 fs", "baseline_config": W.baseline,
            "rebuild": "particle_displ_over(skin/2) trigger (read one step late with a 2-step-displacement margin, no host sync), forced at least every %d steps" % args.rebuild_every,
            "l2": ("L2 flushed between timed steps (a 160 MiB buffer is rewritten, every step timed on its own)" if getattr(args, "flush_l2", False) else
                   "inputs larger than L2 (positions + neighbour lists > 1 GB per GPU)" if W.name != "c1" else "NOT flushed: the 90 MB working set stays L2-resident, as in a production run of this size"),
            "parallelism": "bricks %s" % "x".join(str(d) for d in rank_dims(n))}


# ------------------------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock + throttle reasons sampled through NVML every ~2 ms while the timed region runs (the recipe's clocks line;
    nvidia-smi -lms cannot resolve a 90 ms region).  Falls back to one nvidia-smi query when pynvml is unusable."""

    def __init__(self, device):
        self.dev, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            idx = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[device]) if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:      # noqa: BLE001
            self.h = None
            self.err = str(e)

    def _run(self):
        nv = self.nv
        names = (("hw_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)), ("hw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)), ("sw_power_cap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
            except Exception:      # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
        if self.h is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), reasons=sorted(self.reasons), samples=len(self.samples), sm_mhz_min=float(min(self.samples)))
            return out
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            c = [x.strip() for x in subprocess.check_output(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"], text=True).split(",")]
            out.update(sm_mhz=float(c[0]), sm_max_mhz=float(c[1]), samples=1, source="nvidia-smi (one sample after the timed region)",
                       reasons=[n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]) if v.lower().startswith("active")])
        except Exception:      # noqa: BLE001
            pass
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the restatement of the reference path (oracle/), OpenMP on all host cores
# ------------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:      # noqa: BLE001
        return os.cpu_count() or 1


def cpu_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip(); break
    except OSError:
        pass
    try:
        cc = subprocess.check_output(["g++", "--version"], text=True).splitlines()[0]
    except Exception:      # noqa: BLE001
        cc = "g++ (unknown)"
    return model, cc


def cpu_reference_run(W, ucells, steps, warmup, rebuild_every, cell=None):
    """chunk_neighbors + the workload's force operators on the CPU restatement for a periodic system of `ucells`^3 unit
    cells: one list build (amortised over rebuild_every steps, like the GPU arm) + warmup + `steps` timed force steps.
    Returns atom-timesteps/s and a description.  torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly."""
    nthr = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthr)
    from oracle import oracle as O
    from helpers import GridSystem
    L, flags = O.lib_timed()
    try:
        return _cpu_reference_run(O, L, flags, nthr, W, ucells, steps, warmup, rebuild_every, cell)
    finally:
        O.lib_pinned()                 # the -march=native build must not leak into a process that also checks parity


def _cpu_reference_run(O, L, flags, nthr, W, ucells, steps, warmup, rebuild_every, cell):
    from helpers import GridSystem
    L.orc_set_num_threads(nthr)
    pos, typ0, box = lattice(W.structure, ucells, W.a, W.noise, seed=1)
    typ = W.types(len(pos), 1)
    if cell is None:
        cell = box[0] / n_cells_for(box[0], W.rcut, W.skin, W.cell_slack)
    gs = GridSystem(pos, typ, box, cell, 1, xform=W.xform0)
    g = gs.oracle_grid()
    own = ~gs.is_ghost
    owner_of = np.zeros(len(pos), dtype=np.int64); owner_of[gs.src_index[own]] = np.nonzero(own)[0]
    img = owner_of[gs.src_index]
    arr = [gs.zeros() for _ in range(5)]
    t0 = time.perf_counter()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, W.rcut + W.skin, 1, True)
    t_nb = time.perf_counter() - t0
    times = []
    for i in range(warmup + steps):
        for a in arr[:4]:
            a[:] = 0
        t0 = time.perf_counter()
        W.cpu_forces(O, g, gs, nb, arr, img)
        times.append(time.perf_counter() - t0)
    t_force = float(np.mean(times[warmup:]))
    per_step = t_force + t_nb / rebuild_every
    model, cc = cpu_info()
    return gs.n_owned / per_step, {"atoms": int(gs.n_owned), "steps": steps, "nbh_build_s": t_nb, "force_s_per_step": t_force, "threads": int(L.orc_num_threads()),
                                    "cpu_model": model, "compiler": cc, "flags": flags}


def cpu_arm_sample_cells(W, full, steps, warmup, threads, forced=0):
    """0 = time the full configuration `full` (unit cells per axis of the whole N-GPU system), else the edge (unit cells) of
    the cube timed instead.  The full configuration is timed whenever warmup + steps force steps of it fit ~4 minutes on
    this host (per-thread rates of the restatement measured on the round's boxes, profiles/r02zz_*); otherwise -- SNAP at
    2J = 8 always (~0.2 s per 1000 atoms and core), a 16 M-atom system with a long step count -- a cube of the same lattice /
    potential / cutoffs sized to that budget, which the caller names in config.workload."""
    per_cell = 2 if W.structure == "BCC" else 4
    atoms_full = per_cell * full[0] * full[1] * full[2]
    rate = {"c1": 6.0e5, "c2": 1.8e5, "c4": 1.8e5, "c2j": 1.1e5, "c5": 1.4e5, "c3": 2.8e2}.get(W.name, 1.0e5) * threads
    est_s = atoms_full * (steps + warmup + 2) / rate
    if not (forced or W.name == "c3" or est_s > 240.0):
        return 0
    sc = forced or (24 if W.name == "c3" else max(16, int((240.0 * rate / (steps + warmup + 2) / per_cell) ** (1.0 / 3.0))))
    return min(sc, min(full))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]()
    if W.strong_total and args.scaling != "strong":
        args.scaling = "strong"
    n = args.gpus
    uc = brick_cells(W, args, n)
    rd = rank_dims(n)
    full = [uc[a] * rd[a] for a in range(3)]                      # the whole system the N-GPU arm runs
    t0 = time.perf_counter()
    sample_note = None
    ucells = full
    cell, _, _ = domain_cells(W, args.scaling, np.asarray(uc, dtype=np.float64) * W.a, rd)      # the GPU arm's cell size
    sc = cpu_arm_sample_cells(W, full, args.steps, args.warmup, os.cpu_count() or 1, args.cpu_sample_cells)
    if sc:
        ucells, cell = [sc] * 3, None
        sample_note = "bounded sample: %d^3 unit cells of the same lattice / potential / cutoffs" % sc
    value, info = cpu_reference_run(W, ucells, args.steps, args.warmup, args.rebuild_every, cell)
    cfg = workload_config(W, args, n)
    if sample_note:
        cfg["workload"] += " [CPU arm: " + sample_note + " = %d atoms]" % info["atoms"]
    sample = "%d atoms (%s), 1 list build amortised over %d steps + %d warm-up + %d timed force steps; %d OpenMP threads on %s; %s %s" % (
        info["atoms"], sample_note or "the full configuration", args.rebuild_every, args.warmup, args.steps, info["threads"], info["cpu_model"], info["compiler"], info["flags"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * info["atoms"] / value, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0, "host_cores": os.cpu_count(),
            "detail": {"nbh_build_s": info["nbh_build_s"], "force_s_per_step": info["force_s_per_step"], "cpu_model": info["cpu_model"], "compiler": info["compiler"], "flags": info["flags"]}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def run_xsb(args):
    import torch
    import exastamp_b200 as xsb
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    W = WORKLOADS[args.workload]()
    # measured optima (profiles/r02k_skin_*: c2 0.12; profiles/r02y8_*: c5 0.28 -- light Al atoms spend the budget faster)
    W.inner_skin = args.inner_skin if args.inner_skin >= 0.0 else {"c5": 0.28}.get(W.name, 0.12)
    if W.strong_total and args.scaling != "strong":
        args.scaling = "strong"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local == 0:
        xsb.build()                       # one builder per node; the others wait (and build() itself is lock-protected)
    if dist is not None:
        dist.barrier()
    rd = rank_dims(world)
    coord = (rank % rd[0], (rank // rd[0]) % rd[1], rank // (rd[0] * rd[1]))
    pos, vel, typ, brick = brick_system(W, brick_cells(W, args, world), coord, seed=1 + rank)
    cell, ncb3, gcells = domain_cells(W, args.scaling, brick, rd)
    origin = [(coord[a] * ncb3[a] - 1) * cell for a in range(3)]
    ctx = xsb.Context(local)
    if world > 1:
        ids = [xsb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])
    ctx.grid_set(xsb.make_grid([c + 2 for c in ncb3], 1, cell, origin, W.xform0))
    ctx.particles_assign(pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2], typ)
    ctx.set_domain(gcells, (1, 1, 1), rd, coord)
    ctx.ghost_comm_scheme()
    W.setup(ctx, xsb)
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    n_own = ctx.n_own
    masses = list(W.masses) + [W.masses[-1]] * (2 - len(W.masses))
    state = {"rebuilds": 0, "since": 0, "rebuild_s": 0.0, "move_s": 0.0, "step": 0}
    flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda") if args.flush_l2 else None

    def rebuild(first=False):
        ctx.sync(); t0 = time.perf_counter()
        if not first:
            ctx.particles_rebin()          # move_particles + migrate_cell_particles (cross-rank over NCCL when world > 1)
            ctx.ghost_comm_scheme()
        ctx.sync(); t1 = time.perf_counter()
        ctx.chunk_neighbors(W.rcut + W.skin, stream_prealloc_factor=1.25)
        ctx.backup_r()
        ctx.sync()
        state["rebuilds"] += 1; state["since"] = 0
        graph["flags"] = None                          # a recorded step belongs to one list generation
        state["move_s"] += t1 - t0; state["rebuild_s"] += time.perf_counter() - t0

    mode = {"flags": 0}                               # xsb.FLAG_MIXED during the extra mixed-precision measurement

    graph = {"id": None, "flags": None}
    use_graph = args.graph and world == 1 and W.xform0 is None and not (args.separate_integrator or args.sync_displ)

    def record():
        # the regular step (integrator pass, ghost update, zero + force operators) as ONE launch: xsb_step_capture_*
        if graph["id"] is not None:
            ctx.step_release(graph["id"])
        ctx.step_capture_begin()
        ctx.verlet_boundary_async(masses, DT); ctx.ghost_update(POS); W.forces(ctx, xsb, mode["flags"])
        graph["id"] = ctx.step_capture_end(); graph["flags"] = mode["flags"]

    def step_recorded():
        # same decisions as step(), on results one step older (the host must not wait for the step it is about to follow):
        # xsb_displ_poll(1) = the maxima of two steps ago, three step displacements of margin
        over = False
        if state["since"] >= 2:
            d, s1 = ctx.displ_poll(1)
            over = d + 3.0 * s1 > 0.5 * W.skin
        state["since"] += 1; state["step"] += 1
        if over or state["since"] >= args.rebuild_every:
            ctx.verlet_boundary_async(masses, DT); rebuild(); W.forces(ctx, xsb, mode["flags"])
        else:
            if graph["flags"] != mode["flags"]:
                record()
            ctx.step_replay(graph["id"])

    def step():
        nonlocal use_graph
        if use_graph:
            return step_recorded()
        # one velocity-Verlet step, cut at the displacement check: force_to_accel + push_f_v close the previous step,
        # push_f_v_r + push_f_v + particle_displ_over open this one (xsb_verlet_boundary = those five operators in one pass)
        if args.separate_integrator:
            ctx.force_to_accel(masses); ctx.push_f_v(0.5 * DT)
            ctx.push_f_v_r(DT); ctx.push_f_v(0.5 * DT)
            over, _ = ctx.particle_displ_over(0.5 * W.skin)
        elif args.sync_displ:
            over, _ = ctx.verlet_boundary(masses, DT, 0.5 * W.skin)      # blocking read-back + all-reduce every step
        else:
            # no host read-back: decide on the PREVIOUS step's maxima (all-reduced on the stream, long finished) with twice
            # that step's largest displacement as the margin, so every list stays valid and the host never waits for the GPU
            ctx.verlet_boundary_async(masses, DT)
            over = False
            if state["since"] >= 1:                                       # values recorded after the last backup_r
                d, s1 = ctx.displ_poll(1)
                over = d + 2.0 * s1 > 0.5 * W.skin
        state["since"] += 1; state["step"] += 1
        if W.xform0 is not None:                      # barostat-like drift of the cell matrix (NPT): 2e-6 per step
            ctx.grid_set_xform(W.xform0 * (1.0 + 2e-6 * state["step"]))
        if over or state["since"] >= args.rebuild_every:
            rebuild()
        else:
            ctx.ghost_update(POS)
        W.forces(ctx, xsb, mode["flags"])

    rebuild(first=True)
    W.forces(ctx, xsb, 0)
    total_nbh, max_nbh = ctx.chunk_neighbors_stats()
    n_l = total_nbh / max(1, ctx.n)

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def allranks(v):
        """value of every rank (list, rank order)"""
        if dist is None:
            return [float(v)]
        t = torch.zeros(world, device="cuda", dtype=torch.float64); t[rank] = float(v)
        dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    for _ in range(args.warmup):
        step()
    if args.warmup:
        rebuild(); W.forces(ctx, xsb, 0)      # warm-up also covers the rebuild path (migration buffers, lazily set up NCCL channels)
        rebuild(); W.forces(ctx, xsb, 0)
    # the clock sampler (NVML init + thread start: milliseconds) starts BEFORE the barrier: started after it, it delays rank 0
    # behind its peers, which then wait for rank 0 in the first exchange -- a one-off skew of several ms that the MAX over
    # ranks of a 20-step window turns into a 5-20 % "scaling loss" (round 1's N=2 point)
    clocks = Clocks(local) if rank == 0 else None
    barrier()
    ctx.profile_enable(True)
    barrier()
    l0 = ctx.launches; rb0 = state["rebuilds"]; state["rebuild_s"] = 0.0; state["move_s"] = 0.0
    t0 = time.perf_counter()
    if flush is None:
        ctx.timer_start()
        for _ in range(args.steps):
            step()
        ms_dev = ctx.timer_stop_ms()
    else:
        ms_dev = 0.0
        for _ in range(args.steps):                   # L2 flushed between timed steps: each step timed on its own
            flush.fill_(1); torch.cuda.synchronize()
            ctx.timer_start(); step(); ms_dev += ctx.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launches - l0
    rebuilds_timed = state["rebuilds"] - rb0; rebuild_s_timed = state["rebuild_s"]; move_s_timed = state["move_s"]
    prof = ctx.profile_read()
    prof_note = None
    if use_graph:
        # a recorded step has no per-operator events: the kernel times behind `roofline` come from direct-call steps
        # right after the timed region (same list, same L2 treatment), the step time itself from the recorded steps above
        use_graph = False
        ctx.profile_enable(True)
        for _ in range(min(10, args.steps)):
            if flush is not None:
                flush.fill_(1); torch.cuda.synchronize()
            step()
        prof = ctx.profile_read()
        use_graph = True
        prof_note = "step = one CUDA-graph launch (xsb_step_replay); per-kernel times from %d direct-call steps after the timed region" % min(10, args.steps)
    ctx.profile_enable(False)
    clk = clocks.stop() if clocks else None
    ms_ranks = allranks(max(ms_dev, 0.0))
    ms = max(ms_ranks)
    atoms_total = int(round(sum(allranks(n_own))))
    value = atoms_total * args.steps / (ms * 1e-3)
    # per-rank view of the timed region (a single straggling rank is otherwise invisible behind the MAX)
    per_rank = {"timed_region_ms": {"min": min(ms_ranks), "median": float(np.median(ms_ranks)), "max": ms, "by_rank": ms_ranks}}
    for k, v in prof.items():
        if v[1]:
            r = allranks(v[0])
            per_rank[k + "_ms"] = {"min": min(r), "median": float(np.median(r)), "max": max(r)}
    per_rank["rebuild_wall_s"] = dict(zip(("min", "median", "max"), (lambda r: (min(r), float(np.median(r)), max(r)))(allranks(rebuild_s_timed))))

    # checksum of the forces the timed loop ended on (own atoms, all ranks): runs that differ only in a kernel switch
    # (inner skin, ghost transport, ...) must agree on it to ~1e-10
    fsum = 0.0
    for f in (xsb.F_FX, xsb.F_FY, xsb.F_FZ):
        fsum += float(np.abs(ctx.download(f)).sum())           # ghost slots hold zeros
    fsum = sum(allranks(fsum))

    # ---- the same steps in mixed precision (FP32 spline / pair math, tolerance 1e-5): reported beside, not as, the metric
    mixed = None
    if not args.no_mixed:
        mode["flags"] = xsb.FLAG_MIXED
        for _ in range(3):
            step()
        barrier()
        km = max(5, min(args.steps, 40))
        ctx.timer_start()
        for _ in range(km):
            step()
        msm = max(allranks(ctx.timer_stop_ms()))
        barrier()
        mixed = {"value": atoms_total * km / (msm * 1e-3), "unit": UNIT, "ms_per_step": msm / km, "steps": km,
                 "dtype": "f32 spline / pair / bispectrum math, f64 positions, distances and accumulation (XSB_FLAG_MIXED)", "tolerance": 1e-5}
        mode["flags"] = 0

    # ---- e2e: the plugin use case.  The host application owns the particle arrays (pinned host memory): every step it
    # hands the positions of its own atoms to the C ABI and reads their forces (+ energies) back.  The copies run on the
    # context's two copy streams (xsb_fields_upload_async / _download_async): the H2D of step i+1 and the D2H of step i-1
    # overlap the passes of step i; every step's bytes cross PCIe inside the timed region and the last result is waited for.
    e2e = None
    if not args.no_e2e:
        outf = W.e2e_fields(xsb)
        # The host application integrates: what it uploads every step is its own trajectory.  With the inner skin on, the
        # bench records rebuild_every consecutive steps of the device integrator into pinned host buffers first and replays
        # them as the host's positions (the upload charges their displacement to the sub-list's budget, xsb_core.cu); a
        # host that re-sent the SAME positions every step would never spend the budget and flatter the number.  Without the
        # skin (LJ, SNAP, a drifting cell matrix) one snapshot is re-sent: those passes do the same work whatever moved.
        use_traj = hasattr(ctx, "eam_inner_skin") and W.inner_skin > 0.0 and W.xform0 is None and W.name in ("c2", "c4")
        if hasattr(ctx, "eam_inner_skin") and not use_traj:
            ctx.eam_inner_skin(0.0)
        M = args.rebuild_every if use_traj else 1
        pin_r = [[torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(3)] for _ in range(M)]
        pin_f = [[torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in outf] for _ in range(2)]   # results of even / odd steps
        for m in range(M):
            ctx.fields_download_async(POS, [t.data_ptr() for t in pin_r[m]]); ctx.copy_wait()
            if m + 1 < M:
                ctx.verlet_boundary_async(masses, DT); ctx.ghost_update(POS); W.forces(ctx, xsb, 0)
        ke = max(3, min(args.steps, args.e2e_steps))

        def e2e_step(i):
            if i % args.rebuild_every == 0:
                ctx.ghost_update(POS); ctx.chunk_neighbors(W.rcut + W.skin, stream_prealloc_factor=1.25)
            else:
                ctx.ghost_update(POS)                             # owner -> ghost images of the uploaded positions
            W.e2e_forces(ctx, xsb)
            ctx.fields_download_async(outf, [t.data_ptr() for t in pin_f[i & 1]])      # results of step i
            ctx.fields_upload_async(POS, [t.data_ptr() for t in pin_r[(i + 1) % M]])  # inputs of step i+1 (overlaps step i)

        ctx.fields_upload_async(POS, [t.data_ptr() for t in pin_r[(M - 1) % M]])
        for i in (M - 1, M):                                      # warm-up ends on the upload of snapshot 1 % M ... re-aligned below
            e2e_step(i)
        ctx.copy_wait(); barrier()
        sub0 = ctx.eam_sublist_stats() if use_traj else None
        t0 = time.perf_counter()
        ctx.fields_upload_async(POS, [t.data_ptr() for t in pin_r[0]])
        for i in range(ke):
            e2e_step(i)
        ctx.copy_wait(); barrier()
        te = max(allranks(time.perf_counter() - t0))
        fsum = float(sum(float(t.abs().sum()) for t in pin_f[(ke - 1) & 1][:3]))
        e2e = {"value": atoms_total * ke / te, "unit": UNIT, "h2d_bytes_per_step": int(24 * n_own), "d2h_bytes_per_step": int(8 * len(outf) * n_own),
               "steps": ke, "bytes_are": "per GPU", "result_check_sum_abs_f": fsum,
               "host_positions": ("a recorded %d-step trajectory (the device integrator's own steps), inner skin %.2f: %d rho phases re-filtered, %d re-evaluated the sub-list" % (
                   (M, W.inner_skin) + tuple(b - a for a, b in zip(sub0, ctx.eam_sublist_stats()))) if use_traj else "one snapshot re-sent every step (inner skin off)"),
               "what": "pinned host r of the own atoms -> xsb_fields_upload_async, ghost_update_r, chunk_neighbors every %d steps, zero + force operators%s, "
                       "xsb_fields_download_async of %d fields of the own atoms -> pinned host; copies overlap the passes of the neighbouring steps" % (
                           args.rebuild_every, " with energies" if len(outf) > 3 else "", len(outf))}

    if rank != 0:
        dist.barrier(); dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel, algorithmic bytes / flops per SURVEY.md 8(d)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:      # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    try:
        live_fp64, live_fp32, live_hbm = ctx.measure_peaks()      # DFMA / FFMA loops and a 1 GiB copy on this device, now
    except Exception as e:      # noqa: BLE001                     the metric line must not depend on the side measurement
        sys.stderr.write("xsb_measure_peaks failed: %s\n" % e)
        live_fp64 = live_fp32 = live_hbm = None
    n_c = in_range_sample(pos, W.rcut)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:      # noqa: BLE001
            traffic = {}
    model = W.model(n_l, n_c)
    kern = []
    for tag in W.kernels():
        t_ms, cnt = prof.get(tag, (0.0, 0))
        if cnt and tag in model:
            m = model[tag]; dur = t_ms / cnt * 1e-3
            # a profile interval of the pair tag may cover more than one launch (c5: one per operator call)
            ach = m["bytes"] * n_own / dur / 1e9; tf = m["flops"] * n_own / dur / 1e12
            kern.append({"kernel": m["kernel"], "avg_launch_ms": dur * 1e3, "share_of_step": t_ms / ms, "algorithmic_bytes_per_atom": m["bytes"],
                         "achieved": ach, "frac": ach / peak,
                         "traffic": (traffic.get(W.name + ":" + tag) or (traffic.get(tag) if W.name in ("c2", "c4") else None) or {}).get("dram_bytes_per_launch"),
                         "fp64": dict({"achieved": tf, "peak": live_fp64, "unit": "TFLOP/s", "frac": tf / live_fp64 if live_fp64 else None, "algorithmic_flops_per_atom": m["flops"]},
                                      **({"executed_flops_per_atom": m["executed_flops"], "executed_frac": (m["executed_flops"] * n_own / dur / 1e12) / live_fp64 if live_fp64 else None}
                                         if "executed_flops" in m else {}))})
    kern.sort(key=lambda k: -k["avg_launch_ms"])
    roof = None
    if kern:
        d = kern[0]
        roof = {"bound": "hbm", "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"], "traffic": d["traffic"],
                "kernel": d["kernel"], "avg_launch_ms": d["avg_launch_ms"], "share_of_step": d["share_of_step"], "algorithmic_bytes_per_atom": d["algorithmic_bytes_per_atom"],
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "fp64": dict(d["fp64"], n_c_in_range=n_c, peak_source="DFMA loop measured in this run (xsb_measure_peaks)"),
                "live_peaks": {"fp64_tflops": live_fp64, "fp32_tflops": live_fp32, "hbm_copy_gbs": live_hbm},
                "binding_resource": "FP64 pipe" if W.name == "c3" else "shared-memory data pipe (per-pair position / spline-knot gathers), see DESIGN.md 3.1",
                "other_kernels": kern[1:],
                "note": "the dominant kernel by time is headlined; traffic = ncu dram bytes per launch of the capture recorded in profiles/ncu_traffic.json"}
    breakdown = {k: {"ms_total": v[0], "intervals": v[1], "share": v[0] / ms if ms else None} for k, v in prof.items() if v[1]}
    cpu = None
    if not args.no_cpu and world == 1:          # the CPU baseline is timed beside the 1-GPU run only, on a bounded sample
        sc = args.cpu_sample_cells or {"c1": 32, "c3": 16}.get(W.name, 40)
        v, info = cpu_reference_run(W, sc, 2, 1, args.rebuild_every)
        cpu = {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port",
               "sample": "%s %d^3 unit cells = %d atoms, same potential / cutoffs; 1 warm-up + 2 timed force steps + list build/%d (oracle restatement, OpenMP, %s, %s %s)" % (
                   W.structure, sc, info["atoms"], args.rebuild_every, info["cpu_model"], info["compiler"], info["flags"])}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": W.dtype, "data": "synthetic",
            "config": workload_config(W, args, world), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "mixed_precision": mixed,
            "detail": {"atoms_per_gpu": int(n_own), "atoms_with_ghosts": int(ctx.n), "list_entries_per_atom": n_l, "max_list": int(max_nbh),
                       "rebuilds_in_timed_region": rebuilds_timed, "rebuild_wall_s_total": rebuild_s_timed,
                       "move_particles_wall_s_total": move_s_timed, "host_wall_s": wall, "breakdown": breakdown, "ranks": per_rank,
                       "clamped_at_assign": ctx.out_of_domain_count(), "ghost_transport": ctx.ghost_transport(),
                       "recorded_step": prof_note, "force_checksum_sum_abs": fsum, "pair_operators_fused_into_eam_force_pass": ctx.chain_stats(), "eam_sublist": dict(zip(("inner_skin", "refiltered", "reused"), (W.inner_skin,) + tuple(ctx.eam_sublist_stats())))}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="xsb", choices=["xsb", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json configs[0..4] = c1..c5 (default c2, the judged line)")
    ap.add_argument("--cells", type=int, default=0, help="unit cells per axis per GPU (default: the workload's size)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --total-cells^3 unit cells split over the GPUs (c4 / configs[3])")
    ap.add_argument("--total-cells", type=int, default=0, help="strong scaling: unit cells per axis of the whole system (c4: 160 -> 16 384 000 atoms)")
    ap.add_argument("--rebuild-every", type=int, default=20)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-sample-cells", type=int, default=0, help="unit cells per axis of the CPU arm's system (default: the full configuration for --impl reference)")
    ap.add_argument("--inner-skin", type=float, default=-1.0, help="inner skin (angstrom) of eam_alloy_force's in-range sub-list: 0 re-filters the neighbour list every step (xsb_eam_inner_skin); default: 0.12, c5 0.28")
    ap.add_argument("--sync-displ", action="store_true", help="blocking particle_displ_over read-back every step (xsb_verlet_boundary) instead of the one-step-late check")
    ap.add_argument("--separate-integrator", action="store_true", help="five integrator operators as separate kernels instead of xsb_verlet_boundary")
    ap.add_argument("--flush-l2", default="auto", choices=["auto", "on", "off"], help="rewrite a 160 MiB buffer between timed steps; auto: on for c1, whose working set fits the 126 MB L2")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="regular steps as one CUDA-graph launch (xsb_step_capture_*); auto: on for c1 (launch-bound), one GPU only")
    ap.add_argument("--no-mixed", action="store_true", help="skip the extra mixed-precision measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.flush_l2 = args.flush_l2 == "on" or (args.flush_l2 == "auto" and args.workload == "c1")
    args.graph = args.graph == "on" or (args.graph == "auto" and args.workload == "c1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_xsb(args)


if __name__ == "__main__":
    main()
