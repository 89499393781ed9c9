#!/usr/bin/env python
"""Side workloads of BASELINE.json (not the judged bench line, that is bench.py = configs[1]):
   configs[0]  Lennard-Jones FCC argon 32^3 unit cells (131 072 atoms), NVE Verlet       python tools/side_bench.py lj
   configs[2]  SNAP tantalum BCC 2J=8, 63^3 unit cells (500 094 atoms), NVE Verlet       python tools/side_bench.py snap
   configs[4]  two-species random FCC alloy, 126^3 unit cells (8 001 504 atoms), tabulated eam/alloy, cell matrix
               (upper-triangular xform) changing every step as under NPT, rebuild on the displacement trigger
                                                                                           python tools/side_bench.py c5 STEPS WARMUP VIRIAL
Same step structure as bench.py (push_f_v_r, push_f_v, displacement trigger, ghost_update_r or rebuild, forces,
force_to_accel, push_f_v), device-timed with CUDA events on the context stream; prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import exastamp_b200 as xsb  # noqa: E402
from helpers import EV, lattice  # noqa: E402

KB = 8.617333262e-5 * EV
W = {"lj": dict(structure="FCC", a=5.0, cells=32, rcut=8.0, skin=1.0, mass=39.948, noise=0.1, label="configs[0] LJ Ar FCC 32^3 (131072 atoms) NVE"),
     "lj2m": dict(structure="FCC", a=5.0, cells=80, rcut=8.0, skin=1.0, mass=39.948, noise=0.1, label="LJ Ar FCC 80^3 (2048000 atoms) NVE (configs[0] potential at the C2 size)"),
     "c2j": dict(structure="FCC", a=3.6, cells=79, rcut=5.8, skin=1.0, mass=63.546, noise=0.1,
                 label="configs[1] analytic variant: johnson_force (Zhou/Johnson/Wadley Cu set) FCC 79^3 (1972156 atoms) NVE"),
     "c5": dict(structure="FCC", a=3.8, cells=126, rcut=6.6825, skin=1.0, mass=45.0, noise=0.08,
                label="configs[4] two-species eam/alloy FCC 126^3 (8001504 atoms), NPT-like time-varying xform, periodic rebuild"),
     "snap": dict(structure="BCC", a=3.316, cells=63, rcut=4.7, skin=1.0, mass=180.95, noise=0.05, label="configs[2] SNAP Ta BCC 2J=8 (500094 atoms) NVE")}


def main(which, steps=100, warmup=10, mixed=0, cells=0):
    w = dict(W[which])
    if cells:
        w["cells"] = cells; w["label"] += " [cells per axis overridden: %d]" % cells
    pos, typ, box = lattice(w["structure"], w["cells"], w["a"], w["noise"], seed=1)
    vel = np.random.default_rng(2).normal(0.0, np.sqrt(KB * 300.0 / w["mass"]), pos.shape); vel -= vel.mean(axis=0)
    X0 = None
    if which == "c5":
        import tempfile
        from helpers import SC_CU, SC_XX, write_setfl
        typ = (np.random.default_rng(3).random(len(pos)) < 0.5).astype(np.uint8)
        X0 = np.array([[1.0, 0.01, 0.005], [0.0, 1.0, 0.01], [0.0, 0.0, 1.0]])
        setfl = write_setfl(os.path.join(tempfile.mkdtemp(), "ab.eam.alloy"), [SC_CU, SC_XX], nrho=10000, drho=0.02, nr=5000, rc=w["rcut"])
    nc = int(box[0] // ((w["rcut"] + w["skin"]) * (1.03 if X0 is not None else 1.0))); cell = box[0] / nc
    ctx = xsb.Context(0)
    ctx.grid_set(xsb.make_grid([nc + 2] * 3, 1, cell, [-cell] * 3, X0))
    ctx.particles_assign(pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2], typ)
    ctx.set_domain([nc] * 3); ctx.ghost_comm_scheme()
    if which == "snap":
        ncoef = xsb.load_library().xsb_snap_ncoeff(8)
        beta = np.random.default_rng(1).normal(0, 1, (1, ncoef + 1)) * 1e-3 * EV
        ctx.snap_set(8, w["rcut"], [0.5], [1.0], beta)
    if which == "c5":
        ctx.eam_alloy_load(setfl)
    POS = [xsb.F_RX, xsb.F_RY, xsb.F_RZ]
    state = {"since": 0, "rebuilds": 0, "step": 0}

    def rebuild(first=False):
        if not first:
            ctx.particles_rebin(); ctx.ghost_comm_scheme()
        ctx.chunk_neighbors(w["rcut"] + w["skin"]); ctx.backup_r()
        state["since"] = 0; state["rebuilds"] += 1

    def forces():
        if which == "c5":
            vf = (xsb.EAM_EFLAG, xsb.FLAG_VIRIAL) if mixed else (0, 0)       # 3rd argument: energies + virial every step
            ctx.zero_force_energy()
            ctx.eam_alloy_force(w["rcut"], xsb.EAM_RHO | xsb.EAM_RHO2EMB | vf[0], vf[1])
            ctx.ghost_update([xsb.F_RHO_DEMB])
            ctx.eam_alloy_force(w["rcut"], xsb.EAM_FORCE | vf[0], vf[1])
        elif which == "c2j":
            from helpers import johnson_params
            ctx.zero_force_energy()
            ctx.eam_johnson_force(johnson_params(), w["rcut"], 1)                 # johnson_emb: rho -> F'(rho)
            ctx.ghost_update([xsb.F_RHO_DEMB])
            ctx.eam_johnson_force(johnson_params(), w["rcut"], 4)                 # johnson_force_reuse_emb
        elif which != "snap":
            ctx.zero_force_energy()
            ctx.pair_force([0.0104 * EV, 3.4], w["rcut"], xsb.FLAG_MIXED if mixed else 0)
        else:
            ctx.zero_force_energy(ghost=True)
            ctx.snap_force(0)
            ctx.ghost_reduce_add([xsb.F_FX, xsb.F_FY, xsb.F_FZ])

    def step():
        over, _ = ctx.verlet_boundary([w["mass"]] * 2, 1e-3, 0.5 * w["skin"])     # force_to_accel, push_f_v | push_f_v_r, push_f_v, particle_displ_over
        state["since"] += 1; state["step"] += 1
        if X0 is not None:                                # barostat-like drift of the cell matrix, 2e-6 per step
            ctx.grid_set_xform(X0 * (1.0 + 2e-6 * state["step"]))
        if over or state["since"] >= 20:
            rebuild()
        else:
            ctx.ghost_update(POS)
        forces()

    rebuild(first=True); forces()
    for _ in range(warmup):
        step()
    ctx.sync(); ctx.profile_enable(True)
    r0 = state["rebuilds"]; l0 = ctx.launches
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop_ms()
    prof = ctx.profile_read()
    tot, mx = ctx.chunk_neighbors_stats()
    print(json.dumps({"workload": w["label"], "metric": "atom-timesteps/s (neighbor+force)", "value": ctx.n_own * steps / (ms * 1e-3), "ms_per_step": ms / steps,
                      "atoms": ctx.n_own, "steps": steps, "dtype": "f64" if which == "c5" or not mixed else "f32 pair math / f64 accumulation", "energy_virial_every_step": bool(mixed) if which == "c5" else None, "list_entries_per_atom": tot / max(1, ctx.n),
                      "rebuilds": state["rebuilds"] - r0, "gpu_launches": ctx.launches - l0,
                      "breakdown_ms_per_call": {k: round(v[0] / v[1], 4) for k, v in prof.items() if v[1]}}))


if __name__ == "__main__":
    main(sys.argv[1], *[int(a) for a in sys.argv[2:]])
