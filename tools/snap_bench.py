#!/usr/bin/env python
"""quick timing of snap_force on config C3 (BCC Ta, 2J=8, ~500k atoms); not the contract bench (bench.py)"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import exastamp_b200 as xsb
from helpers import EV, lattice

def main(nc=63, twoj=8, reps=5):
    a = 3.316
    pos, typ, box = lattice("BCC", nc, a, 0.05, seed=1)
    rcut = 4.7; skin = 1.0   # rcutfac 4.7 * 2*0.5
    ncell = int(box[0] // (rcut + skin)); cell = box[0] / ncell
    ctx = xsb.Context(0)
    ctx.grid_set(xsb.make_grid([ncell + 2] * 3, 1, cell, [-cell] * 3))
    ctx.particles_assign(pos[:, 0], pos[:, 1], pos[:, 2], None, None, None, typ)
    ctx.set_domain([ncell] * 3); ctx.ghost_comm_scheme()
    ncoef = xsb.load_library().xsb_snap_ncoeff(twoj)
    beta = np.random.default_rng(1).normal(0, 1, (1, ncoef + 1)) * 1e-3 * EV
    ctx.snap_set(twoj, rcut, [0.5], [1.0], beta)
    ctx.chunk_neighbors(rcut + skin)
    ctx.zero_force_energy(ghost=True)
    ctx.snap_force(xsb.FLAG_ENERGY); ctx.sync()
    if os.environ.get("XSB_SNAP_CLOCKS"):
        import ctypes as C
        L = xsb.load_library(); L.xsbdbg_snap_clocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.xsbdbg_snap_clocks(ctx.h, 1, None)
        ctx.snap_force(xsb.FLAG_ENERGY); ctx.sync()
        out = (C.c_uint64 * 8)(); L.xsbdbg_snap_clocks(ctx.h, 0, out)
        tot = float(sum(out)) or 1.0
        print("phase cycles: " + ", ".join("%s %.1f%%" % (n, 100 * v / tot) for n, v in zip(["init+filter", "sweepU", "mirror", "Y", "energy", "sweep_dU+force"], out)))
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.snap_force(xsb.FLAG_ENERGY)
    ctx.sync(); dt = (time.perf_counter() - t0) / reps
    ms, cnt = ctx.profile_read()["snap"]
    tot, mx = ctx.chunk_neighbors_stats()
    print("atoms %d (with ghosts %d) list %.1f/atom  snap_force %.2f ms/call (events %.2f)  %.3e atom-steps/s" % (ctx.n_own, ctx.n, tot / ctx.n, dt * 1e3, ms / cnt, ctx.n_own / dt))

if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
