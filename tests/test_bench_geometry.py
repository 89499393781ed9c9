"""bench.py's domain geometry for every workload and every rank grid the driver launches (N = 1, 2, 4, 8): the bricks must
be made of whole cells of one common size >= rc + skin and tile the global cell grid; the CPU arm runs end to end on a
small sample of every workload (same operators as the GPU arm, oracle restatement)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402


def args_for(W, **kw):
    a = type("A", (), dict(cells=0, total_cells=0, scaling="strong" if W.strong_total else "weak", rebuild_every=20, flush_l2=False))()
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("wl", sorted(bench.WORKLOADS))
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_bricks_are_whole_cells_of_one_size(wl, n):
    W = bench.WORKLOADS[wl]()
    a = args_for(W)
    rd = bench.rank_dims(n)
    uc = bench.brick_cells(W, a, n)
    brick = np.asarray(uc, dtype=np.float64) * W.a
    cell, ncb3, gcells = bench.domain_cells(W, a.scaling, brick, rd)
    assert cell >= (W.rcut + W.skin) * W.cell_slack
    for ax in range(3):
        assert abs(ncb3[ax] * cell - brick[ax]) < 1e-9 * brick[ax]          # a brick is a whole number of cells
        assert gcells[ax] == ncb3[ax] * rd[ax]                                 # the bricks tile the global grid
    per_cell = 2 if W.structure == "BCC" else 4
    atoms = per_cell * uc[0] * uc[1] * uc[2] * n
    want = {"c1": 131072 * n, "c2": 1972156 * n, "c2j": 1972156 * n, "c3": 500094 * n, "c4": 16384000, "c5": 8001504 * n}[wl]
    assert atoms == want
    assert str(atoms) in bench.workload_config(W, a, n)["workload"]


@pytest.mark.parametrize("wl,cells", [("c1", 6), ("c2", 6), ("c3", 6), ("c5", 6)])
def test_cpu_arm_runs_every_workload_on_a_small_sample(wl, cells):
    W = bench.WORKLOADS[wl]()
    v, info = bench.cpu_reference_run(W, cells, 1, 0, 20)
    assert v > 0 and info["atoms"] == (2 if W.structure == "BCC" else 4) * cells ** 3 and info["threads"] >= 1


def test_cpu_arm_times_the_full_configuration_unless_it_cannot_fit_its_budget():
    """--impl reference: the stated configuration at the driver's step counts (N = 1 and N = 8), SNAP and over-long runs on a
    labelled cube that is smaller than the configuration"""
    c2 = bench.WORKLOADS["c2"]()
    assert bench.cpu_arm_sample_cells(c2, [79, 79, 79], 20, 5, 16) == 0
    assert bench.cpu_arm_sample_cells(c2, [158, 158, 158], 20, 5, 32) == 0            # 15.8 M atoms, 27 steps: ~75 s on 32 cores
    sc = bench.cpu_arm_sample_cells(c2, [158, 158, 158], 100, 10, 16)
    assert 16 <= sc < 158
    assert 4 * sc ** 3 * 112 / (1.8e5 * 16) <= 240.0 * 1.01
    assert bench.cpu_arm_sample_cells(bench.WORKLOADS["c3"](), [63, 63, 63], 3, 1, 64) == 24
    assert bench.cpu_arm_sample_cells(c2, [79, 79, 79], 20, 5, 16, forced=40) == 40
    assert bench.cpu_arm_sample_cells(c2, [12, 12, 12], 20, 5, 16, forced=40) == 12
