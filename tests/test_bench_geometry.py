"""bench.py's domain geometry for every rank grid the driver launches (N = 1, 2, 4, 8; weak and strong scaling): the
bricks must be made of whole cells of one common size >= rc + skin and tile the global cell grid."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402


class Args:
    cells, total_cells = 79, 160


@pytest.mark.parametrize("scaling", ["weak", "strong"])
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_bricks_are_whole_cells_of_one_size(scaling, n):
    a = Args(); a.scaling = scaling
    rd = bench.rank_dims(n)
    uc = bench.brick_cells(a, n)
    brick = np.asarray(uc, dtype=np.float64) * bench.A_CU
    cell, ncb3, gcells = bench.domain_cells(scaling, brick, rd)
    assert cell >= bench.RCUT + bench.SKIN
    for ax in range(3):
        assert abs(ncb3[ax] * cell - brick[ax]) < 1e-9 * brick[ax]          # a brick is a whole number of cells
        assert gcells[ax] == ncb3[ax] * rd[ax]                                 # the bricks tile the global grid
    atoms = 4 * uc[0] * uc[1] * uc[2] * n
    assert atoms == (4 * 160 ** 3 if scaling == "strong" else 4 * 79 ** 3 * n)
    cfg = bench.workload_config(type("A", (), dict(cells=79, total_cells=160, scaling=scaling, rebuild_every=20))(), n)
    assert str(atoms) in cfg["workload"]
