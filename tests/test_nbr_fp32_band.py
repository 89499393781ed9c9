"""The list build tests candidates in FP32 and lets an exact FP64 test decide inside a guard band around the cut-off
(xsb_nbr.cu, nbr_count_kernel / nbr_fp32_band).  The list stays bit-identical to an all-FP64 sweep as long as the band
bounds |d2_fp32 - d2_exact| for every candidate near the cut-off.  This CPU test replays the kernel's FP32 arithmetic in
numpy (float32 coordinates relative to a tile atom, float32 cell matrix, float32 products and sums) on candidates placed
within 1 % of the cut-off at the far corners of a tile's stage, and compares with the exact distance."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import exastamp_b200 as xsb  # noqa: E402


def band(cell, X, tx, R, dist):
    L = xsb.load_library()
    L.xsbdbg_nbr_fp32_band.restype = C.c_double
    L.xsbdbg_nbr_fp32_band.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_double]
    x = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
    return L.xsbdbg_nbr_fp32_band(float(cell), None if x is None else x.ctypes.data_as(C.c_void_p), int(tx), int(R), float(dist))


CASES = [  # cell size, cut-off, tile width, search range, cell matrix, box offset of the tile (large absolute coordinates)
    (8.365, 8.29, 1, 1, None, 280.0),
    (9.30, 9.0, 3, 1, None, 395.0),
    (7.93, 7.6825, 1, 1, np.array([[1.0, 0.01, 0.005], [0.0, 1.0, 0.01], [0.0, 0.0, 1.0]]), 450.0),
    (5.2, 9.0, 2, 2, np.array([[1.02, 0.03, 0.01], [0.0, 0.98, 0.02], [0.0, 0.0, 1.01]]), 150.0),
]


@pytest.mark.parametrize("cell,dist,tx,R,X,off", CASES)
def test_fp32_distance_error_stays_inside_the_guard_band(cell, dist, tx, R, X, off):
    b = band(cell, X, tx, R, dist)
    assert 0 < b < 0.02 * dist * dist                                  # a thin shell, not a second cut-off
    rng = np.random.default_rng(11)
    n = 200000
    M = np.eye(3) if X is None else X
    Minv = np.linalg.inv(M)
    # tile atom (origin of the FP32 coordinates) and central atoms anywhere in the tile; candidates at physical
    # distance within 1 % of the cut-off in random directions (grid-space offset = M^-1 * physical offset)
    o = off + rng.random(3) * cell
    a = o + (rng.random((n, 3)) * np.array([tx, 1, 1]) - np.array([0.0, 0.5, 0.5])) * cell
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    rphys = dist * (1.0 + (rng.random(n) - 0.5) * 0.02)
    c = a + (u * rphys[:, None]) @ Minv.T
    # exact (FP64) distance in the reference's formulation
    d = (c - a) @ M.T
    d2 = (d * d).sum(axis=1)
    # the kernel's FP32 replay
    af = (a - o).astype(np.float32); cf = (c - o).astype(np.float32)
    df = cf - af
    Mf = M.astype(np.float32)
    if X is not None:
        df = np.stack([Mf[k, 0] * df[:, 0] + Mf[k, 1] * df[:, 1] + Mf[k, 2] * df[:, 2] for k in range(3)], axis=1).astype(np.float32)
    d2f = (df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1] + df[:, 2] * df[:, 2]).astype(np.float32)
    err = np.abs(d2f.astype(np.float64) - d2).max()
    assert err < 0.5 * b, (err, b)                                     # factor-2 margin on top of the analytic bound
    # and the classification rule of the kernel is exact wherever it does not defer to FP64
    d2max = np.float32(dist * dist); bf = np.float32(b)
    sure_in = d2f < d2max - bf
    sure_out = d2f > d2max + bf
    assert np.all(d2[sure_in] < dist * dist) and np.all(d2[sure_out] >= dist * dist)
    assert (sure_in | sure_out).mean() > 0.5                           # most candidates this close are still decided in FP32
