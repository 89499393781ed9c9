"""Precision model of the mixed-precision EAM passes (XSB_FLAG_MIXED, xsb_eam.cu EamRhoTileOp32 / EamForceTileOp32),
replayed in numpy on the bench's table (setfl header of Cu.eam.alloy: nr 5000, rc 7.29): the FP32 rows must be the cubic
coefficients {c6, c5, c4, c3} rounded once from FP64.  Re-deriving c4, c3 from FP32 knots {f[m], c5[m], f[m+1], c5[m+1]}
-- the layout the FP64 passes use -- cancels (f[m+1] - f[m] ~ 1e-3 f) and misses the 1e-5 bar by an order of magnitude;
this test pins both statements so the table layout cannot silently regress."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import SC_CU  # noqa: E402

NR, RC = 5000, 7.29


def spline_rows(f, delta):
    """LAMMPS pair_eam interpolate() (eam_alloy.cpp:29-58), rows 1..n, columns c0..c6"""
    n = len(f) - 1
    s = np.zeros((n + 1, 7))
    s[1:, 6] = f[1:]
    s[1, 5] = s[2, 6] - s[1, 6]; s[2, 5] = 0.5 * (s[3, 6] - s[1, 6])
    s[n - 1, 5] = 0.5 * (s[n, 6] - s[n - 2, 6]); s[n, 5] = s[n, 6] - s[n - 1, 6]
    m = np.arange(3, n - 1)
    s[m, 5] = ((s[m - 2, 6] - s[m + 2, 6]) + 8.0 * (s[m + 1, 6] - s[m - 1, 6])) / 12.0
    m = np.arange(1, n)
    s[m, 4] = 3.0 * (s[m + 1, 6] - s[m, 6]) - 2.0 * s[m, 5] - s[m + 1, 5]
    s[m, 3] = s[m, 5] + s[m + 1, 5] - 2.0 * (s[m + 1, 6] - s[m, 6])
    s[:, 2] = s[:, 5] / delta; s[:, 1] = 2.0 * s[:, 4] / delta; s[:, 0] = 3.0 * s[:, 3] / delta
    return s


def test_fp32_cubic_rows_meet_the_mixed_bar_and_fp32_knots_do_not():
    dr = RC / NR
    r_tab = np.maximum(np.arange(NR) * dr, 0.6)
    f = np.concatenate([[0.0], (SC_CU["a"] / r_tab) ** SC_CU["m"]])           # rhor table of the bench potential, 1-based
    S = spline_rows(f, dr)
    rng = np.random.default_rng(2)
    r = rng.uniform(2.2, RC * 0.999, 400000)
    rdr = 1.0 / dr
    p64 = r * rdr + 1.0
    m = np.minimum(p64.astype(np.int64), NR - 1)
    p64 = np.minimum(p64 - m, 1.0)
    val64 = ((S[m, 3] * p64 + S[m, 4]) * p64 + S[m, 5]) * p64 + S[m, 6]
    der64 = (S[m, 0] * p64 + S[m, 1]) * p64 + S[m, 2]
    # the kernel's FP32 path: r from an FP32 square root of an FP32 copy of d2, p and the polynomial in FP32
    f32 = np.float32
    rf = np.sqrt((r * r).astype(f32)).astype(f32)
    pf = (rf * f32(rdr) + f32(1.0)).astype(f32)
    mf = np.minimum(pf.astype(np.int64), NR - 1)
    pf = np.minimum(pf - mf.astype(f32), f32(1.0)).astype(f32)

    def evaluate(c6, c5, c4, c3):
        v = (((c3 * pf + c4) * pf + c5) * pf + c6).astype(f32)
        d = (((f32(3.0) * c3 * pf + f32(2.0) * c4) * pf + c5) * f32(rdr)).astype(f32)
        return v.astype(np.float64), d.astype(np.float64)

    # (a) rows of cubic coefficients rounded once from FP64 (what xsb_eam_alloy_set uploads)
    va, da = evaluate(S[mf, 6].astype(f32), S[mf, 5].astype(f32), S[mf, 4].astype(f32), S[mf, 3].astype(f32))
    # (b) c4, c3 re-derived in FP32 from FP32 knots
    k0f, k0c, k1f, k1c = S[mf, 6].astype(f32), S[mf, 5].astype(f32), S[np.minimum(mf + 1, NR), 6].astype(f32), S[np.minimum(mf + 1, NR), 5].astype(f32)
    df = (k1f - k0f).astype(f32)
    vb, db = evaluate(k0f, k0c, (f32(3.0) * df - f32(2.0) * k0c - k1c).astype(f32), (k0c + k1c - f32(2.0) * df).astype(f32))
    # a row mismatch (mf != m at an interval boundary) is legitimate: the cubic pieces join continuously
    rel = lambda x, y: np.abs(x - y) / np.abs(y)
    assert rel(va, val64).max() < 2e-6 and rel(da, der64).max() < 1e-5
    assert np.sqrt(np.mean(rel(da, der64) ** 2)) < 2e-6                      # what a sum over ~140 neighbours sees
    assert rel(db, der64).max() > 5e-5, "FP32 knots unexpectedly accurate: revisit the table layout comment in xsb_eam.cu"
