"""CPU replays (numpy) of the arithmetic two SNAP kernels of csrc/xsb_snap.cu use, written before the kernels went to the GPU:

* snap_sweep_rev (reverse-mode force sweep): the adjoint recursion over the levels, including the mailbox that carries the
  adjoint of a row's birth back to the row it was mirrored from, gives dG/da, dG/db of G(a, b) = sum w Re(conj(Y) u) --
  checked against central finite differences of the forward recursion (the reference computes the same derivative with one
  chain per Cartesian direction: snaplmp/snap_force_op.h:250-320, compute_duidrj + compute_deidrj).
* snap_y_block (register-blocked compute_yi): work items = blocks of consecutive ma, loop over ma2 with a sliding window of
  u1 elements, coefficients re-tabulated as D[ma2][m]; replayed index for index (window rotation, clamped loads, table
  offsets) against the idxz formulation of compute_yi the oracle restates.

No GPU, no library: this pins the index algebra, the GPU parity tests pin the kernels."""
import math

import numpy as np
import pytest

ROOTPQ = np.zeros((10, 10))
for _a in range(1, 10):
    for _b in range(1, 10):
        ROOTPQ[_a][_b] = math.sqrt(_a / _b)


def idxu_blocks(tj):
    out, n = [], 0
    for j in range(tj + 1):
        out.append(n); n += (j + 1) ** 2
    return out, n


def seed_w(J, mb, ma):
    if 2 * mb == J:
        return 1.0 if ma < mb else (0.5 if ma == mb else 0.0)
    return 1.0


def forward(tj, Y, blk, a, b, hist=None):
    rows = {0: np.array([1 + 0j])}
    G = 0.5 * Y[0].real
    for J in range(1, tj + 1):
        new_rows = {}
        for mb in range(J // 2 + 1):
            if 2 * mb == J:      # birth: mirror of row mb-1 of level J-1
                src = rows[mb - 1]
                old = np.array([(-1.0 if ((mb - 1 + (J - 1 - ma)) & 1) else 1.0) * np.conj(src[J - 1 - ma]) for ma in range(J)])
            else:
                old = rows[mb]
            if hist is not None:
                hist[(J, mb)] = old.copy()
            new = np.zeros(J + 1, complex)
            for ma in range(J + 1):
                if ma < J:
                    new[ma] += ROOTPQ[J - ma][J - mb] * np.conj(a) * old[ma]
                if ma > 0:
                    new[ma] -= ROOTPQ[ma][J - mb] * np.conj(b) * old[ma - 1]
            new_rows[mb] = new
        rows = new_rows
        for mb in range(J // 2 + 1):
            base = blk[J] + (J + 1) * mb
            for ma in range(J + 1):
                G += seed_w(J, mb, ma) * (rows[mb][ma] * np.conj(Y[base + ma])).real
    return G


def backward(tj, Y, blk, a, b, hist):
    """the way down of snap_sweep_rev: ubar rows, abar / bbar accumulation, mirrored adjoints through the mailbox"""
    abar = bbar = 0j
    ub, mbox = {}, {}
    for J in range(tj, 0, -1):
        for mb in range(J // 2 + 1):
            base = blk[J] + (J + 1) * mb
            nb = np.array([seed_w(J, mb, ma) * Y[base + ma] for ma in range(J + 1)])
            if mb in ub:
                nb = nb + ub[mb]
            if (J, mb) in mbox:
                nb = nb + mbox.pop((J, mb))
            old = hist[(J, mb)]
            ob = np.zeros(J, complex)
            for m in range(J):
                qa, qb = ROOTPQ[J - m][J - mb], ROOTPQ[m + 1][J - mb]
                p, s = qa * nb[m], qb * nb[m + 1]
                ob[m] = a * p - b * s
                abar += (p.real * old[m].real + p.imag * old[m].imag) + 1j * (p.real * old[m].imag - p.imag * old[m].real)
                bbar -= (s.real * old[m].real + s.imag * old[m].imag) + 1j * (s.real * old[m].imag - s.imag * old[m].real)
            if 2 * mb == J:
                m = np.zeros(J, complex)
                for ma in range(J):
                    mp = J - 1 - ma
                    m[mp] = (-1.0 if ((mb - 1 + mp) & 1) else 1.0) * np.conj(ob[ma])
                mbox[(J - 1, mb - 1)] = m
                ub.pop(mb, None)
            else:
                ub[mb] = ob
    assert not mbox
    return abar, bbar


@pytest.mark.parametrize("tj", [1, 2, 3, 5, 8])
def test_reverse_mode_sweep_equals_finite_differences(tj):
    rng = np.random.default_rng(tj)
    blk, n = idxu_blocks(tj)
    Y = rng.normal(size=n) + 1j * rng.normal(size=n)
    a, b = complex(0.3, -0.5), complex(0.6, 0.2)
    hist = {}
    forward(tj, Y, blk, a, b, hist)
    abar, bbar = backward(tj, Y, blk, a, b, hist)
    e = 1e-6
    fd = [(forward(tj, Y, blk, a + e, b) - forward(tj, Y, blk, a - e, b)) / (2 * e),
          (forward(tj, Y, blk, a + 1j * e, b) - forward(tj, Y, blk, a - 1j * e, b)) / (2 * e),
          (forward(tj, Y, blk, a, b + e) - forward(tj, Y, blk, a, b - e)) / (2 * e),
          (forward(tj, Y, blk, a, b + 1j * e) - forward(tj, Y, blk, a, b - 1j * e)) / (2 * e)]
    got = [abar.real, abar.imag, bbar.real, bbar.imag]
    scale = max(1.0, max(abs(x) for x in fd))
    assert max(abs(g - f) for g, f in zip(got, fd)) < 1e-7 * scale


def history_slots(tj):
    """SnapHist<TJ>::total(): rows keep levels max(1, 2mb) .. TJ-1, level J holds J elements"""
    S = lambda J: J * (J - 1) // 2
    return sum(max(0, S(tj) - S(max(1, 2 * mb))) for mb in range(tj // 2 + 1))


def test_history_size_of_the_force_kernel():
    assert history_slots(8) == 90 and history_slots(1) == 0 and history_slots(2) == 1


def cg_tables(tj):
    f = math.factorial
    tri = [(j1, j2, j) for j1 in range(tj + 1) for j2 in range(j1 + 1) for j in range(j1 - j2, min(tj, j1 + j2) + 1, 2)]
    cg = {}
    for (j1, j2, j) in tri:
        t = np.zeros((j1 + 1, j2 + 1))
        for m1 in range(j1 + 1):
            for m2 in range(j2 + 1):
                aa2, bb2 = 2 * m1 - j1, 2 * m2 - j2
                m = (aa2 + bb2 + j) // 2
                if m < 0 or m > j:
                    continue
                zlo = max(0, max(-(j - j2 + aa2) // 2, -(j - j1 - bb2) // 2)); zhi = min((j1 + j2 - j) // 2, min((j1 - aa2) // 2, (j2 + bb2) // 2))
                sm = sum((-1.0 if z & 1 else 1.0) / (f(z) * f((j1 + j2 - j) // 2 - z) * f((j1 - aa2) // 2 - z) * f((j2 + bb2) // 2 - z) * f((j - j2 + aa2) // 2 + z) * f((j - j1 - bb2) // 2 + z))
                         for z in range(zlo, zhi + 1))
                cc2 = 2 * m - j
                dcg = math.sqrt(f((j1 + j2 - j) // 2) * f((j1 - j2 + j) // 2) * f((-j1 + j2 + j) // 2) / f((j1 + j2 + j) // 2 + 1))
                sf = math.sqrt(f((j1 + aa2) // 2) * f((j1 - aa2) // 2) * f((j2 + bb2) // 2) * f((j2 - bb2) // 2) * f((j + cc2) // 2) * f((j - cc2) // 2) * (j + 1))
                t[m1, m2] = sm * dcg * sf
        cg[(j1, j2, j)] = t
    return tri, cg


@pytest.mark.parametrize("tj", [2, 5, 8])
def test_blocked_compute_yi_equals_the_idxz_formulation(tj):
    rng = np.random.default_rng(10 + tj)
    blk, n = idxu_blocks(tj)
    U = rng.normal(size=n) + 1j * rng.normal(size=n)
    tri, cg = cg_tables(tj)
    beta = {t: rng.normal() for t in tri}
    # compute_yi as the oracle restates it (one idxz entry per (triple, mb, ma))
    Yref = np.zeros(n, complex)
    for (j1, j2, j) in tri:
        c = cg[(j1, j2, j)]
        for mb in range(j // 2 + 1):
            for ma in range(j + 1):
                ma1min = max(0, (2 * ma - j - j2 + j1) // 2); ma2max = (2 * ma - j - (2 * ma1min - j1) + j2) // 2; na = min(j1, (2 * ma - j + j2 + j1) // 2) - ma1min + 1
                mb1min = max(0, (2 * mb - j - j2 + j1) // 2); mb2max = (2 * mb - j - (2 * mb1min - j1) + j2) // 2; nb = min(j1, (2 * mb - j + j2 + j1) // 2) - mb1min + 1
                z = 0
                for ib in range(nb):
                    sm = sum(c[ma1min + ia, ma2max - ia] * U[blk[j1] + (j1 + 1) * (mb1min + ib) + ma1min + ia] * U[blk[j2] + (j2 + 1) * (mb2max - ib) + ma2max - ia] for ia in range(na))
                    z += c[mb1min + ib, mb2max - ib] * sm
                Yref[blk[j] + (j + 1) * mb + ma] += beta[(j1, j2, j)] * z
    # the kernel's tables (xsb_snap_set) and loops (snap_y_block)
    tl, first = [], []
    for j in range(tj + 1):
        first.append(len(tl))
        tl += [(j1, j2, j) for j1 in range(tj + 1) for j2 in range(j1 + 1) if j1 - j2 <= j <= min(tj, j1 + j2) and ((j1 + j2 - j) & 1) == 0]
    first.append(len(tl))
    D, y2tri = [], []
    for (j1, j2, j) in tl:
        P, cgp, sh = (j + 2) & ~1, len(D), (j1 + j2 - j) // 2
        for m2 in range(j2 + 1):
            for m in range(P):
                m1 = m + sh - m2
                D.append(cg[(j1, j2, j)][m1, m2] if (m <= j and 0 <= m1 <= j1) else 0.0)
        y2tri.append((j1, j2, sh, P, cgp))
    D = np.array(D + [0.0, 0.0])
    assert all(t[4] % 2 == 0 for t in y2tri)                 # aligned pair loads
    Y = np.full(n, np.nan, complex)
    for j in range(2, tj + 1, 2):
        for k in range(j // 2):
            Y[blk[j] + (j + 1) * (j // 2) + j // 2 + 1 + k] = 0
    for j in range(tj + 1):
        for mb in range(j // 2 + 1):
            nn = mb + 1 if 2 * mb == j else j + 1
            ma0 = 0
            for W in {6: [4, 2], 7: [4, 3], 8: [4, 4], 9: [4, 5]}.get(nn, [nn]):
                assert ma0 % 2 == 0
                yr = np.zeros(W, complex)
                for it in range(first[j], first[j + 1]):
                    j1, j2, s, P, cgp = y2tri[it]
                    mb1min = max(0, mb + s - j2); nb = min(j1, mb + s) - mb1min + 1
                    lo2 = max(0, ma0 + s - j1); n_it = min(j2, ma0 + W - 1 + s) - lo2 + 1
                    t0 = ma0 + s - lo2
                    assert n_it >= 1 and nb >= 1 and t0 <= j1
                    for ib in range(nb):
                        r1 = mb1min + ib; r2 = mb + s - r1
                        cb = beta[tl[it]] * D[cgp + r2 * P + mb]
                        u1row = blk[j1] + r1 * (j1 + 1); u2i = blk[j2] + r2 * (j2 + 1) + lo2
                        t, cp = t0, cgp + lo2 * P + ma0
                        w = [U[u1row + min(t + d, j1)] for d in range(W)]
                        sr = np.zeros(W, complex)
                        for k in range(0, n_it, W):
                            for p in range(W):
                                if k + p < n_it:
                                    u2 = U[u2i]; u2i += 1
                                    for d in range(W):
                                        assert cgp <= cp + d < cgp + (j2 + 1) * P
                                        sr[d] += D[cp + d] * w[(d - p + W) % W] * u2
                                    cp += P; t -= 1
                                    w[(W - 1 - p) % W] = U[u1row + max(t, 0)]
                        yr += cb * sr
                for d in range(W):
                    Y[blk[j] + (j + 1) * mb + ma0 + d] = yr[d]
                ma0 += W
    for j in range(tj + 1):
        for mb in range(j // 2 + 1):
            for ma in range(j + 1):
                i = blk[j] + (j + 1) * mb + ma
                if 2 * mb == j and ma > mb:
                    assert Y[i] == 0
                else:
                    assert abs(Y[i] - Yref[i]) < 1e-12 * max(1.0, np.abs(Yref).max())
