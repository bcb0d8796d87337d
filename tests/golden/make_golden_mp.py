"""High-precision (mpmath, 60 digits) known answers for the orders the reference cannot compute at bw = 2048
-> tests/golden/mp_high_orders.npz.

The reference's Pmm_L2 (src/legendre_polynomials/pmm.c:21-33) overflows for m >= 2044 (the running product of
sqrt((m - i/2)/(m - i)) reaches inf before pow(2, -m/2) brings it back), so FSTSemiMemo returns NaN for the orders
|m| >= 2044 and the committed reference samples (oracle_vectors_large.npz) leave them unpinned.  The definitions
themselves are finite there; this script evaluates them exactly:

  P~_m^m(theta) = (-1)^m sqrt(m + 1/2) prod_{i<m} sqrt((m - i/2)/(m - i)) 2^(-m/2) sin^m(theta)       (pmm.c:21-33)
  P~_{l+1}^m    = L2_an(m,l) cos(theta) P~_l^m + L2_cn(m,l) P~_{l-1}^m                                   (l2_norms.c:16-38,
                                                                                         pml.c / cospml.c recurrences)
  weights w_j   = (2/bw) sin(theta_j) sum_{k<bw} sin((2k+1) theta_j)/(2k+1),  theta_j = (2j+1) pi/(4 bw)  (weights.c:32-47)
  DLT           : c_l  = sum_j w_j f_j P~_l^m(theta_j)                                       (naive.c:35-63, = DLTSemi)
  inverse DLT   : f_j  = sum_l c_l P~_l^m(theta_j)                                          (naive.c:81-100, = InvDLTSemi)

Before anything is written the mpmath evaluation is checked against the reference itself (oracle/_ref) where the
reference is finite: every order at bw = 24 (tables through the naive transform) and the orders 2040..2043 at bw = 2048
(DLTSemi / InvDLTSemi of the same seeded columns).  Run in the build container (needs /root/reference); ~2 minutes.
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

mp.mp.dps = 60


def nodes(bw):
    return [mp.pi * (2 * j + 1) / (4 * bw) for j in range(2 * bw)]


def weights(bw):
    """GenerateWeightsForDLT, first half (the plain weights)."""
    th = nodes(bw)
    out = []
    for t in th:
        s = mp.mpf(0)
        for k in range(bw):
            s += mp.sin((2 * k + 1) * t) / (2 * k + 1)
        out.append(2 * mp.sin(t) / bw * s)
    return out


def an(m, l):
    m, l = mp.mpf(m), mp.mpf(l)
    return mp.sqrt(((2 * l + 3) / (2 * l + 1)) * ((l - m + 1) / (l + m + 1))) * ((2 * l + 1) / (l - m + 1))


def cn(m, l):
    if l == 0:
        return mp.mpf(0)
    m, l = mp.mpf(m), mp.mpf(l)
    return -mp.sqrt(((2 * l + 3) / (2 * l - 1)) * ((l - m + 1) / (l + m + 1)) * ((l - m) / (l + m))) * ((l + m) / (l - m + 1))


def pmm_coeff(m):
    c = mp.sqrt(mp.mpf(m) + mp.mpf(1) / 2)
    for i in range(m):
        c *= mp.sqrt((m - mp.mpf(i) / 2) / (m - i))
    if m:
        c *= mp.power(2, -mp.mpf(m) / 2)
    return -c if m % 2 else c


def pml_rows(bw, m, th=None):
    """[(bw - m) rows][2 bw nodes] of P~_l^m(theta_j), l = m .. bw-1."""
    th = th or nodes(bw)
    c = pmm_coeff(m)
    prev = [mp.mpf(0)] * len(th)
    cur = [c * mp.sin(t) ** m for t in th]
    x = [mp.cos(t) for t in th]
    rows = [cur]
    for l in range(m, bw - 1):
        a, cc = an(m, l), cn(m, l)
        nxt = [a * x[j] * cur[j] + cc * prev[j] for j in range(len(th))]
        prev, cur = cur, nxt
        rows.append(cur)
    return rows


def mp_dlt(rows, w, data):
    return [sum(w[j] * mp.mpf(float(data[j])) * r[j] for j in range(len(w))) for r in rows]


def mp_inv_dlt(rows, coeffs):
    n = len(rows[0])
    return [sum(mp.mpf(float(c)) * r[j] for c, r in zip(coeffs, rows)) for j in range(n)]


def f64(v):
    return np.array([float(x) for x in v], dtype=np.float64)


def main():
    import oracle
    from oracle import _p

    # ---- 1. the mpmath evaluation against the reference where the reference is finite: every order at bw = 24
    bw = 24
    R = oracle.Oracle(bw, "ref")
    th, w = nodes(bw), weights(bw)
    wref = R.weights()[:2 * bw]
    assert np.abs(f64(w) - wref).max() < 1e-15, "weights"
    worst = 0.0
    rng = np.random.RandomState(24)
    for m in range(bw):
        rows = pml_rows(bw, m, th)
        data = rng.uniform(-1, 1, 2 * bw)
        got = np.zeros(bw - m)
        R.L.ref_dlt_semi(R.h, _p(data), m, _p(got))
        want = f64(mp_dlt(rows, w, data))
        worst = max(worst, np.abs(got - want).max() / max(1.0, np.abs(want).max()))
        co = rng.uniform(-1, 1, bw - m)
        back = np.zeros(2 * bw)
        R.L.ref_inv_dlt_semi(R.h, _p(co), m, _p(back))
        wantb = f64(mp_inv_dlt(rows, co))
        worst = max(worst, np.abs(back - wantb).max() / max(1.0, np.abs(wantb).max()))
    R.close()
    print(f"bw = 24: reference DLTSemi / InvDLTSemi vs mpmath, all orders: {worst:.2e}")
    assert worst < 1e-13

    # ---- 2. bw = 2048: orders 2040..2047 (2040..2043: the reference is still finite -- cross-checked against the
    # committed reference samples by tests/test_oracle.py; 2044..2047: mpmath is the only source)
    bw = 2048
    th, w = nodes(bw), weights(bw)
    out = {"bw": np.int64(bw), "orders": np.arange(2040, 2048, dtype=np.int64),
           "weights": f64(w), "dps": np.int64(mp.mp.dps)}
    # the seeded columns of make_golden_large.py (RandomState(7), drawn in DLT_ORDERS_2048 order) for m = 2042, 2043:
    # the reference's own DLTSemi / InvDLTSemi outputs for them are committed in oracle_vectors_large.npz
    large = np.load(os.path.join(HERE, "oracle_vectors_large.npz"))
    rng7 = np.random.RandomState(7)
    replay = {}
    for mm in [0, 1, 2, 511, 778, 1023, 1024, 1999, 2042, 2043]:
        replay[mm] = (rng7.uniform(-1, 1, 2 * bw), rng7.uniform(-1, 1, bw - mm))
    for m in range(2040, 2048):
        rows = pml_rows(bw, m, th)
        if m in (2042, 2043):  # mpmath vs the reference at bw = 2048, where the reference is still finite
            col, co7 = replay[m]
            fw, bk = f64(mp_dlt(rows, w, col)), f64(mp_inv_dlt(rows, co7))
            e1 = np.abs(fw - large[f"bw2048_dlt_m{m}"]).max() / np.abs(fw).max()
            e2 = np.abs(bk - large[f"bw2048_invdlt_m{m}"]).max() / np.abs(bk).max()
            print(f"m = {m}: mpmath vs committed reference DLTSemi {e1:.2e}, InvDLTSemi {e2:.2e}")
            assert e1 < 1e-11 and e2 < 1e-11
            out[f"crosscheck_m{m}_dlt"], out[f"crosscheck_m{m}_inv"] = fw, bk
        rng = np.random.RandomState(20480 + m)
        data = rng.uniform(-1, 1, 2 * bw)
        co = rng.uniform(-1, 1, bw - m)
        out[f"m{m}_data"] = data
        out[f"m{m}_dlt"] = f64(mp_dlt(rows, w, data))
        out[f"m{m}_coeffs"] = co
        out[f"m{m}_inv"] = f64(mp_inv_dlt(rows, co))
        out[f"m{m}_pmm"] = f64(rows[0])  # P~_m^m at the 2 bw nodes (what a finite Pmm_L2 must return)
        print(m, out[f"m{m}_dlt"][:2], np.abs(out[f"m{m}_pmm"]).max())
    np.savez_compressed(os.path.join(HERE, "mp_high_orders.npz"), **out)
    print("wrote mp_high_orders.npz")


if __name__ == "__main__":
    main()
