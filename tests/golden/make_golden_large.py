"""Golden samples of the reference's outputs at the LARGE bandwidths (bw = 512, 1024 Memo; bw = 2048 = BASELINE
configs[4]) -> tests/golden/oracle_vectors_large.npz.  Run in the build container (needs /root/reference, ~25 GB of
RAM and ~15 minutes of CPU: the reference builds 2 x 11.5 GB of tables at bw = 2048).

The reference (oracle/_ref: its unmodified sources + oracle/fftw_stub) is the only source of numbers here:
  bw512_*, bw1024_*   seeded coefficients (seed 1000) -> InvFSTSemiMemo -> FSTSemiMemo, COMPLEX; strided samples
  bw2048_fwd_*        FSTSemiMemo (src/FST_semi_memo.c:68-202) of the RandomState(2048) grid, strided sample of the bw^2
                      coefficients plus whole orders; NaN where the reference returns NaN (|m| >= 2044,
                      src/legendre_polynomials/pmm.c:22-30) -- parity is defined on the finite entries
  bw2048_dlt_*        DLTSemi / InvDLTSemi (src/legendre_transform/seminaive.c:153-198 / 56-115) of seeded columns
                      for a list of orders m <= 2043
  bw2048_inv_*        the 2-D inverse.  The reference's own InvFSTSemiMemo returns NaN everywhere at bw = 2048 (the NaN
                      tables of orders >= 2044 reach every grid point through the longitude FFT), so the golden is
                      COMPOSED from the reference's per-order InvDLTSemi for |m| <= 2043 exactly as
                      InvFSTSemiMemo (src/FST_semi_memo.c:228-351) assembles them, with numpy's FFT along phi; the
                      same composition is checked against the reference's real InvFSTSemiMemo at bw = 64 below.
                      Input: seed-1000 coefficients with the orders |m| >= 2044 zeroed.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from oracle import _p, coef_index  # noqa: E402

DLT_ORDERS_2048 = [0, 1, 2, 511, 778, 1023, 1024, 1999, 2042, 2043]
WHOLE_ORDERS_2048 = [0, 1, -1, 700, -777, 1500, 2043, -2043]


def composed_inverse(R, rc, ic, m_max):
    """InvFSTSemiMemo (COMPLEX) assembled from the reference's InvDLTSemi per order, orders |m| <= m_max."""
    bw = R.bw
    n = 2 * bw
    F = np.zeros((n, n), dtype=complex)  # [order row m'][latitude j]
    a, b = np.zeros(n), np.zeros(n)
    for m in range(0, m_max + 1):
        at = coef_index(m, m, bw)
        R.L.ref_inv_dlt_semi(R.h, _p(np.ascontiguousarray(rc[at:at + bw - m])), m, _p(a))
        R.L.ref_inv_dlt_semi(R.h, _p(np.ascontiguousarray(ic[at:at + bw - m])), m, _p(b))
        F[m] = a + 1j * b
        if m:
            at = coef_index(-m, m, bw)
            sg = -1.0 if m & 1 else 1.0  # FST_semi_memo.c:303-331
            R.L.ref_inv_dlt_semi(R.h, _p(np.ascontiguousarray(rc[at:at + bw - m])), m, _p(a))
            R.L.ref_inv_dlt_semi(R.h, _p(np.ascontiguousarray(ic[at:at + bw - m])), m, _p(b))
            F[n - m] = sg * (a + 1j * b)
    F *= 1.0 / np.sqrt(2.0 * np.pi)  # FST_semi_memo.c:344
    g = np.fft.ifft(F, axis=0) * n   # data[j, k] = sum_m' F[m', j] e^{+2 pi i m' k / n}
    return np.ascontiguousarray(g.T.real), np.ascontiguousarray(g.T.imag)


def main():
    oracle.build()
    out = {}
    t0 = time.time()
    # the composition is the reference's InvFSTSemiMemo: check at a size where that function works
    R = oracle.Oracle(64, "ref")
    rc, ic = R.gen_coeffs(1000)
    want = R.inverse(rc, ic, 0)
    got = composed_inverse(R, rc, ic, 63)
    err = max(np.abs(got[0] - want[0]).max(), np.abs(got[1] - want[1]).max()) / np.abs(want[0]).max()
    assert err < 1e-13, err
    print("composition vs InvFSTSemiMemo at bw 64:", err)
    R.close()

    for bw, gs, cs in ((512, 1031, 127), (1024, 4099, 509)):
        R = oracle.Oracle(bw, "ref")
        rc, ic = R.gen_coeffs(1000)
        rd, idt = R.inverse(rc, ic, 0)
        fr, fi = R.forward(rd, idt, 0)
        out[f"bw{bw}_inv_sample_r"] = rd.ravel()[::gs].copy()
        out[f"bw{bw}_inv_sample_i"] = idt.ravel()[::gs].copy()
        out[f"bw{bw}_fwd_sample_r"] = fr[::cs].copy()
        out[f"bw{bw}_fwd_sample_i"] = fi[::cs].copy()
        out[f"bw{bw}_strides"] = np.array([gs, cs])
        R.close()
        print(f"bw {bw} done, {time.time() - t0:.0f} s")

    bw = 2048
    n = 2 * bw
    R = oracle.Oracle(bw, "ref")
    print(f"bw 2048 reference tables built, {time.time() - t0:.0f} s")
    rng = np.random.RandomState(2048)
    rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    fr, fi = R.forward(rd, idt, 0)
    out["bw2048_fwd_sample_r"] = fr[::509].copy()
    out["bw2048_fwd_sample_i"] = fi[::509].copy()
    for m in WHOLE_ORDERS_2048:
        a0 = coef_index(m, abs(m), bw)
        tag = f"m{m}" if m >= 0 else f"mneg{-m}"
        out[f"bw2048_fwd_order_{tag}_r"] = fr[a0:a0 + bw - abs(m)].copy()
        out[f"bw2048_fwd_order_{tag}_i"] = fi[a0:a0 + bw - abs(m)].copy()
    nan_orders = [m for m in range(-(bw - 1), bw)
                  if not np.isfinite(fr[coef_index(m, abs(m), bw):coef_index(m, abs(m), bw) + bw - abs(m)]).all()]
    out["bw2048_ref_nan_orders"] = np.array(nan_orders)
    print("reference NaN orders:", min(nan_orders), "..", max(nan_orders), len(nan_orders))
    # 1-D transforms of single orders
    rng = np.random.RandomState(7)
    for m in DLT_ORDERS_2048:
        col = rng.uniform(-1, 1, n)
        res = np.zeros(bw)
        R.L.ref_dlt_semi(R.h, _p(col), m, _p(res))
        out[f"bw2048_dlt_m{m}"] = res[:bw - m].copy()
        co = rng.uniform(-1, 1, bw - m)
        grid = np.zeros(n)
        R.L.ref_inv_dlt_semi(R.h, _p(co), m, _p(grid))
        out[f"bw2048_invdlt_m{m}"] = grid.copy()
    # composed 2-D inverse, orders |m| <= 2043
    rc, ic = R.gen_coeffs(1000)
    for m in range(2044, bw):
        for sm in (m, -m):
            a0 = coef_index(sm, m, bw)
            rc[a0:a0 + bw - m] = 0.0
            ic[a0:a0 + bw - m] = 0.0
    gr, gi = composed_inverse(R, rc, ic, 2043)
    out["bw2048_inv_sample_r"] = gr.ravel()[::4099].copy()
    out["bw2048_inv_sample_i"] = gi.ravel()[::4099].copy()
    out["bw2048_inv_rows_r"] = gr[[0, 1, 1000, 2047, 2048, 4095]].copy()
    R.close()
    print(f"bw 2048 done, {time.time() - t0:.0f} s")
    path = os.path.join(HERE, "oracle_vectors_large.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
