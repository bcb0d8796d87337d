"""Parity of the CUDA path at the LARGE bandwidths against committed samples of the reference's own outputs
(tests/golden/oracle_vectors_large.npz, generator tests/golden/make_golden_large.py): bw = 512 and 1024 Memo, and
bw = 2048 = BASELINE configs[4] (FSTSemiMemo, src/FST_semi_memo.c:68-202) on the orders the reference can compute
(|m| <= 2043: its P_m^m overflows above, src/legendre_polynomials/pmm.c:22-30).

Tolerance 1e-10 relative max-abs (north_star).  Nothing here reads /root/reference or runs the CPU oracle's
transforms; only the port's drand48 restatement regenerates the seeded coefficients.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-10
DLT_ORDERS_2048 = [0, 1, 2, 511, 778, 1023, 1024, 1999, 2042, 2043]
WHOLE_ORDERS_2048 = [0, 1, -1, 700, -777, 1500, 2043, -2043]


@pytest.fixture(scope="module")
def large():
    return np.load(os.path.join(GOLDEN, "oracle_vectors_large.npz"))


@pytest.fixture(scope="module")
def s2():
    import s2kit_b200

    s2kit_b200.lib()
    yield s2kit_b200
    s2kit_b200.release()


def seeded_coeffs(oracle_mod, bw, seed=1000):
    """test_s2_semi_memo.c:156-172 with a fixed seed (the port's drand48 restatement; no transform involved)."""
    L = oracle_mod._load("port")
    rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
    L.orc_gen_coeffs(bw, seed, oracle_mod._p(rc), oracle_mod._p(ic))
    return rc, ic


def cat(pair):
    return np.concatenate([np.asarray(pair[0]).ravel(), np.asarray(pair[1]).ravel()])


@pytest.fixture(scope="module")
def plan2048(s2):
    P = s2.Plan(2048, s2.MEMO, max_batch=1)
    yield P
    P.close()


@pytest.mark.parametrize("bw", [512, 1024])
def test_memo_large_bw_vs_reference_samples(s2, oracle_mod, large, bw):
    gs, cs = (int(v) for v in large[f"bw{bw}_strides"])
    rc, ic = seeded_coeffs(oracle_mod, bw)
    P = s2.Plan(bw, s2.MEMO, max_batch=1)
    rd, idt = P.inverse(rc, ic, 0)
    assert relerr(cat((rd.ravel()[::gs], idt.ravel()[::gs])),
                  cat((large[f"bw{bw}_inv_sample_r"], large[f"bw{bw}_inv_sample_i"]))) < TOL
    fr, fi = P.forward(rd, idt, 0)
    assert relerr(cat((fr[::cs], fi[::cs])), cat((large[f"bw{bw}_fwd_sample_r"], large[f"bw{bw}_fwd_sample_i"]))) < TOL
    # REAL format agrees with COMPLEX on a real field (the reference's two code paths, FST_semi_memo.c:131-145)
    fr2, fi2 = P.forward(rd, np.zeros_like(rd), 1)
    fr3, fi3 = P.forward(rd, np.zeros_like(rd), 0)
    assert relerr(cat((fr2, fi2)), cat((fr3, fi3))) < 1e-11
    P.close()


def test_c5_bw2048_forward_vs_reference(s2, large, plan2048):
    """BASELINE configs[4]: FSTSemiMemo at bw = 2048 on the RandomState(2048) grid, every finite reference entry of a
    strided sample and eight whole orders; ours must also be finite where the reference overflows."""
    bw = 2048
    n = 2 * bw
    rng = np.random.RandomState(2048)
    rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    fr, fi = plan2048.forward(rd, idt, 0)
    assert np.isfinite(fr).all() and np.isfinite(fi).all()
    want_r, want_i = large["bw2048_fwd_sample_r"], large["bw2048_fwd_sample_i"]
    ok = np.isfinite(want_r) & np.isfinite(want_i)
    assert ok.sum() > 0.99 * ok.size
    scale = max(np.abs(want_r[ok]).max(), np.abs(want_i[ok]).max())
    err = max(np.abs(fr[::509][ok] - want_r[ok]).max(), np.abs(fi[::509][ok] - want_i[ok]).max()) / scale
    assert err < TOL, err
    for m in WHOLE_ORDERS_2048:
        a0 = s2.index_of_harmonic_coeff(m, abs(m), bw)
        tag = f"m{m}" if m >= 0 else f"mneg{-m}"
        wr, wi = large[f"bw2048_fwd_order_{tag}_r"], large[f"bw2048_fwd_order_{tag}_i"]
        e = max(np.abs(fr[a0:a0 + bw - abs(m)] - wr).max(), np.abs(fi[a0:a0 + bw - abs(m)] - wi).max()) / scale
        assert e < TOL, (m, e)
    # one copy of the tiled table is resident (the reference needs two packed tables: 2 x 11.46 GB)
    assert plan2048.table_bytes() == plan2048.table_stream_bytes() < 12.0e9
    nan_orders = set(int(m) for m in large["bw2048_ref_nan_orders"])
    assert nan_orders == {m for m in range(-(bw - 1), bw) if abs(m) >= 2044}


def test_c5_bw2048_dlt_vs_reference(large, plan2048):
    """DLTSemi / InvDLTSemi (seminaive.c:153-198 / 56-115) of single orders at bw = 2048."""
    bw = 2048
    n = 2 * bw
    rng = np.random.RandomState(7)
    for m in DLT_ORDERS_2048:
        col = rng.uniform(-1, 1, n)
        assert relerr(plan2048.dlt_semi(col, m)[0], large[f"bw2048_dlt_m{m}"]) < TOL, m
        co = rng.uniform(-1, 1, bw - m)
        assert relerr(plan2048.inv_dlt_semi(co, m)[0], large[f"bw2048_invdlt_m{m}"]) < TOL, m


def test_bw2048_orders_beyond_the_reference_vs_mpmath(plan2048):
    """Orders m >= 2044 at bw = 2048: the reference's Pmm_L2 overflows (pmm.c:21-33) and its transforms return NaN, so the
    committed reference samples cannot pin them.  tests/golden/mp_high_orders.npz holds the exact transforms of seeded
    columns for m = 2040..2047, evaluated with mpmath (60 digits) from the reference's own definitions and cross-checked
    against the reference where it is finite (make_golden_mp.py: every order at bw = 24 to 3e-15, m = 2042 / 2043 at
    bw = 2048 to 2e-14 / 1.4e-13)."""
    mp = np.load(os.path.join(GOLDEN, "mp_high_orders.npz"))
    bw = int(mp["bw"])
    for m in (int(v) for v in mp["orders"]):
        got = plan2048.dlt_semi(mp[f"m{m}_data"], m)[0]
        assert np.isfinite(got).all(), m
        assert relerr(got, mp[f"m{m}_dlt"]) < TOL, (m, relerr(got, mp[f"m{m}_dlt"]))
        back = plan2048.inv_dlt_semi(mp[f"m{m}_coeffs"], m)[0]
        assert np.isfinite(back).all(), m
        assert relerr(back, mp[f"m{m}_inv"]) < TOL, (m, relerr(back, mp[f"m{m}_inv"]))
        assert len(got) == bw - m


@pytest.mark.parametrize("bw", [1024, 2048])
def test_real_format_agrees_with_complex_format_at_large_bw(s2, oracle_mod, plan2048, bw):
    """REAL data format at the large bandwidths (FST_semi_memo.c:110-145, 300-341: only the orders m >= 0 are transformed,
    the others follow from the symmetry of a real field).  The committed reference samples are COMPLEX; here the REAL path
    -- at n = 4096 the staged longitude transforms with half of the spectral rows, the conjugate mirror on the way back --
    must agree with the COMPLEX path on a real field, in both directions."""
    n = 2 * bw
    P = plan2048 if bw == 2048 else s2.Plan(bw, s2.MEMO, max_batch=1)
    rng = np.random.RandomState(bw + 1)
    rd, zero = rng.uniform(-1, 1, (n, n)), np.zeros((n, n))
    cr, ci = P.forward(rd, zero, 0)
    gr, gi = P.forward(rd, zero, 1)
    scale = max(np.abs(cr).max(), np.abs(ci).max())
    assert np.isfinite(gr).all() and np.isfinite(gi).all()
    assert max(np.abs(gr - cr).max(), np.abs(gi - ci).max()) / scale < TOL
    # inverse of the coefficients of a real field (test_s2_semi_memo.c:156-172 symmetry)
    rc, ic = seeded_coeffs(oracle_mod, bw, seed=77)
    xr, xi = P.inverse(rc, ic, 0)
    yr, _ = P.inverse(rc, ic, 1)
    gscale = np.abs(xr).max()
    assert np.abs(xi).max() / gscale < 1e-9  # a real field
    assert np.abs(yr - xr).max() / gscale < TOL
    if bw != 2048:
        P.close()


def test_fly_equals_memo_at_bw2048(s2, oracle_mod, plan2048):
    """FSTSemiFly / InvFSTSemiFly at bw = 2048: the tables are generated per call in six order groups (2 GiB scratch,
    FST_semi_fly.c:96,259-261) by the half-grid generator with a tile offset per group, and both directions read A-order
    tiles.  Same tables as the resident (Memo) plan, so the results must agree to rounding."""
    bw = 2048
    n = 2 * bw
    F = s2.Plan(bw, s2.FLY, max_batch=1)
    rng = np.random.RandomState(4096)
    rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    mr, mi = plan2048.forward(rd, idt, 0)
    fr, fi = F.forward(rd, idt, 0)
    scale = max(np.abs(mr).max(), np.abs(mi).max())
    assert max(np.abs(fr - mr).max(), np.abs(fi - mi).max()) / scale < 1e-12
    rc, ic = seeded_coeffs(oracle_mod, bw, seed=5)
    xr, xi = plan2048.inverse(rc, ic, 0)
    yr, yi = F.inverse(rc, ic, 0)
    assert max(np.abs(yr - xr).max(), np.abs(yi - xi).max()) / np.abs(xr).max() < 1e-12
    F.close()


def test_c5_bw2048_inverse_vs_composed_reference(s2, oracle_mod, large, plan2048):
    """InvFSTSemiMemo at bw = 2048.  The reference's own 2-D inverse is NaN everywhere at this size; the golden is
    composed from its per-order InvDLTSemi for |m| <= 2043 (make_golden_large.py), input = seed-1000 coefficients
    with the orders |m| >= 2044 zeroed."""
    bw = 2048
    rc, ic = seeded_coeffs(oracle_mod, bw)
    for m in range(2044, bw):
        for sm in (m, -m):
            a0 = s2.index_of_harmonic_coeff(sm, m, bw)
            rc[a0:a0 + bw - m] = 0.0
            ic[a0:a0 + bw - m] = 0.0
    rd, idt = plan2048.inverse(rc, ic, 0)
    assert relerr(cat((rd.ravel()[::4099], idt.ravel()[::4099])),
                  cat((large["bw2048_inv_sample_r"], large["bw2048_inv_sample_i"]))) < TOL
    assert relerr(rd[[0, 1, 1000, 2047, 2048, 4095]], large["bw2048_inv_rows_r"]) < TOL
    # symmetric coefficients <=> real field (size-independent property)
    assert np.abs(idt).max() / np.abs(rd).max() < 1e-11


def test_c5_bw2048_sharded_eight_ranks_emulated(s2, large):
    """configs[4] as written: m-sharded tables over 8 ranks, ring -> order exchange (emulated on one device so the
    driver's single-GPU box runs it); coefficients must match the reference sample and every slot must be written."""
    import torch

    bw, nranks = 2048, 8
    n = 2 * bw
    rng = np.random.RandomState(2048)
    rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    plans = [s2.ShardedPlan(bw, r, nranks) for r in range(nranks)]
    nr, blk = plans[0].rings, plans[0].block_doubles
    gr, gi = torch.tensor(rd, device="cuda"), torch.tensor(idt, device="cuda")
    send = []
    for r, P in enumerate(plans):
        s = torch.zeros(nranks * blk, device="cuda", dtype=torch.float64)
        P.fst_rings(gr[r * nr:(r + 1) * nr].contiguous(), gi[r * nr:(r + 1) * nr].contiguous(), s)
        P.synchronize()
        send.append(s.view(nranks, -1))
    out_r = torch.full((bw * bw,), float("nan"), device="cuda", dtype=torch.float64)
    out_i = torch.full_like(out_r, float("nan"))
    for r, P in enumerate(plans):
        recv = torch.stack([send[s][r] for s in range(nranks)]).contiguous().view(-1)
        P.fst_orders(recv, out_r, out_i)
        P.synchronize()
    fr, fi = out_r.cpu().numpy(), out_i.cpu().numpy()
    assert np.isfinite(fr).all() and np.isfinite(fi).all()
    want_r, want_i = large["bw2048_fwd_sample_r"], large["bw2048_fwd_sample_i"]
    ok = np.isfinite(want_r) & np.isfinite(want_i)
    scale = max(np.abs(want_r[ok]).max(), np.abs(want_i[ok]).max())
    err = max(np.abs(fr[::509][ok] - want_r[ok]).max(), np.abs(fi[::509][ok] - want_i[ok]).max()) / scale
    assert err < TOL, err
    assert sum(P.table_bytes() for P in plans) < 1.02 * 8 * plans[0].table_bytes()
    for P in plans:
        P.close()
