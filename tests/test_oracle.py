"""Pins the CPU oracle (both the compiled reference and the independent port) to every stored vector the
reference's tests hold for this path: the four golden convolution outputs (dist/test.sh:39-61, eps 1e-15),
the four Y_l^m known-answer grids (dist/S2kitHowTo.pdf 2.4.2), plus the committed outputs of the
reference itself (tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest

from conftest import relerr

KINDS = ["port", "ref"]


def _oracle(oracle_mod, bw, kind, **kw):
    if kind == "ref" and not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_mod.Oracle(bw, kind, **kw)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("bw", [64, 128])
@pytest.mark.parametrize("variant", ["memo", "fly"])
def test_golden_convolution(oracle_mod, refdata, kind, bw, variant):
    if kind == "port" and variant == "fly":
        pytest.skip("port has one code path")
    n = 2 * bw
    O = _oracle(oracle_mod, bw, kind, variant=variant, tables=False)
    s, f = refdata[f"s{bw}"].reshape(n, n), refdata[f"f{bw}"].reshape(n, n)
    z = np.zeros_like(s)
    rr, ir = O.conv(s, z, f, z)
    gold = refdata[f"o{bw}_conv_semi_{variant}_original"].reshape(n, n)
    assert np.abs(rr - gold).max() <= 1e-15  # the reference's own eps


def _kat_coeffs(O, grid_interleaved, bw):
    n = 2 * bw
    g = grid_interleaved.reshape(n * n, 2)
    return O.forward(g[:, 0].reshape(n, n).copy(), g[:, 1].reshape(n, n).copy(), 0)


@pytest.mark.parametrize("kind", KINDS)
def test_known_answer_ylm(oracle_mod, refdata, kind):
    cases = [
        ("y20_bw8", 8, {(0, 2): 1.0}),
        ("y31_bw8", 8, {(1, 3): 1.0}),
        ("y43_bw23", 23, {(3, 4): complex(math.sqrt(2.0), math.pi)}),
        ("yMix_bw17", 17, {(1, 1): 1.0, (-2, 5): complex(3.0, -2.0)}),
    ]
    for name, bw, expect in cases:
        O = _oracle(oracle_mod, bw, kind)
        rc, ic = _kat_coeffs(O, refdata[name], bw)
        want = np.zeros(bw * bw, dtype=complex)
        for (m, l), v in expect.items():
            want[oracle_mod.coef_index(m, l, bw)] = v
        assert np.abs((rc + 1j * ic) - want).max() < 2e-14, name


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("bw", [16, 64])
def test_committed_reference_outputs(oracle_mod, vectors, kind, bw):
    O = _oracle(oracle_mod, bw, kind)
    rc, ic = O.gen_coeffs(1000)
    assert np.array_equal(rc, vectors[f"coef_seed1000_bw{bw}_r"])
    assert np.array_equal(ic, vectors[f"coef_seed1000_bw{bw}_i"])
    assert relerr(O.weights(), vectors[f"weights_bw{bw}"]) < 1e-15
    for m in ((0, 1, 2, 7, 14, 15) if bw == 16 else (0, 1, 2, 31, 62, 63)):
        assert relerr(O.table(m), vectors[f"table_bw{bw}_m{m}"]) < 1e-14
    for fmt, tag in ((0, "complex"), (1, "real")):
        rd, idt = O.inverse(rc, ic, fmt)
        g = np.concatenate([vectors[f"inv_{tag}_bw{bw}_r"], vectors[f"inv_{tag}_bw{bw}_i"]])
        assert relerr(np.concatenate([rd, idt]), g) < 1e-13
        fr, fi = O.forward(vectors[f"inv_{tag}_bw{bw}_r"], vectors[f"inv_{tag}_bw{bw}_i"], fmt)
        c = np.concatenate([vectors[f"fwd_{tag}_bw{bw}_r"], vectors[f"fwd_{tag}_bw{bw}_i"]])
        assert relerr(np.concatenate([fr, fi]), c) < 1e-13
        # and the round trip returns the seeded coefficients
        assert relerr(np.concatenate([fr, fi]), np.concatenate([rc, ic])) < 1e-11


@pytest.mark.parametrize("kind", KINDS)
def test_config_c1_s64_forward(oracle_mod, refdata, vectors, kind):
    bw = 64
    O = _oracle(oracle_mod, bw, kind)
    s = refdata["s64"].reshape(128, 128)
    z = np.zeros_like(s)
    for fmt, tag in ((0, "complex"), (1, "real")):
        fr, fi = O.forward(s, z, fmt)
        c = np.concatenate([vectors[f"s64_fwd_{tag}_r"], vectors[f"s64_fwd_{tag}_i"]])
        assert relerr(np.concatenate([fr, fi]), c) < 1e-13
    zr, _ = O.zonal(refdata["f64"].reshape(128, 128), z, 1)
    assert relerr(zr, vectors["f64_zonal_r"]) < 1e-13


@pytest.mark.parametrize("kind", KINDS)
def test_full_complex_coefficients(oracle_mod, vectors, kind):
    O = _oracle(oracle_mod, 64, kind)
    rd, idt = O.inverse(vectors["coef_full_bw64_r"], vectors["coef_full_bw64_i"], 0)
    g = np.concatenate([vectors["inv_full_bw64_r"], vectors["inv_full_bw64_i"]])
    assert relerr(np.concatenate([rd, idt]), g) < 1e-13
    fr, fi = O.forward(rd, idt, 0)
    assert relerr(np.concatenate([fr, fi]), np.concatenate([vectors["fwd_full_bw64_r"], vectors["fwd_full_bw64_i"]])) < 1e-13


def test_port_matches_reference_bw256(oracle_mod, vectors):
    """Config C3 sample: port vs the committed strided sample of the reference's bw=256 outputs."""
    bw = 256
    O = oracle_mod.Oracle(bw, "port")
    rc, ic = O.gen_coeffs(1000)
    rd, idt = O.inverse(rc, ic, 0)
    assert relerr(rd.ravel()[::257], vectors["bw256_inv_sample_r"]) < 1e-12
    assert relerr(idt.ravel()[::257], vectors["bw256_inv_sample_i"] + 0.0) < 1e-9 or np.abs(idt).max() < 1e-9
    fr, fi = O.forward(rd, idt, 0)
    assert relerr(fr[::61], vectors["bw256_fwd_sample_r"]) < 1e-12
    assert relerr(fi[::61], vectors["bw256_fwd_sample_i"]) < 1e-12
    assert relerr(O.table(200), vectors["bw256_table_m200"]) < 1e-14
    assert relerr(O.table(1)[:4096], vectors["bw256_table_m1_head"]) < 1e-14


def test_layout_helpers(oracle_mod):
    for bw in (8, 16, 17, 64):
        seen = set()
        for m in range(-(bw - 1), bw):
            for l in range(abs(m), bw):
                seen.add(oracle_mod.coef_index(m, l, bw))
        assert seen == set(range(bw * bw))
    # totals quoted in SURVEY.md section 3.5
    for bw, total in ((64, 44736), (128, 353664), (256, 2812672)):
        assert sum(oracle_mod.table_size(m, bw) for m in range(bw)) == total


def test_port_matches_reference_samples_bw512(oracle_mod):
    """The port against the committed samples of the REFERENCE's bw = 512 outputs (make_golden_large.py)."""
    import os

    from conftest import GOLDEN, relerr

    large = np.load(os.path.join(GOLDEN, "oracle_vectors_large.npz"))
    bw = 512
    gs, cs = (int(v) for v in large[f"bw{bw}_strides"])
    O = oracle_mod.Oracle(bw, "port")
    rc, ic = O.gen_coeffs(1000)
    rd, idt = O.inverse(rc, ic, 0)
    assert relerr(rd.ravel()[::gs], large[f"bw{bw}_inv_sample_r"]) < 1e-12
    fr, fi = O.forward(rd, idt, 0)
    assert relerr(fr[::cs], large[f"bw{bw}_fwd_sample_r"]) < 1e-12
    assert relerr(fi[::cs], large[f"bw{bw}_fwd_sample_i"]) < 1e-12
    O.close()
    # the bw = 2048 fixtures: the reference returns NaN exactly for the eight orders |m| >= 2044 (pmm.c:22-30)
    assert sorted(int(m) for m in large["bw2048_ref_nan_orders"]) == [-2047, -2046, -2045, -2044, 2044, 2045, 2046, 2047]
    assert np.isfinite(large["bw2048_inv_sample_r"]).all()


def test_mpmath_fixture_agrees_with_the_reference_where_it_is_finite():
    """tests/golden/mp_high_orders.npz (make_golden_mp.py) pins the orders the reference cannot compute at bw = 2048.  Its
    cross-check entries -- the mpmath DLT / inverse DLT of the SAME seeded columns make_golden_large.py fed to the
    reference's DLTSemi / InvDLTSemi at m = 2042, 2043 -- must agree with the committed reference outputs."""
    import os

    from conftest import GOLDEN

    mp = np.load(os.path.join(GOLDEN, "mp_high_orders.npz"))
    large = np.load(os.path.join(GOLDEN, "oracle_vectors_large.npz"))
    for m in (2042, 2043):
        a, b = mp[f"crosscheck_m{m}_dlt"], large[f"bw2048_dlt_m{m}"]
        assert np.abs(a - b).max() / np.abs(a).max() < 1e-11
        a, b = mp[f"crosscheck_m{m}_inv"], large[f"bw2048_invdlt_m{m}"]
        assert np.abs(a - b).max() / np.abs(a).max() < 1e-11
    # the quadrature weights of the fixture are the reference's (weights.c:32-47) to the last bits
    w = mp["weights"]
    assert abs(w.sum() - 2.0) < 1e-13  # the weights integrate 1 over [-1, 1]
