"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the committed golden vectors.

Tolerance: BASELINE.json's north_star asks for a relative max-abs error <= 1e-10 on the coefficients and on the
round-trip grid, measured as max|ours - ref| / max|ref|.  Tables and small cases are held
to much tighter bounds.  All tests need a CUDA device (-m gpu); nothing here reads /root/reference.
"""
import math

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def s2():
    import s2kit_b200

    s2kit_b200.lib()
    yield s2kit_b200
    s2kit_b200.release()


@pytest.fixture(scope="module")
def oracles(oracle_mod):
    cache = {}

    def get(bw):
        if bw not in cache:
            cache[bw] = oracle_mod.Oracle(bw, oracle_mod.best_kind())
        return cache[bw]

    yield get
    for o in cache.values():
        o.close()


def cat(pair):
    return np.concatenate([np.asarray(pair[0]).ravel(), np.asarray(pair[1]).ravel()])


# ------------------------------------------------------------------------------------------------ tables (K7)
@pytest.mark.parametrize("bw", [16, 64, 256, 24])
def test_tables_match_oracle(s2, oracles, bw):
    O, P = oracles(bw), s2.Plan(bw)
    for m in sorted({0, 1, 2, 3, bw // 2, bw // 2 + 1, bw - 2, bw - 1}):
        assert relerr(P.table(m), O.table(m)) < 1e-13, (bw, m)
    P.close()


def test_table_golden_and_reference_symbols(s2, vectors):
    # GenerateCosPmlTable / TransposeCosPmlTable through the drop-in symbols vs committed reference output
    for bw, ms in ((16, (0, 1, 2, 7, 14, 15)), (64, (0, 1, 2, 31, 62, 63))):
        for m in ms:
            assert relerr(s2.GenerateCosPmlTable(bw, m), vectors[f"table_bw{bw}_m{m}"]) < 1e-13
    assert relerr(s2.GenerateCosPmlTable(256, 200), vectors["bw256_table_m200"]) < 1e-13
    assert relerr(s2.GenerateCosPmlTable(256, 1)[:4096], vectors["bw256_table_m1_head"]) < 1e-13
    assert np.array_equal(s2.GenerateWeightsForDLT(64), vectors["weights_bw64"])


# ------------------------------------------------------------------------------------------------ transforms
@pytest.mark.parametrize("bw", [16, 64, 128, 256, 24])
@pytest.mark.parametrize("fmt", [0, 1])
def test_forward_inverse_match_oracle(s2, oracles, bw, fmt):
    O, P = oracles(bw), s2.Plan(bw)
    rc, ic = O.gen_coeffs(1000)
    want_g = O.inverse(rc, ic, fmt)
    assert relerr(cat(P.inverse(rc, ic, fmt)), cat(want_g)) < TOL
    want_c = O.forward(want_g[0], want_g[1], fmt)
    got_c = P.forward(want_g[0], want_g[1], fmt)
    assert relerr(cat(got_c), cat(want_c)) < TOL
    assert relerr(cat(got_c), cat((rc, ic))) < 1e-9  # round trip back to the seeded coefficients
    P.close()


def test_committed_vectors_without_oracle(s2, refdata, vectors):
    """Golden fixtures only (tests/golden): config C1 forward of data/s64.dat and the seeded round trips."""
    s = refdata["s64"].reshape(128, 128)
    z = np.zeros_like(s)
    for fmt, tag in ((0, "complex"), (1, "real")):
        got = s2.FSTSemiMemo(s, z, 64, fmt)
        assert relerr(cat(got), cat((vectors[f"s64_fwd_{tag}_r"], vectors[f"s64_fwd_{tag}_i"]))) < TOL
        got = s2.FSTSemiFly(s, z, 64, fmt)
        assert relerr(cat(got), cat((vectors[f"s64_fwd_{tag}_r"], vectors[f"s64_fwd_{tag}_i"]))) < TOL
    for bw in (16, 64):
        rc, ic = vectors[f"coef_seed1000_bw{bw}_r"], vectors[f"coef_seed1000_bw{bw}_i"]
        for fmt, tag in ((0, "complex"), (1, "real")):
            g = s2.InvFSTSemiMemo(rc, ic, bw, fmt)
            want = (vectors[f"inv_{tag}_bw{bw}_r"], vectors[f"inv_{tag}_bw{bw}_i"])
            assert relerr(cat(g), cat(want)) < TOL
            c = s2.FSTSemiMemo(want[0], want[1], bw, fmt)
            assert relerr(cat(c), cat((vectors[f"fwd_{tag}_bw{bw}_r"], vectors[f"fwd_{tag}_bw{bw}_i"]))) < TOL
            g = s2.InvFSTSemiFly(rc, ic, bw, fmt)
            assert relerr(cat(g), cat(want)) < TOL
    # fully complex coefficients (independent negative orders)
    g = s2.InvFSTSemiMemo(vectors["coef_full_bw64_r"], vectors["coef_full_bw64_i"], 64, 0)
    assert relerr(cat(g), cat((vectors["inv_full_bw64_r"], vectors["inv_full_bw64_i"]))) < TOL
    c = s2.FSTSemiMemo(vectors["inv_full_bw64_r"], vectors["inv_full_bw64_i"], 64, 0)
    assert relerr(cat(c), cat((vectors["fwd_full_bw64_r"], vectors["fwd_full_bw64_i"]))) < TOL
    # strided sample of the reference's bw = 256 outputs (config C3, function 0)
    zr = s2.FZTSemiMemo(refdata["f64"].reshape(128, 128), z, 64, 1)[0]
    assert relerr(zr, vectors["f64_zonal_r"]) < TOL


def test_config_c3_sample_bw256(s2, oracle_mod, vectors):
    bw = 256
    # regenerate the seeded coefficients with the port's drand48 restatement (no oracle transform involved)
    O = oracle_mod.Oracle(16, "port")
    rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
    O.L.orc_gen_coeffs(bw, 1000, oracle_mod._p(rc), oracle_mod._p(ic))
    P = s2.Plan(bw)
    rd, idt = P.inverse(rc, ic, 0)
    assert relerr(rd.ravel()[::257], vectors["bw256_inv_sample_r"]) < TOL
    fr, fi = P.forward(rd, idt, 0)
    assert relerr(fr[::61], vectors["bw256_fwd_sample_r"]) < TOL
    assert relerr(fi[::61], vectors["bw256_fwd_sample_i"]) < TOL
    P.close()


@pytest.mark.parametrize("variant", ["Memo", "Fly"])
@pytest.mark.parametrize("bw", [64, 128])
def test_golden_convolution(s2, refdata, variant, bw):
    """Config C2 and dist/test.sh:39-61: ConvOn2SphereSemi{Memo,Fly} on the reference's data files."""
    n = 2 * bw
    s, f = refdata[f"s{bw}"].reshape(n, n), refdata[f"f{bw}"].reshape(n, n)
    z = np.zeros_like(s)
    rr, ir = getattr(s2, f"ConvOn2SphereSemi{variant}")(s, z, f, z, bw)
    gold = refdata[f"o{bw}_conv_semi_{variant.lower()}_original"].reshape(n, n)
    # the reference's own regression eps is 1e-15 absolute on values of O(0.1)
    assert np.abs(rr - gold).max() <= 1e-15
    assert np.abs(ir).max() < 1e-15


def test_known_answer_ylm(s2, oracle_mod, refdata):
    """dist/S2kitHowTo.pdf 2.4.2: sampled Y_l^m grids must give exactly those coefficients (odd bw included)."""
    cases = [("y20_bw8", 8, {(0, 2): 1.0}), ("y31_bw8", 8, {(1, 3): 1.0}),
             ("y43_bw23", 23, {(3, 4): complex(math.sqrt(2.0), math.pi)}),
             ("yMix_bw17", 17, {(1, 1): 1.0, (-2, 5): complex(3.0, -2.0)})]
    for name, bw, expect in cases:
        n = 2 * bw
        g = refdata[name].reshape(n * n, 2)
        rc, ic = s2.FSTSemiMemo(g[:, 0].reshape(n, n), g[:, 1].reshape(n, n), bw, 0)
        want = np.zeros(bw * bw, dtype=complex)
        for (m, l), v in expect.items():
            want[oracle_mod.coef_index(m, l, bw)] = v
        assert np.abs((rc + 1j * ic) - want).max() < 2e-14, name


def test_fly_equals_memo(s2, oracles):
    bw = 64
    O = oracles(bw)
    rc, ic = O.gen_coeffs(1001)
    Pm, Pf = s2.Plan(bw, s2.MEMO), s2.Plan(bw, s2.FLY)
    gm, gf = Pm.inverse(rc, ic, 0), Pf.inverse(rc, ic, 0)
    assert relerr(cat(gf), cat(gm)) < 1e-14
    assert relerr(cat(Pf.forward(gm[0], gm[1], 0)), cat(Pm.forward(gm[0], gm[1], 0))) < 1e-14
    Pm.close()
    Pf.close()


def test_zonal_transmult_dlt(s2, oracles):
    bw = 64
    n = 2 * bw
    O, P = oracles(bw), s2.Plan(bw)
    rng = np.random.RandomState(3)
    a, b = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    for fmt in (0, 1):
        want = O.zonal(a, b, fmt)
        rr, ir = np.zeros(bw), np.zeros(bw)
        P.fzt(a, b, rr, ir, fmt)
        assert relerr(cat((rr, ir)), cat(want)) < TOL
    # TransMult with a non-zero imaginary filter exercises ComplexMult's sign (util.c:26)
    rd, idt = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)
    rf, ifl = rng.uniform(-1, 1, bw), rng.uniform(-1, 1, bw)
    want = O.spectral_multiply(rd, idt, rf, ifl)
    rr, ir = np.zeros(bw * bw), np.zeros(bw * bw)
    P.trans_mult(rd, idt, rf, ifl, rr, ir)
    assert relerr(cat((rr, ir)), cat(want)) < 1e-14
    if O.kind == "ref":
        from oracle import _p
        for m in (0, 1, 2, 33, 63):
            col = rng.uniform(-1, 1, n)
            want = np.zeros(bw)
            O.L.ref_dlt_semi(O.h, _p(col), m, _p(want))
            assert relerr(P.dlt_semi(col, m)[0], want[: bw - m]) < TOL
            co = rng.uniform(-1, 1, bw - m)
            want = np.zeros(n)
            O.L.ref_inv_dlt_semi(O.h, _p(co), m, _p(want))
            assert relerr(P.inv_dlt_semi(co, m)[0], want) < TOL
    P.close()


def test_naive_dlt_matches_reference_and_seminaive(s2, oracles):
    """DLTNaive / InvDLTNaive (naive.c) on the GPU with the theta-space table of GeneratePmlTable: against the
    reference's own functions where the compiled reference is available, and against the seminaive transform of the
    same column (the two algorithms compute the same coefficients, test/test_DLT_naive.c vs test_DLT_semi.c)."""
    bw = 32
    n = 2 * bw
    O, P = oracles(bw), s2.Plan(bw)
    rng = np.random.RandomState(11)
    w = s2.GenerateWeightsForDLT(bw)
    theta = (2.0 * np.arange(n) + 1.0) * np.pi / (2.0 * n)
    for m in (0, 1, 6, 31):
        tab = s2.GeneratePmlTable(bw, m)
        assert relerr(tab[:n], s2.Pmm_L2(m, theta)) < 1e-14  # first row is P_m^m itself (pml.c:52-56)
        co = rng.uniform(-1, 1, bw - m)
        grid = s2.InvDLTNaive(co, bw, m, tab)
        assert relerr(grid, tab.reshape(bw - m, n).T @ co) < 1e-14
        back = s2.DLTNaive(grid, bw, m, w, tab)
        assert relerr(back, co) < 1e-11  # exact quadrature for band-limited columns
        assert relerr(P.dlt_semi(grid, m)[0], back) < TOL
        if O.kind == "ref":
            import ctypes

            from oracle import _p
            rtab, want = np.zeros_like(tab), np.zeros(bw - m)
            ws = np.zeros(16 * bw)
            O.L.GeneratePmlTable(ctypes.c_int(bw), ctypes.c_int(m), _p(rtab), _p(ws))
            assert relerr(tab, rtab) < 1e-14
            O.L.DLTNaive(_p(grid), ctypes.c_int(bw), ctypes.c_int(m), _p(w), _p(want), _p(rtab), _p(ws))
            assert relerr(back, want) < 1e-13
            wantg = np.zeros(n)
            O.L.InvDLTNaive(_p(co), ctypes.c_int(bw), ctypes.c_int(m), _p(wantg), _p(rtab))
            assert relerr(grid, wantg) < 1e-14
    P.close()


# ------------------------------------------------------------------------------------------------ batched device path
def test_batched_device_pointers(s2, oracles):
    """Device-resident batch through s2kit_cuda_inv_fst / s2kit_cuda_fst: ragged chunking (batch 5, chunk 2)."""
    import torch

    bw, batch = 64, 5
    n = 2 * bw
    O = oracles(bw)
    P = s2.Plan(bw, s2.MEMO, max_batch=2)
    coefs = [O.gen_coeffs(1000 + k) for k in range(batch)]
    rc = torch.tensor(np.stack([c[0] for c in coefs]), device="cuda")
    ic = torch.tensor(np.stack([c[1] for c in coefs]), device="cuda")
    for fmt in (0, 1):
        rd = torch.zeros(batch, n, n, device="cuda", dtype=torch.float64)
        idt = torch.zeros_like(rd)
        P.inv_fst(rc, ic, rd, idt, fmt)
        rc2, ic2 = torch.zeros_like(rc), torch.zeros_like(ic)
        P.fst(rd, idt, rc2, ic2, fmt)
        P.synchronize()
        for k in range(batch):
            want_g = O.inverse(coefs[k][0], coefs[k][1], fmt)
            assert relerr(cat((rd[k].cpu().numpy(), idt[k].cpu().numpy())), cat(want_g)) < TOL
            want_c = O.forward(want_g[0], want_g[1], fmt)
            assert relerr(cat((rc2[k].cpu().numpy(), ic2[k].cpu().numpy())), cat(want_c)) < TOL
    # empty batch is a no-op
    e = torch.zeros(0, device="cuda", dtype=torch.float64)
    P.fst(e, e, e, e, 0)
    P.close()


@pytest.mark.parametrize("bw,batch", [(128, 19), (256, 11)])
def test_batched_wide_panels_match_oracle(s2, oracles, bw, batch):
    """Batches wide enough for the persistent warp-specialised kernels (kernels_pipe.cu: 32-column panels, ragged last
    panel, chunking), every function checked against the oracle in both data formats; fully complex coefficients so
    that the negative orders carry independent data."""
    import torch

    n = 2 * bw
    O = oracles(bw)
    P = s2.Plan(bw, s2.MEMO, max_batch=12)
    rng = np.random.RandomState(bw + batch)
    coefs = [O.gen_coeffs(2000 + k) for k in range(batch)]
    for k in range(1, batch, 2):  # independent +-m data on every other function
        coefs[k] = (rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw))
    rc = torch.tensor(np.stack([c[0] for c in coefs]), device="cuda")
    ic = torch.tensor(np.stack([c[1] for c in coefs]), device="cuda")
    for fmt in (0, 1):
        rd = torch.zeros(batch, n, n, device="cuda", dtype=torch.float64)
        idt = torch.zeros_like(rd)
        P.inv_fst(rc, ic, rd, idt, fmt)
        rc2, ic2 = torch.full_like(rc, float("nan")), torch.full_like(ic, float("nan"))
        P.fst(rd, idt, rc2, ic2, fmt)
        P.synchronize()
        for k in range(batch):
            want_g = O.inverse(coefs[k][0], coefs[k][1], fmt)
            assert relerr(cat((rd[k].cpu().numpy(), idt[k].cpu().numpy())), cat(want_g)) < TOL
            want_c = O.forward(want_g[0], want_g[1], fmt)
            assert relerr(cat((rc2[k].cpu().numpy(), ic2[k].cpu().numpy())), cat(want_c)) < TOL
    P.close()


@pytest.mark.parametrize("mode", ["0", "1"])
def test_alternative_forward_contractions(mode):
    """The batched forward contraction has three implementations selected once per process by S2KIT_CUDA_PIPE: the fused
    persistent kernel (default, covered by every other test), K2 + plain K3 ("0") and K2 + the streamed persistent K3
    ("1").  Re-run the wide-panel parity test in a child process for the two alternatives."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, S2KIT_CUDA_PIPE=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_batched_wide_panels_match_oracle"], env=env, cwd=root, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout, r.stdout[-500:]


def test_batched_convolution_device(s2, oracles):
    import torch

    bw, batch = 32, 3
    n = 2 * bw
    O = oracles(bw)
    rng = np.random.RandomState(5)
    sig = rng.uniform(-1, 1, (batch, n, n))
    fil = rng.uniform(-1, 1, (n, n))
    z = np.zeros((n, n))
    P = s2.Plan(bw, s2.MEMO, max_batch=2)
    d = lambda a: torch.tensor(a, device="cuda")  # noqa: E731
    rr, ir = torch.zeros(batch, n, n, device="cuda", dtype=torch.float64), torch.zeros(batch, n, n, device="cuda", dtype=torch.float64)
    P.conv(d(sig), d(np.zeros_like(sig)), d(fil), d(z), rr, ir, shared_filter=True)
    P.synchronize()
    for k in range(batch):
        want = O.conv(sig[k], z, fil, z)
        assert relerr(rr[k].cpu().numpy(), want[0]) < TOL
    P.close()


def test_full_size_properties_bw256_batch1024(s2):
    """BASELINE configs[2] at full size: 1024 functions at bw = 256, device resident.  Size-independent properties:
    round trip coefficients -> grid -> coefficients, real-valuedness of the field, and linearity."""
    import torch

    from bench import synth_coeffs

    bw, batch = 256, 1024
    n = 2 * bw
    dev = torch.device("cuda", 0)
    P = s2.Plan(bw, s2.MEMO, max_batch=64)
    rc, ic = synth_coeffs(torch, bw, batch, dev, 1000)
    rd = torch.empty(batch, n, n, device=dev, dtype=torch.float64)
    idt = torch.empty_like(rd)
    P.inv_fst(rc, ic, rd, idt, 0)
    rc2, ic2 = torch.empty_like(rc), torch.empty_like(ic)
    P.fst(rd, idt, rc2, ic2, 0)
    P.synchronize()
    scale = float(torch.maximum(rc.abs().max(), ic.abs().max()))
    err = float(torch.maximum((rc2 - rc).abs().max(), (ic2 - ic).abs().max())) / scale
    assert err < 1e-10, err
    assert float(idt.abs().max()) / float(rd.abs().max()) < 1e-12  # symmetric coefficients <=> real field
    # linearity on a slice: T(a f0 + b f1) = a T(f0) + b T(f1)
    mix_r, mix_i = 0.75 * rd[0] - 1.5 * rd[1], 0.75 * idt[0] - 1.5 * idt[1]
    o_r, o_i = torch.empty(1, bw * bw, device=dev, dtype=torch.float64), torch.empty(1, bw * bw, device=dev, dtype=torch.float64)
    P.fst(mix_r.contiguous(), mix_i.contiguous(), o_r, o_i, 0)
    P.synchronize()
    lin = float((o_r[0] - (0.75 * rc2[0] - 1.5 * rc2[1])).abs().max()) / scale
    assert lin < 1e-11, lin
    P.close()


def test_odd_bandwidth_inverse_is_a_true_inverse(s2, oracles):
    """The reference's inverse is wrong for odd bw (cospml.c:270-288, SURVEY.md section 0 trap 4); ours is the exact
    transpose, so coefficients survive the round trip.  Forward parity with the reference still holds."""
    bw = 17
    O = oracles(bw)
    rc, ic = O.gen_coeffs(1000)
    g = s2.InvFSTSemiMemo(rc, ic, bw, 0)
    c = s2.FSTSemiMemo(g[0], g[1], bw, 0)
    assert relerr(cat(c), cat((rc, ic))) < 1e-12
    assert relerr(cat(c), cat(O.forward(g[0], g[1], 0))) < TOL


def test_single_field_bw1024_fly(s2, oracle_mod):
    """BASELINE configs[3]: single field at bw = 1024 with on-the-fly tables vs the reference (its Memo and Fly
    outputs are identical; Memo is used for the oracle because reference Fly takes ~100 s of CPU)."""
    bw = 1024
    O = oracle_mod.Oracle(bw, oracle_mod.best_kind())
    rc, ic = O.gen_coeffs(1000)
    want_g = O.inverse(rc, ic, 0)
    want_c = O.forward(want_g[0], want_g[1], 0)
    O.close()
    P = s2.Plan(bw, s2.FLY)
    assert relerr(cat(P.inverse(rc, ic, 0)), cat(want_g)) < TOL
    assert relerr(cat(P.forward(want_g[0], want_g[1], 0)), cat(want_c)) < TOL
    P.close()


# ------------------------------------------------------------------------------------------------ sharded single field
def _exchange(send, nranks):
    """all_to_all of equal blocks, emulated on one device: recv[d][s] = send[s][d]."""
    import torch

    blocks = [s.view(nranks, -1) for s in send]
    return [torch.stack([blocks[s][d] for s in range(nranks)]).contiguous().view(-1) for d in range(nranks)]


@pytest.mark.parametrize("bw,nranks", [(64, 4), (128, 2), (256, 8)])
def test_sharded_single_field_emulated_ranks(s2, oracles, bw, nranks):
    """latitude rings -> all-to-all -> orders, every rank's share run on one GPU and the
    exchange emulated; results must equal the reference on the full field."""
    import torch

    n = 2 * bw
    O = oracles(bw)
    rng = np.random.RandomState(11)
    rc, ic = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)  # fully complex coefficients
    want_g = O.inverse(rc, ic, 0)
    want_c = O.forward(want_g[0], want_g[1], 0)
    plans = [s2.ShardedPlan(bw, r, nranks) for r in range(nranks)]
    nr, blk = plans[0].rings, plans[0].block_doubles
    assert nr == n // nranks and sum(len(p.orders) for p in plans) == bw
    dev = "cuda"
    # forward
    gr, gi = torch.tensor(want_g[0], device=dev), torch.tensor(want_g[1], device=dev)
    send = [torch.zeros(nranks * blk, device=dev, dtype=torch.float64) for _ in range(nranks)]
    for r, P in enumerate(plans):
        P.fst_rings(gr[r * nr:(r + 1) * nr].contiguous(), gi[r * nr:(r + 1) * nr].contiguous(), send[r])
        P.synchronize()
    recv = _exchange(send, nranks)
    out_r = torch.full((bw * bw,), float("nan"), device=dev, dtype=torch.float64)
    out_i = torch.full_like(out_r, float("nan"))
    for r, P in enumerate(plans):
        P.fst_orders(recv[r], out_r, out_i)
        P.synchronize()
    assert relerr(cat((out_r.cpu().numpy(), out_i.cpu().numpy())), cat(want_c)) < TOL  # also: every slot written
    masks = np.stack([P.owned_coefficient_mask() for P in plans])
    assert (masks.sum(0) == 1).all()
    # inverse
    cr, ci = torch.tensor(rc, device=dev), torch.tensor(ic, device=dev)
    send = [torch.zeros(nranks * blk, device=dev, dtype=torch.float64) for _ in range(nranks)]
    for r, P in enumerate(plans):
        P.inv_fst_orders(cr, ci, send[r])
        P.synchronize()
    recv = _exchange(send, nranks)
    og_r = torch.zeros(n, n, device=dev, dtype=torch.float64)
    og_i = torch.zeros_like(og_r)
    for r, P in enumerate(plans):
        a = torch.zeros(nr, n, device=dev, dtype=torch.float64)
        b = torch.zeros_like(a)
        P.inv_fst_rings(recv[r], a, b)
        P.synchronize()
        og_r[r * nr:(r + 1) * nr], og_i[r * nr:(r + 1) * nr] = a, b
    assert relerr(cat((og_r.cpu().numpy(), og_i.cpu().numpy())), cat(want_g)) < TOL
    for P in plans:
        P.close()


def test_sharded_single_field_nccl_two_gpus():
    """Real exchange: two ranks, NCCL all_to_all over NVLink (skipped on a single-GPU box)."""
    import json
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(root, "tools", "bench_single_field.py"), "--bw", "256",
           "--steps", "2", "--warmup", "1", "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["sharded_vs_single_gpu_rel_err"] < 1e-12
    assert res["idempotence_rel_err"] < 1e-10
