"""GPU tests of the drop-in boundary beyond single calls: concurrent callers (the reference is re-entrant), plan clones,
the single-process multi-GPU transform, non-power-of-two bandwidths above 512, the four double** table builders'
layouts against the compiled reference, and the tuning switches of the library (each run in a child process)."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def s2():
    import s2kit_b200

    s2kit_b200.lib()
    yield s2kit_b200
    s2kit_b200.release()


def cat(pair):
    return np.concatenate([np.asarray(pair[0]).ravel(), np.asarray(pair[1]).ravel()])


# ------------------------------------------------------------------------------------------------ threads
def test_concurrent_callers_of_the_reference_api(s2, oracle_mod):
    """Eight host threads call InvFSTSemiMemo / FSTSemiMemo concurrently on their own buffers -- the way
    oracle/ref_harness.c:225-262 drives the (re-entrant) reference -- at two bandwidths, several rounds each.
    ctypes releases the GIL, so the calls really overlap inside the library."""
    kind = oracle_mod.best_kind()
    want = {}
    for bw in (32, 64):
        O = oracle_mod.Oracle(bw, kind)
        for t in range(8):
            rc, ic = O.gen_coeffs(3000 + t)
            g = O.inverse(rc, ic, 0)
            want[(bw, t)] = (rc, ic, g, O.forward(g[0], g[1], 0))
        O.close()
    errs = []

    def work(t):
        try:
            for rnd in range(3):
                for bw in ((32, 64) if (t + rnd) % 2 else (64, 32)):
                    rc, ic, g, c = want[(bw, t)]
                    got_g = s2.InvFSTSemiMemo(rc, ic, bw, 0)
                    got_c = s2.FSTSemiMemo(g[0], g[1], bw, 0)
                    e = max(relerr(cat(got_g), cat(g)), relerr(cat(got_c), cat(c)))
                    if not e < TOL:
                        errs.append((t, rnd, bw, e))
        except Exception as ex:  # noqa: BLE001
            errs.append((t, repr(ex)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs[:4]


def test_plan_cache_eviction_under_use(s2, oracle_mod):
    """More distinct bandwidths than cache entries (8), from four threads: entries in use must never be destroyed."""
    O = {bw: oracle_mod.Oracle(bw, "port") for bw in (8, 10, 12, 14, 16, 18, 20, 22, 24, 26)}
    cases = {}
    for bw, o in O.items():
        rc, ic = o.gen_coeffs(10 + bw)
        cases[bw] = (rc, ic, o.inverse(rc, ic, 0))
        o.close()
    errs = []

    def work(t):
        bws = sorted(cases)
        for k in range(20):
            bw = bws[(3 * k + t) % len(bws)]
            rc, ic, g = cases[bw]
            e = relerr(cat(s2.InvFSTSemiMemo(rc, ic, bw, 0)), cat(g))
            if not e < TOL:
                errs.append((t, bw, e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs[:4]


def test_plan_clone_shares_tables_and_runs_concurrently(s2, oracle_mod):
    bw = 64
    O = oracle_mod.Oracle(bw, oracle_mod.best_kind())
    P = s2.Plan(bw, s2.MEMO, max_batch=4)
    Q = P.clone(max_batch=2)
    F = s2.Plan(bw, s2.FLY).clone()  # a Fly clone owns its scratch table
    data = []
    for t in range(3):
        rc, ic = O.gen_coeffs(500 + t)
        data.append((rc, ic, O.inverse(rc, ic, 0)))
    errs = []

    def work(plan, t):
        for _ in range(5):
            rc, ic, g = data[t]
            e = relerr(cat(plan.inverse(rc, ic, 0)), cat(g))
            if not e < TOL:
                errs.append((t, e))

    threads = [threading.Thread(target=work, args=(pl, t)) for t, pl in enumerate((P, Q, F))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    assert relerr(Q.table(5), P.table(5)) == 0.0
    Q.close()
    F.close()
    P.close()
    O.close()


# ------------------------------------------------------------------------------------------------ multi-GPU, one process
@pytest.mark.parametrize("bw,ngpu", [(64, 1), (256, 1), (512, 1), (128, 2), (512, 2), (256, 4), (256, 8)])
def test_multi_gpu_single_process_transform(s2, oracle_mod, bw, ngpu):
    """s2kit_cuda_multi_*: rings and orders split over the GPUs, the exchange done by the DCT kernels on peer-mapped
    memory.  ngpu = 1 runs the same code path (peer pointers = own buffers) on the driver's single-GPU box."""
    import torch

    if torch.cuda.device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    O = oracle_mod.Oracle(bw, oracle_mod.best_kind())
    rng = np.random.RandomState(bw + ngpu)
    rc, ic = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)  # fully complex coefficients
    want_g = O.inverse(rc, ic, 0)
    want_c = O.forward(want_g[0], want_g[1], 0)
    M = s2.MultiPlan(bw, ngpu)
    for _ in range(3):  # both ring-buffer generations
        got_c = M.forward(want_g[0], want_g[1])
        assert relerr(cat(got_c), cat(want_c)) < TOL
        got_g = M.inverse(rc, ic)
        assert relerr(cat(got_g), cat(want_g)) < TOL
    assert M.run(inverse=False, iters=5) > 0.0 and M.run(inverse=True, iters=5) > 0.0
    got_c = M.forward(want_g[0], want_g[1])  # still correct after the back-to-back runs
    assert relerr(cat(got_c), cat(want_c)) < TOL
    M.close()
    O.close()


def test_reference_api_routes_large_fields_to_all_gpus(s2, oracle_mod):
    """S2KIT_CUDA_NGPU > 1: FSTSemiMemo / InvFSTSemiMemo split a bw >= 512 field over the GPUs (child process: the
    switch is read per call but the plans are cached per process)."""
    import torch

    ngpu = min(torch.cuda.device_count(), 8)
    ngpu = 1 << (ngpu.bit_length() - 1)
    code = (
        "import numpy as np, os, sys; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import oracle, s2kit_b200 as s2\n"
        "bw = 512\n"
        "L = oracle._load('port'); rc, ic = np.zeros(bw*bw), np.zeros(bw*bw); L.orc_gen_coeffs(bw, 1000, oracle._p(rc), oracle._p(ic))\n"
        "g = s2.InvFSTSemiMemo(rc, ic, bw, 0)\n"
        "c = s2.FSTSemiMemo(g[0], g[1], bw, 0)\n"
        "large = np.load(os.path.join(%r, 'tests', 'golden', 'oracle_vectors_large.npz'))\n"
        "gs, cs = (int(v) for v in large['bw512_strides'])\n"
        "e1 = np.abs(g[0].ravel()[::gs] - large['bw512_inv_sample_r']).max() / np.abs(large['bw512_inv_sample_r']).max()\n"
        "e2 = np.abs(c[0][::cs] - large['bw512_fwd_sample_r']).max() / np.abs(large['bw512_fwd_sample_r']).max()\n"
        "print('ERR', e1, e2); s2.release()\n" % (ROOT, ROOT, ROOT))
    env = dict(os.environ, S2KIT_CUDA_NGPU=str(ngpu), S2KIT_CUDA_MULTI_MIN_BW="512", S2KIT_CUDA_MULTI_PROF="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    e1, e2 = (float(v) for v in r.stdout.split("ERR")[1].split()[:2])
    assert e1 < TOL and e2 < TOL
    if ngpu > 1:  # S2KIT_CUDA_MULTI_PROF: the per-stage event times of every GPU on stderr
        assert "[s2kit multi] gpu 0:" in r.stderr and "K3" in r.stderr and "K6" in r.stderr


# ------------------------------------------------------------------------------------------------ any bandwidth
def test_non_power_of_two_bandwidth_above_512(s2, oracle_mod):
    """The reference accepts any bandwidth; bandwidths that are not powers of two run on the direct O(n^2) kernels.
    A full reference transform at such a size costs minutes on the CPU (the FFT stub is O(n^2) there), so parity is
    checked per order: the reference's own GenerateCosPmlTable rows contracted in numpy with the DCT of the weighted
    spectral row (seminaive.c:153-198 restated with scipy's REDFT10), plus the coefficient round trip."""
    import scipy.fft

    bw = 516
    n = 2 * bw
    O = oracle_mod.Oracle(bw, oracle_mod.best_kind(), tables=False)
    rc, ic = O.gen_coeffs(1000)
    P = s2.Plan(bw)
    rd, idt = P.inverse(rc, ic, 0)
    fr, fi = P.forward(rd, idt, 0)
    assert relerr(cat((fr, fi)), cat((rc, ic))) < TOL  # round trip of band-limited data
    w = s2.GenerateWeightsForDLT(bw)
    F = np.fft.fft(rd + 1j * idt, axis=1) * (np.sqrt(2.0 * np.pi) / n)  # [latitude j][order row m']
    for m in (0, 1, 258, 515):
        tab = O.table(m)
        assert relerr(P.table(m), tab) < 1e-12
        for sgn in ((1,) if m == 0 else (1, -1)):
            col = F[:, m if sgn > 0 else n - m] * w[(n if m & 1 else 0):(n if m & 1 else 0) + n]
            want = np.zeros(bw - m, dtype=complex)
            for part, x in ((0, col.real), (1, col.imag)):
                y = scipy.fft.dct(x, type=2)  # REDFT10
                y[0] *= np.sqrt(0.5)
                y *= 1.0 / np.sqrt(2.0 * n)
                at = 0
                for l in range(m, bw):
                    par, ln = (l - m) & 1, ((l - 1) // 2 + 1 if m & 1 else l // 2 + 1)
                    v = float(np.dot(y[par:par + 2 * ln:2], tab[at:at + ln]))
                    want[l - m] += v if part == 0 else 1j * v
                    at += ln
            if sgn < 0 and m & 1:
                want = -want  # (-1)^m on the negative orders, FST_semi_memo.c:181-186
            a0 = s2.index_of_harmonic_coeff(sgn * m, m, bw)
            got = fr[a0:a0 + bw - m] + 1j * fi[a0:a0 + bw - m]
            assert np.abs(got - want).max() / np.abs(fr).max() < TOL, (m, sgn)
    P.close()
    O.close()


# ------------------------------------------------------------------------------------------------ host-visible tables
def test_table_builders_match_reference_layout(s2, oracle_mod):
    """Spharmonic_Pml_Table, SemiNaive_Naive_Pml_Table (cutoff < bw: cosine tables below the cutoff, theta-space tables
    from it on) and their transposes (cospml.c:387-518): the concatenated resultspace and every pointer offset against
    the reference's own builders."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    L, R = s2.lib(), ctypes.CDLL(oracle_mod.REF_SO)
    P = ctypes.POINTER(ctypes.c_double)
    PP = ctypes.POINTER(P)
    ci = ctypes.c_int
    bw, cutoff = 32, 8
    for lib_ in (L, R):
        lib_.Spharmonic_Pml_Table.restype = PP
        lib_.Transpose_Spharmonic_Pml_Table.restype = PP
        lib_.SemiNaive_Naive_Pml_Table.restype = PP
        lib_.Transpose_SemiNaive_Naive_Pml_Table.restype = PP
        for name in ("Reduced_Naive_TableSize", "Reduced_SpharmonicTableSize"):
            getattr(lib_, name).argtypes = [ci, ci]
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]

    def build(lib_):
        out = {}
        ws = np.zeros(64 * bw)
        size = lib_.Spharmonic_TableSize(bw) + 2 * bw
        a, at = np.zeros(size), np.zeros(size)
        t = lib_.Spharmonic_Pml_Table(ci(bw), a.ctypes.data_as(P), ws.ctypes.data_as(P))
        tt = lib_.Transpose_Spharmonic_Pml_Table(t, ci(bw), at.ctypes.data_as(P))
        base, baset = a.ctypes.data, at.ctypes.data
        out["sph"] = (a, [(ctypes.addressof(t[m].contents) - base) // 8 for m in range(bw)])
        out["sph_t"] = (at, [(ctypes.addressof(tt[m].contents) - baset) // 8 for m in range(bw)])
        size = lib_.Reduced_Naive_TableSize(bw, cutoff) + lib_.Reduced_SpharmonicTableSize(bw, cutoff) + 2 * bw
        b, bt = np.zeros(size), np.zeros(size)
        u = lib_.SemiNaive_Naive_Pml_Table(ci(bw), ci(cutoff), b.ctypes.data_as(P), ws.ctypes.data_as(P))
        ut = lib_.Transpose_SemiNaive_Naive_Pml_Table(u, ci(bw), ci(cutoff), bt.ctypes.data_as(P), ws.ctypes.data_as(P))
        out["mix"] = (b, [(ctypes.addressof(u[m].contents) - b.ctypes.data) // 8 for m in range(bw)])
        out["mix_t"] = (bt, [(ctypes.addressof(ut[m].contents) - bt.ctypes.data) // 8 for m in range(bw)])
        for q in (t, tt, u, ut):
            libc.free(ctypes.cast(q, ctypes.c_void_p))  # the caller frees the pointer arrays (test_s2_semi_memo.c:270-271)
        return out

    ours, ref = build(L), build(R)
    for key in ("sph", "sph_t", "mix", "mix_t"):
        assert ours[key][1] == ref[key][1], key  # same offsets of every order inside the resultspace
        assert relerr(ours[key][0], ref[key][0]) < 1e-13, key


# ------------------------------------------------------------------------------------------------ tuning switches
@pytest.mark.parametrize("env", [{"S2KIT_CUDA_NO_TMA": "1"}, {"S2KIT_CUDA_FFT16": "0"}, {"S2KIT_CUDA_L2PERSIST": "1"},
                                 {"S2KIT_CUDA_NC": "16"}, {"S2KIT_CUDA_L2PF_MAX": "0"}, {"S2KIT_CUDA_UNI": "0"},
                                 {"S2KIT_CUDA_UNI_INV": "1"}, {"S2KIT_CUDA_FLOW": "1"}, {"S2KIT_CUDA_K4_QUAD": "0"}, {"S2KIT_CUDA_UNI_LEAD": "2", "S2KIT_CUDA_UNI_SLEEP": "0", "S2KIT_CUDA_UNI_CAP": "8"},
                                 {"S2KIT_CUDA_TMA_TABLES": "1"},
                                 {"S2KIT_CUDA_TABLE_LCH": "64", "S2KIT_CUDA_FLY_RING_MB": "8"}])
def test_tuning_switches_keep_parity(env):
    """Every environment switch the library reads selects code that must still match the oracle: the wide-panel batch
    test (bw 128 / 256, both formats), the table test and the Fly test re-run in a child process per switch."""
    e = dict(os.environ, **env)
    sel = "test_batched_wide_panels_match_oracle or test_fly_equals_memo or (test_tables_match_oracle and 64)"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m",
                        "gpu", "-k", sel], env=e, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "4 passed" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("env", [{"S2KIT_CUDA_TABLE_COPIES": "2"}, {"S2KIT_CUDA_TMA_TABLES": "1"}, {"S2KIT_CUDA_TMA_TABLES": "2"}])
def test_large_bandwidth_switches_keep_parity(env):
    """Switches that only matter at bw >= 512 (second table copy, TMA-staged table tiles instead of cp.async): the bw = 512
    reference-sample test in a child process."""
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_large.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_memo_large_bw_vs_reference_samples and 512"], env=e, cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "1 passed" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("env", [{"S2KIT_CUDA_TABLE_FULL": "1"}, {"S2KIT_CUDA_PHI_ROWS": "2048"}, {"S2KIT_CUDA_PHI_ROWS": "0"},
                                 {"S2KIT_CUDA_TABLE_LCH": "32"}, {"S2KIT_CUDA_FLY_RING_MB": "256"}])
def test_bw1024_switches_keep_parity(env):
    """Switches that only matter at bw >= 1024: the full-length table generator instead of the half-grid one, the staged
    longitude transforms (ring-major plane + tiled transpose) switched on at n = 2048 / off everywhere, and smaller
    generator work units (checkpoint spacing), and several Fly order groups per transform.  The bw = 1024 Memo and Fly
    reference tests in a child process."""
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_large.py"),
                        os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu", "-k",
                        "(test_memo_large_bw_vs_reference_samples and 1024) or test_single_field_bw1024_fly"], env=e, cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout, r.stdout[-500:]


def test_sub_batch_split_over_two_streams():
    """S2KIT_CUDA_SPLIT=2 (plan.cu, run_split: a chunk cut into sub-batches that alternate between two streams -- measured
    slower than one stream, kept as an experiment): the split path must give the results of the unsplit one.  Child
    processes, because the switch is read once per plan."""
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r); import s2kit_b200 as s2; from bench import synth_coeffs\n"
        "bw, batch = 256, 256; n = 2 * bw; dev = torch.device('cuda', 0)\n"
        "P = s2.Plan(bw, s2.MEMO, max_batch=batch)\n"
        "rc, ic = synth_coeffs(torch, bw, batch, dev, 7)\n"
        "rd = torch.empty(batch, n, n, device=dev, dtype=torch.float64); idt = torch.empty_like(rd)\n"
        "P.inv_fst(rc, ic, rd, idt, 0); rc2, ic2 = torch.empty_like(rc), torch.empty_like(ic); P.fst(rd, idt, rc2, ic2, 0); P.synchronize()\n"
        "print('SUM %%.17g %%.17g %%.3e' %% (float(rd.double().sum()), float(rc2.abs().sum()), float((rc2 - rc).abs().max())))\n" % ROOT)
    outs = []
    for split in ("1", "2"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, S2KIT_CUDA_SPLIT=split), cwd=ROOT,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("SUM")][-1].split())
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]  # bit-identical: same kernels on the same data
    assert float(outs[1][3]) < 1e-10
