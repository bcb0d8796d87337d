"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the headers declare,
the host-side layout arithmetic and libm seeds agree with the oracle, and compute entry points fail loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def s2():
    import __graft_entry__ as ge

    ge.build()
    import s2kit_b200

    return s2kit_b200


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(s2):
    L = s2.lib()
    for header in ("s2kit_cuda.h", "s2kit.h"):
        names = declared_functions(header)
        assert len(names) > 15
        for name in names:
            assert hasattr(L, name), f"{name} declared in include/{header} but not exported"


def test_layout_helpers_match_oracle(s2, oracle_mod):
    L = s2.lib()
    for bw in (8, 16, 17, 23, 64, 256):
        for m in range(bw):
            assert L.TableSize(m, bw) == oracle_mod.table_size(m, bw)
            for l in (m, (m + bw) // 2, bw - 1):
                assert L.IndexOfHarmonicCoeff(m, l, bw) == oracle_mod.coef_index(m, l, bw)
                assert L.IndexOfHarmonicCoeff(-m, l, bw) == oracle_mod.coef_index(-m, l, bw)
                assert L.TableOffset(m, l) == sum(L.RowSize(m, d) for d in range(m, l))
        assert L.Reduced_SpharmonicTableSize(bw, bw) == sum(oracle_mod.table_size(m, bw) for m in range(bw))
        assert L.Reduced_Naive_TableSize(bw, bw // 2) == 2 * bw * sum(bw - o for o in range(bw // 2, bw))
        assert L.Spharmonic_TableSize(bw) >= L.Reduced_SpharmonicTableSize(bw, bw)
        if bw % 2 == 0:
            for m in range(bw):
                assert sum(L.Transpose_RowSize(r, m, bw) for r in range(bw)) == L.TableSize(m, bw)


def test_layout_helpers_match_reference_build(s2, oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    L, R = s2.lib(), ctypes.CDLL(oracle_mod.REF_SO)
    P = ctypes.POINTER(ctypes.c_double)
    for bw in (16, 64, 128):
        for m in range(bw):
            assert L.TableSize(m, bw) == R.TableSize(m, bw)
            for l in range(m, bw):
                assert L.TableOffset(m, l) == R.TableOffset(m, l)
                assert L.RowSize(m, l) == R.RowSize(m, l)
            for row in range(bw + 1):
                assert L.Transpose_RowSize(row, m, bw) == R.Transpose_RowSize(row, m, bw)
        assert L.Spharmonic_TableSize(bw) == R.Spharmonic_TableSize(bw)
    # TransposeCosPmlTable is pure index arithmetic on the host: same gather as the reference
    O = oracle_mod.Oracle(32, "ref")
    for m in (0, 1, 2, 15, 30, 31):
        t = O.table(m)
        want = np.zeros_like(t)
        R.TransposeCosPmlTable(32, m, t.ctypes.data_as(P), want.ctypes.data_as(P))
        assert np.array_equal(s2.TransposeCosPmlTable(32, m, t), want)


def test_host_seeds_bit_identical_to_oracle(s2, oracle_mod):
    """weights.c:32-47 through the drop-in symbol; must be bit-identical (same libm, same expression order)."""
    for bw in (16, 64, 100):
        O = oracle_mod.Oracle(bw, "port")
        assert np.array_equal(s2.GenerateWeightsForDLT(bw), O.weights())


def test_pmm_l2_is_finite_where_the_reference_overflows(s2):
    """Pmm_L2 (pmm.c:21-33) through the drop-in symbol: bit-identical to the reference's expression where that is finite
    (covered by the table tests); for m >= 2044 the reference's running product overflows to inf -- ours re-associates
    only there and must reproduce the exact values (mpmath, tests/golden/mp_high_orders.npz)."""
    mp = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mp_high_orders.npz"))
    bw = int(mp["bw"])
    theta = (2 * np.arange(2 * bw) + 1) * np.pi / (4 * bw)
    for m in range(2040, 2048):
        got = s2.Pmm_L2(m, theta)
        want = mp[f"m{m}_pmm"]
        assert np.isfinite(got).all(), m
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (m, np.abs(got - want).max())


def test_compute_fails_loudly_without_gpu(s2):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(s2.S2kitCudaError, match="no CUDA device"):
        s2.Plan(16)
    with pytest.raises(s2.S2kitCudaError):
        s2.measure_fp64_peak(0)


def test_product_does_not_touch_the_oracle():
    """The shipped path must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "s2kit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "libs2kit_ref" not in txt, f
    out = os.popen(f"ldd {os.path.join(pkg, 'libs2kit_cuda.so')}").read()
    assert "oracle" not in out and "fftw" not in out


@pytest.mark.parametrize("check,expect", [("layout_check", "layout mismatches: 0"), ("fft16_check", "exactly once: 0"),
                                          ("dct16_check", "dct16 check: OK")])
def test_host_side_checks_of_device_helpers(tmp_path, check, expect):
    """Host-only programs built from the same headers as the kernels: (1) the closed-form tile layout the persistent
    kernels use instead of dependent loads against the tabulated layout for bandwidths 2 .. 2048; (2) the register-level
    pieces of the one-warp 512-point FFT (s2k_fft16.cuh) for 32 emulated lanes against a long-double DFT; (3) the forward
    DCT pair on that FFT with its in-place shared-memory exchange and shuffle-based separation (s2k_dct16.cuh)."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / check)
    src = os.path.join(ROOT, "tests", "host_checks", check + ".cu")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-o", exe, src], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and expect in r.stdout, r.stdout + r.stderr


def test_per_file_headers_forward_to_the_api(tmp_path):
    """A caller of the reference includes "s2kit/FST_semi_memo.h" etc. (reference include/s2kit/*.h): the drop-in tree
    ships those file names, each forwarding to include/s2kit.h; every one compiles alone and declares the entry points
    of the reference header of the same name."""
    import subprocess

    expect = {"FST_semi_memo": "FSTSemiMemo", "FST_semi_fly": "InvFSTSemiFly", "cospml": "Spharmonic_Pml_Table",
              "pml": "GeneratePmlTable", "pmm": "Pmm_L2", "seminaive": "InvDLTSemi", "naive": "DLTNaive",
              "weights": "GenerateWeightsForDLT", "util": "TransMult", "chebyshev_nodes": "ChebyshevNodes"}
    for name, symbol in expect.items():
        src = tmp_path / f"use_{name}.c"
        src.write_text(f'#include "s2kit/{name}.h"\nvoid* probe(void) {{ return (void*){symbol}; }}\nDataFormat fmt = REAL;\n')
        r = subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_text_files_in_the_reference_formats(tmp_path, refdata):
    """s2kit_b200.textio: the file formats of the reference's example mains (test_s2_semi_memo_fwd.c:119-151)."""
    from s2kit_b200 import textio

    bw = 8
    rng = np.random.RandomState(3)
    rc, ic = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)
    for human in (False, True):
        path = tmp_path / f"c{int(human)}.dat"
        textio.write_coeffs(path, rc, ic, bw, human_readable=human)
        br, bi = textio.read_coeffs(path, bw)
        assert np.abs(br - rc).max() < 1e-15 + 5e-16 and np.abs(bi - ic).max() < 1e-15 + 5e-16  # "%.15f"
    lines = open(tmp_path / "c1.dat").read().splitlines()
    assert len(lines) == bw * bw and lines[0].startswith("l = 0\t m = 0\t ") and lines[1].startswith("l = 1\t m = -1\t ")
    # a grid written the way the reference's mains read it, and the reference's own s64.dat as committed
    n = 2 * bw
    g_r, g_i = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    textio.write_interleaved(tmp_path / "g.dat", g_r, g_i)
    h_r, h_i = textio.read_grid(tmp_path / "g.dat", bw)
    assert np.abs(h_r - g_r).max() < 1e-15 and np.abs(h_i - g_i).max() < 1e-15
    s64 = refdata["s64"]  # the reference's data/s64.dat: a real-valued 128 x 128 signal, one value per line
    textio.write_real(tmp_path / "s64.dat", s64)
    assert np.abs(textio.read_real(tmp_path / "s64.dat", 128 * 128) - s64).max() < 1e-15
