"""The reference's own test programs (test/*.c, unmodified, compiled in place by `make -C oracle relink`) linked against
libs2kit_cuda.so instead of the reference objects -- the drop-in claim as an executable check.
The binaries are built in the build container (they need /root/reference) and travel to the GPU box under
oracle/_ref/relink/; the data files are regenerated from tests/golden/reference_data.npz."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "relink")


def _need(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    return path


def _write(path, arr, fmt="%.16f"):
    np.savetxt(path, np.asarray(arr).ravel(), fmt=fmt)


def _run(cmd, cwd):
    out = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:] + out.stdout[-2000:]
    return out.stdout


@pytest.mark.parametrize("variant", ["memo", "fly"])
@pytest.mark.parametrize("bw", [64, 128])
def test_conv_mains_reproduce_goldens(tmp_path, refdata, variant, bw):
    """dist/test.sh:39-61: test_conv_semi_{memo,fly} on data/s*.dat, f*.dat; diff against the *_original.dat goldens."""
    exe = _need(f"test_conv_semi_{variant}")
    _write(tmp_path / "s.dat", refdata[f"s{bw}"])
    _write(tmp_path / "f.dat", refdata[f"f{bw}"])
    _run([exe, "s.dat", "f.dat", "o.dat", str(bw)], tmp_path)
    got = np.loadtxt(tmp_path / "o.dat")
    gold = refdata[f"o{bw}_conv_semi_{variant}_original"]
    assert np.abs(got - gold).max() <= 1e-15 + 5e-17  # test.sh eps 1e-15; the file holds 16 decimals


def test_forward_main_on_known_answers(tmp_path, refdata, oracle_mod):
    """test_s2_semi_memo_fwd on the sampled Y_l^m files (dist/S2kitHowTo.pdf 2.4.2)."""
    exe = _need("test_s2_semi_memo_fwd")
    for name, bw, expect in (("y20_bw8", 8, {(0, 2): 1.0}), ("y31_bw8", 8, {(1, 3): 1.0}),
                             ("y43_bw23", 23, {(3, 4): complex(2 ** 0.5, np.pi)}),
                             ("yMix_bw17", 17, {(1, 1): 1.0, (-2, 5): complex(3.0, -2.0)})):
        _write(tmp_path / "in.dat", refdata[name])
        _run([exe, "in.dat", "out.dat", str(bw)], tmp_path)
        c = np.loadtxt(tmp_path / "out.dat").reshape(-1, 2)
        want = np.zeros(bw * bw, dtype=complex)
        for (m, l), v in expect.items():
            want[oracle_mod.coef_index(m, l, bw)] = v
        assert np.abs((c[:, 0] + 1j * c[:, 1]) - want).max() < 1e-13, name  # file holds 15 decimals


def test_inverse_main_round_trip(tmp_path, vectors):
    """test_s2_semi_memo_inv on committed seeded coefficients vs the committed reference grid."""
    exe = _need("test_s2_semi_memo_inv")
    bw = 16
    rc, ic = vectors["coef_seed1000_bw16_r"], vectors["coef_seed1000_bw16_i"]
    _write(tmp_path / "c.dat", np.stack([rc, ic], axis=1), fmt="%.17g")
    _run([exe, "c.dat", "g.dat", str(bw)], tmp_path)
    g = np.loadtxt(tmp_path / "g.dat").reshape(-1, 2)
    want = np.stack([vectors["inv_complex_bw16_r"].ravel(), vectors["inv_complex_bw16_i"].ravel()], axis=1)
    assert np.abs(g - want).max() / np.abs(want).max() < 1e-12


@pytest.mark.parametrize("exe_name,args", [("test_s2_semi_memo", ["64", "2"]), ("test_s2_semi_fly", ["64", "2"]),
                                            ("test_DLT_semi", ["3", "128", "4"]),
                                            ("test_DLT_naive", ["3", "128", "4"])])
def test_round_trip_mains_report_small_errors(tmp_path, exe_name, args):
    """The self-consistency mains print their round-trip errors (HowTo 2.2: ~1e-12 at bw 123)."""
    exe = _need(exe_name)
    out = _run([exe] + args, tmp_path)
    m = re.search(r"Average r-o error:\s+([0-9.eE+-]+)", out)
    assert m, out[-1500:]
    assert float(m.group(1)) < 1e-11, out[-1500:]
