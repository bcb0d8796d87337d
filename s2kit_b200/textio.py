"""Text files in the formats of the reference's example programs (host-side helpers, no GPU involved).

The reference has no I/O functions in its library; its example mains read and write plain text:
  * samples / coefficients "code format": one value per line, real part then imaginary part of each element
    (test/test_s2_semi_memo_fwd.c:119-124 reads a 2bw x 2bw grid this way, :141-143 writes bw*bw coefficients with
    "%.15f"; test/test_s2_semi_memo_inv.c:131-134, test/test_conv_semi_memo.c:70-80 likewise);
  * real-valued grids, one value per line: the signal, filter and result files of the convolution examples
    (test/test_conv_semi_memo.c:70-80, data/s64.dat, data/f64.dat, data/o64_conv_semi_memo_original.dat);
  * coefficients "human-readable format": "l = %d\\t m = %d\\t %.15f + %.15f I" per (l, m), degrees ascending, orders
    -l..l (test/test_s2_semi_memo_fwd.c:144-149).
These helpers let a Python caller exchange files with those programs (or with the same mains relinked against
libs2kit_cuda.so, tests/test_relinked_reference_mains.py).
"""
import re

import numpy as np


def index_of_harmonic_coeff(m, l, bw):
    """IndexOfHarmonicCoeff (src/util/util.c:42-49)."""
    if m >= 0:
        return m * bw - (m * (m - 1)) // 2 + (l - m)
    big = bw - 1
    return (big * (big + 3)) // 2 + 1 + ((big + m) * (big + m + 1)) // 2 + (l - abs(m))


def read_interleaved(path, count=None):
    """Real and imaginary parts of `count` elements stored one value per line, real part first."""
    v = np.loadtxt(path, dtype=np.float64).ravel()
    if count is not None:
        if v.size < 2 * count:
            raise ValueError(f"{path}: {v.size} values, expected {2 * count}")
        v = v[:2 * count]
    if v.size % 2:
        raise ValueError(f"{path}: odd number of values")
    return np.ascontiguousarray(v[0::2]), np.ascontiguousarray(v[1::2])


def read_real(path, count=None):
    """Real-valued samples, one per line (the signal / filter / output files of test/test_conv_semi_memo.c:70-80, :129)."""
    v = np.loadtxt(path, dtype=np.float64).ravel()
    if count is not None:
        if v.size < count:
            raise ValueError(f"{path}: {v.size} values, expected {count}")
        v = v[:count]
    return np.ascontiguousarray(v)


def write_real(path, values):
    np.savetxt(path, np.asarray(values, dtype=np.float64).ravel(), fmt="%.16f")


def read_grid(path, bw):
    """A 2bw x 2bw sample grid (latitude-major), as the reference's mains read it."""
    n = 2 * bw
    re_, im_ = read_interleaved(path, n * n)
    return re_.reshape(n, n), im_.reshape(n, n)


def write_interleaved(path, real, imag):
    """"%.15f" per line, real part then imaginary part of each element."""
    real, imag = np.asarray(real, dtype=np.float64).ravel(), np.asarray(imag, dtype=np.float64).ravel()
    if real.shape != imag.shape:
        raise ValueError("real and imaginary parts differ in size")
    out = np.empty(2 * real.size)
    out[0::2], out[1::2] = real, imag
    np.savetxt(path, out, fmt="%.15f")


def write_coeffs(path, rcoeffs, icoeffs, bw, human_readable=False):
    """bw*bw coefficients in the library's own order (code format) or listed by (l, m)."""
    rc, ic = np.asarray(rcoeffs, dtype=np.float64).ravel(), np.asarray(icoeffs, dtype=np.float64).ravel()
    if rc.size != bw * bw or ic.size != bw * bw:
        raise ValueError("expected bw*bw coefficients")
    if not human_readable:
        return write_interleaved(path, rc, ic)
    with open(path, "w") as f:
        for l in range(bw):
            for m in range(-l, l + 1):
                i = index_of_harmonic_coeff(m, l, bw)
                f.write("l = %d\t m = %d\t %.15f + %.15f I\n" % (l, m, rc[i], ic[i]))


_HUMAN = re.compile(r"l = (-?\d+)\s+m = (-?\d+)\s+(\S+) \+ (\S+) I")


def read_coeffs(path, bw):
    """Either coefficient format back into the library's order."""
    with open(path) as f:
        first = f.readline()
    if not first.startswith("l ="):
        return read_interleaved(path, bw * bw)
    rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
    with open(path) as f:
        for ln in f:
            mt = _HUMAN.match(ln)
            if not mt:
                raise ValueError(f"{path}: cannot parse {ln!r}")
            l, m = int(mt.group(1)), int(mt.group(2))
            i = index_of_harmonic_coeff(m, l, bw)
            rc[i], ic[i] = float(mt.group(3)), float(mt.group(4))
    return rc, ic
