"""CPU oracle for the S2kit spherical-harmonic-transform hot path -- TEST INFRASTRUCTURE ONLY.

Two checkers live here, behind one Python interface:

* ``kind="ref"``  -- ``oracle/_ref/libs2kit_ref.so``: the reference's own C sources compiled unmodified
  (in place from /root/reference, see ``oracle/Makefile``) against the FFTW-API stub in
  ``oracle/fftw_stub``; it reproduces the reference's four golden convolution files to 2e-16.
* ``kind="port"`` -- ``oracle/liboracle.so``: the independent restatement ``oracle/s2_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product (``s2kit_b200``) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libs2kit_ref.so")
PORT_SO = os.path.join(_HERE, "liboracle.so")

_P = ctypes.POINTER(ctypes.c_double)
COMPLEX, REAL = 0, 1


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


def build(ref=True, port=True):
    """Compile the checkers (building the checker is not using it)."""
    targets = []
    if port:
        targets.append("port")
    if ref and os.path.isdir("/root/reference/src"):
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)
    # the reference's test mains relinked against the product library (needs the product .so to exist)
    if ref and os.path.isdir("/root/reference/test") and os.path.exists(
            os.path.join(_HERE, "..", "s2kit_b200", "libs2kit_cuda.so")):
        subprocess.run(["make", "-s", "-C", _HERE, "relink"], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = REF_SO if kind == "ref" else PORT_SO
    if not os.path.exists(path):
        raise FileNotFoundError(f"oracle library {path} is not built (run `make -C oracle`)")
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    ci = ctypes.c_int
    if kind == "ref":
        L.ref_ctx_create.restype = vp
        L.ref_ctx_create.argtypes = [ci, ci, ci]
        L.ref_ctx_destroy.argtypes = [vp]
        L.ref_ctx_table.restype = _P
        L.ref_ctx_table.argtypes = [vp, ci]
        L.ref_ctx_trans_table.restype = _P
        L.ref_ctx_trans_table.argtypes = [vp, ci]
        L.ref_ctx_weights.restype = _P
        L.ref_ctx_weights.argtypes = [vp]
        for name in ("ref_fst_memo", "ref_inv_fst_memo", "ref_fzt_memo", "ref_fst_fly", "ref_inv_fst_fly",
                     "ref_fzt_fly"):
            getattr(L, name).argtypes = [vp, _P, _P, _P, _P, ci]
        L.ref_conv_memo.argtypes = [ci, _P, _P, _P, _P, _P, _P]
        L.ref_conv_fly.argtypes = [ci, _P, _P, _P, _P, _P, _P]
        L.ref_trans_mult.argtypes = [ci, _P, _P, _P, _P, _P, _P]
        L.ref_dlt_semi.argtypes = [vp, _P, ci, _P]
        L.ref_inv_dlt_semi.argtypes = [vp, _P, ci, _P]
        L.ref_gen_cos_pml_table.argtypes = [ci, ci, _P]
        L.ref_gen_coeffs.argtypes = [ci, ctypes.c_long, _P, _P]
        L.ref_bench_pairs.restype = ctypes.c_double
        L.ref_bench_pairs.argtypes = [vp, ci, ci, ctypes.c_long, ci, ci, _P]
    else:
        L.orc_create.restype = vp
        L.orc_create.argtypes = [ci]
        L.orc_destroy.argtypes = [vp]
        L.orc_table.restype = _P
        L.orc_table.argtypes = [vp, ci]
        L.orc_weights_ptr.restype = _P
        L.orc_weights_ptr.argtypes = [vp]
        for name in ("orc_forward", "orc_inverse", "orc_zonal"):
            getattr(L, name).argtypes = [vp, _P, _P, _P, _P, ci]
        L.orc_conv.argtypes = [vp, _P, _P, _P, _P, _P, _P]
        L.orc_spectral_multiply.argtypes = [ci, _P, _P, _P, _P, _P, _P]
        L.orc_cos_table.argtypes = [ci, ci, _P]
        L.orc_gen_coeffs.argtypes = [ci, ctypes.c_long, _P, _P]
        L.orc_total_len.restype = ctypes.c_long
    _libs[kind] = L
    return L


def table_size(m, bw):
    """TableSize(m, bw) (cospml.c:39-59), even or odd bw."""
    return sum((l - 1) // 2 + 1 if m % 2 else l // 2 + 1 for l in range(m, bw))


def coef_index(m, l, bw):
    """IndexOfHarmonicCoeff (util.c:42-49)."""
    if m >= 0:
        return m * bw - (m * (m - 1)) // 2 + (l - m)
    big = bw - 1
    return (big * (big + 3)) // 2 + 1 + ((big + m) * (big + m + 1)) // 2 + (l - abs(m))


class Oracle:
    """Forward / inverse / zonal / convolution through one of the two CPU checkers.

    All arrays are float64; grids are (2bw, 2bw) latitude-major, coefficient arrays have bw*bw entries in
    the reference's order.  ``variant`` picks the reference's Memo or Fly entry points (``kind="ref"``
    only; the port has a single code path -- the reference's Memo and Fly outputs are identical).
    """

    def __init__(self, bw, kind="ref", variant="memo", tables=True):
        self.bw, self.kind, self.variant = bw, kind, variant
        self.L = _load(kind)
        if kind == "ref":
            self.h = self.L.ref_ctx_create(bw, bw, 1 if (tables and variant == "memo") else 0)
        else:
            self.h = self.L.orc_create(bw)

    def close(self):
        if self.h:
            (self.L.ref_ctx_destroy if self.kind == "ref" else self.L.orc_destroy)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fn(self, memo, fly, port):
        if self.kind == "port":
            return getattr(self.L, port)
        return getattr(self.L, memo if self.variant == "memo" else fly)

    def forward(self, rdata, idata, data_format=COMPLEX):
        bw = self.bw
        rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
        self._fn("ref_fst_memo", "ref_fst_fly", "orc_forward")(
            self.h, _p(np.ascontiguousarray(rdata)), _p(np.ascontiguousarray(idata)), _p(rc), _p(ic), data_format)
        return rc, ic

    def inverse(self, rco, ico, data_format=COMPLEX):
        n = 2 * self.bw
        rd, idt = np.zeros((n, n)), np.zeros((n, n))
        self._fn("ref_inv_fst_memo", "ref_inv_fst_fly", "orc_inverse")(
            self.h, _p(np.ascontiguousarray(rco)), _p(np.ascontiguousarray(ico)), _p(rd), _p(idt), data_format)
        return rd, idt

    def zonal(self, rdata, idata, data_format=REAL):
        bw = self.bw
        rr, ir = np.zeros(2 * bw), np.zeros(2 * bw)  # reference clears 2bw entries of ires
        self._fn("ref_fzt_memo", "ref_fzt_fly", "orc_zonal")(
            self.h, _p(np.ascontiguousarray(rdata)), _p(np.ascontiguousarray(idata)), _p(rr), _p(ir), data_format)
        return rr[:bw].copy(), ir[:bw].copy()

    def conv(self, rdata, idata, rfilter, ifilter):
        n = 2 * self.bw
        rr, ir = np.zeros((n, n)), np.zeros((n, n))
        args = [_p(np.ascontiguousarray(a)) for a in (rdata, idata, rfilter, ifilter)] + [_p(rr), _p(ir)]
        if self.kind == "port":
            self.L.orc_conv(self.h, *args)
        elif self.variant == "memo":
            self.L.ref_conv_memo(self.bw, *args)
        else:
            self.L.ref_conv_fly(self.bw, *args)
        return rr, ir

    def spectral_multiply(self, rd, idt, rf, ifl):
        bw = self.bw
        rr, ir = np.zeros(bw * bw), np.zeros(bw * bw)
        fn = self.L.orc_spectral_multiply if self.kind == "port" else self.L.ref_trans_mult
        fn(bw, _p(rd), _p(idt), _p(rf), _p(ifl), _p(rr), _p(ir))
        return rr, ir

    def table(self, m):
        """Packed cosine-series table of order m in the reference layout (copy)."""
        size = table_size(m, self.bw)
        out = np.zeros(size + 2 * self.bw)
        if self.kind == "ref":
            self.L.ref_gen_cos_pml_table(self.bw, m, _p(out))
        else:
            self.L.orc_cos_table(self.bw, m, _p(out))
        return out[:size].copy()

    def weights(self):
        ptr = self.L.ref_ctx_weights(self.h) if self.kind == "ref" else self.L.orc_weights_ptr(self.h)
        return np.ctypeslib.as_array(ptr, shape=(4 * self.bw,)).copy()

    def gen_coeffs(self, seed):
        bw = self.bw
        rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
        (self.L.ref_gen_coeffs if self.kind == "ref" else self.L.orc_gen_coeffs)(bw, seed, _p(rc), _p(ic))
        return rc, ic

    def bench_pairs(self, nfun, nthreads, seed0=1000, data_format=COMPLEX):
        """Wall seconds for nfun inverse+forward pairs on nthreads host threads (kind="ref" only)."""
        assert self.kind == "ref"
        busy = np.zeros(1)
        wall = self.L.ref_bench_pairs(self.h, nfun, nthreads, seed0, data_format,
                                      0 if self.variant == "memo" else 1, _p(busy))
        return wall, float(busy[0])


def best_kind():
    """The strongest checker available: the compiled reference if present, else the port."""
    if have_ref():
        return "ref"
    if have_port():
        return "port"
    raise FileNotFoundError("no oracle library built: run `make -C oracle`")
