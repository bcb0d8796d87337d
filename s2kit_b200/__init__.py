"""s2kit_b200 -- B200-native FP64 spherical harmonic transforms behind S2kit's C API.

The product is ``libs2kit_cuda.so`` (hand-written CUDA for sm_100a + a C host layer exporting the
reference's API: ``FSTSemiMemo``, ``InvFSTSemiMemo``, ``FZTSemiMemo``, ``ConvOn2SphereSemiMemo``, the
``-SemiFly`` twins and the table/weight helpers).  This package is the thin Python binding used by the tests
and the benchmark: ``ctypes`` over the C-ABI declared in ``include/s2kit_cuda.h`` / ``include/s2kit.h``.
PyTorch is only used by callers for device memory and streams; no torch types cross the boundary.

There is no CPU fallback: importing works anywhere (the library is loaded lazily), but every transform
raises ``S2kitCudaError`` when the extension is missing or no CUDA device is present.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2kit_cuda.so")

COMPLEX, REAL = 0, 1          # DataFormat, include/s2kit/util.h:10-13
MEMO, FLY = 0, 1
HOST, DEVICE = 0, 1

KERNEL_KINDS = ["phi_fft_fwd", "dct_fwd", "legendre_fwd", "legendre_inv", "dct_inv", "phi_fft_inv", "table_gen",
                "zonal", "spectral_mul", "fused_fwd", "fused_inv"]

_P = ctypes.POINTER(ctypes.c_double)
_lib = None


class S2kitCudaError(RuntimeError):
    pass


def build(force=False):
    from . import buildlib as _b

    return _b.build(force=force)


def lib():
    """The loaded C-ABI library (raises if it has not been built -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S2kitCudaError(f"{LIB_PATH} is missing: run `python -m s2kit_b200.buildlib` (no CPU fallback exists)")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cl, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t
    L.s2kit_cuda_plan_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci]
    L.s2kit_cuda_plan_create_sharded.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci]
    L.s2kit_cuda_plan_destroy.argtypes = [vp]
    L.s2kit_cuda_plan_clone.argtypes = [ctypes.POINTER(vp), vp, ci]
    L.s2kit_cuda_multi_create.argtypes = [ctypes.POINTER(vp), ci, ci, ctypes.POINTER(ci)]
    L.s2kit_cuda_multi_destroy.argtypes = [vp]
    L.s2kit_cuda_multi_ngpu.argtypes = [vp]
    L.s2kit_cuda_multi_table_bytes_per_gpu.restype = cs
    L.s2kit_cuda_multi_table_bytes_per_gpu.argtypes = [vp]
    L.s2kit_cuda_multi_fst.argtypes = [vp, vp, vp, vp, vp]
    L.s2kit_cuda_multi_inv_fst.argtypes = [vp, vp, vp, vp, vp]
    L.s2kit_cuda_multi_buffers.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                           ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.s2kit_cuda_multi_run.argtypes = [vp, ci, ci, _P]
    L.s2kit_cuda_plan_set_stream.argtypes = [vp, vp]
    L.s2kit_cuda_plan_stream.restype = vp
    L.s2kit_cuda_plan_stream.argtypes = [vp]
    L.s2kit_cuda_synchronize.argtypes = [vp]
    L.s2kit_cuda_plan_bw.argtypes = [vp]
    L.s2kit_cuda_plan_table_bytes.restype = cs
    L.s2kit_cuda_plan_table_bytes.argtypes = [vp]
    L.s2kit_cuda_plan_table_stream_bytes.restype = cs
    L.s2kit_cuda_plan_table_stream_bytes.argtypes = [vp]
    L.s2kit_cuda_fst.argtypes = [vp, vp, vp, vp, vp, ci, cl, cl, ci, ci]
    L.s2kit_cuda_inv_fst.argtypes = [vp, vp, vp, vp, vp, ci, cl, cl, ci, ci]
    L.s2kit_cuda_fzt.argtypes = [vp, vp, vp, vp, vp, ci, cl, cl, ci, ci]
    L.s2kit_cuda_conv.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, cl, cl, ci]
    L.s2kit_cuda_trans_mult.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, cl, ci]
    L.s2kit_cuda_dlt_semi.argtypes = [vp, vp, ci, vp, ci, ci]
    L.s2kit_cuda_inv_dlt_semi.argtypes = [vp, vp, ci, vp, ci, ci]
    L.s2kit_cuda_dlt_naive.argtypes = [vp, ci, ci, vp, vp, vp, ci]
    L.s2kit_cuda_inv_dlt_naive.argtypes = [vp, ci, ci, vp, vp, ci]
    L.s2kit_cuda_fst_rings.argtypes = [vp, vp, vp, vp]
    L.s2kit_cuda_fst_orders.argtypes = [vp, vp, vp, vp]
    L.s2kit_cuda_inv_fst_orders.argtypes = [vp, vp, vp, vp]
    L.s2kit_cuda_inv_fst_rings.argtypes = [vp, vp, vp, vp]
    L.s2kit_cuda_shard_info.argtypes = [vp, ctypes.POINTER(cl), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.s2kit_cuda_shard_layout.argtypes = [ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.s2kit_cuda_table_export.argtypes = [vp, ci, vp]
    L.s2kit_cuda_table_generate.argtypes = [vp, ci, vp]
    L.s2kit_cuda_profile_enable.argtypes = [vp, ci]
    L.s2kit_cuda_profile_get.argtypes = [vp, _P, ctypes.POINTER(cl)]
    L.s2kit_cuda_profile_reset.argtypes = [vp]
    L.s2kit_cuda_measure_fp64_peak.argtypes = [ci, _P, _P]
    L.s2kit_cuda_measure_copy_bw.argtypes = [ci, cs, _P]
    L.s2kit_cuda_host_alloc.restype = vp
    L.s2kit_cuda_host_alloc.argtypes = [cs]
    L.s2kit_cuda_host_free.argtypes = [vp]
    L.s2kit_cuda_last_error.restype = ctypes.c_char_p
    L.s2kit_cuda_version.restype = ctypes.c_char_p
    # reference API (include/s2kit.h)
    for name in ("TableSize", "Reduced_SpharmonicTableSize", "Reduced_Naive_TableSize", "TableOffset", "RowSize"):
        getattr(L, name).argtypes = [ci, ci]
    L.Spharmonic_TableSize.argtypes = [ci]
    L.Transpose_RowSize.argtypes = [ci, ci, ci]
    L.IndexOfHarmonicCoeff.argtypes = [ci, ci, ci]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise S2kitCudaError(f"{what}: {lib().s2kit_cuda_last_error().decode()}")


def _ptr(a):
    """(address, where) of a numpy array (host) or of any object with data_ptr() on a CUDA device."""
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "float64 C-contiguous arrays only"
        return a.ctypes.data, HOST
    if hasattr(a, "data_ptr"):
        assert str(a.dtype) == "torch.float64" and a.is_contiguous()
        return a.data_ptr(), (DEVICE if a.is_cuda else HOST)
    raise TypeError(type(a))


def _same_where(*ws):
    assert len(set(ws)) == 1, "all arrays of one call must live on the same side (host or device)"
    return ws[0]


def table_size(m, bw):
    return lib().TableSize(m, bw)


def index_of_harmonic_coeff(m, l, bw):
    return lib().IndexOfHarmonicCoeff(m, l, bw)


class Plan:
    """A transform plan for one bandwidth (s2kit_cuda_plan_create).

    Arrays may be numpy float64 (host; copied in and out, synchronous) or torch.float64 CUDA tensors
    (device; asynchronous on the plan's stream).  Batched arrays are (batch, 2bw, 2bw) grids and
    (batch, bw*bw) coefficient arrays.
    """

    def __init__(self, bw, variant=MEMO, max_batch=1, device=0):
        self.bw, self.n, self.variant = bw, 2 * bw, variant
        h = ctypes.c_void_p()
        _check(lib().s2kit_cuda_plan_create(ctypes.byref(h), bw, variant, max_batch, device), "plan_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().s2kit_cuda_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self, max_batch=1):
        """A plan that shares this plan's device tables but owns its stream and workspaces (one per host thread)."""
        q = Plan.__new__(Plan)
        q.bw, q.n, q.variant = self.bw, self.n, self.variant
        h = ctypes.c_void_p()
        _check(lib().s2kit_cuda_plan_clone(ctypes.byref(h), self.h, max_batch), "plan_clone")
        q.h = h
        return q

    # -- stream / measurement
    def set_stream(self, cuda_stream_ptr):
        _check(lib().s2kit_cuda_plan_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        _check(lib().s2kit_cuda_synchronize(self.h), "synchronize")

    def table_bytes(self):
        return lib().s2kit_cuda_plan_table_bytes(self.h)

    def table_stream_bytes(self):
        """Table bytes one transform reads (one copy of the tiles)."""
        return lib().s2kit_cuda_plan_table_stream_bytes(self.h)

    def profile(self, on=True):
        _check(lib().s2kit_cuda_profile_enable(self.h, 1 if on else 0), "profile_enable")
        _check(lib().s2kit_cuda_profile_reset(self.h), "profile_reset")

    def profile_get(self):
        ms = (ctypes.c_double * len(KERNEL_KINDS))()
        cnt = (ctypes.c_long * len(KERNEL_KINDS))()
        _check(lib().s2kit_cuda_profile_get(self.h, ms, cnt), "profile_get")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_KINDS)}

    # -- transforms
    def _batch(self, a, per):
        size = a.size if isinstance(a, np.ndarray) else a.numel()
        assert size % per == 0
        return size // per

    def fst(self, rdata, idata, rcoeffs, icoeffs, data_format=COMPLEX):
        gs, cs = self.n * self.n, self.bw * self.bw
        batch = self._batch(rdata, gs)
        (a, w1), (b, w2), (c, w3), (d, w4) = map(_ptr, (rdata, idata, rcoeffs, icoeffs))
        _check(lib().s2kit_cuda_fst(self.h, a, b, c, d, batch, gs, cs, data_format, _same_where(w1, w2, w3, w4)),
               "fst")

    def inv_fst(self, rcoeffs, icoeffs, rdata, idata, data_format=COMPLEX):
        gs, cs = self.n * self.n, self.bw * self.bw
        batch = self._batch(rcoeffs, cs)
        (a, w1), (b, w2), (c, w3), (d, w4) = map(_ptr, (rcoeffs, icoeffs, rdata, idata))
        _check(lib().s2kit_cuda_inv_fst(self.h, a, b, c, d, batch, cs, gs, data_format, _same_where(w1, w2, w3, w4)),
               "inv_fst")

    def fzt(self, rdata, idata, rres, ires, data_format=REAL):
        gs = self.n * self.n
        batch = self._batch(rdata, gs)
        (a, w1), (b, w2), (c, w3), (d, w4) = map(_ptr, (rdata, idata, rres, ires))
        _check(lib().s2kit_cuda_fzt(self.h, a, b, c, d, batch, gs, self.bw, data_format, _same_where(w1, w2, w3, w4)),
               "fzt")

    def conv(self, rdata, idata, rfilter, ifilter, rres, ires, shared_filter=False):
        gs = self.n * self.n
        batch = self._batch(rdata, gs)
        ptrs = list(map(_ptr, (rdata, idata, rfilter, ifilter, rres, ires)))
        where = _same_where(*[w for _, w in ptrs])
        _check(lib().s2kit_cuda_conv(self.h, *[p for p, _ in ptrs], batch, gs, 0 if shared_filter else gs, where),
               "conv")

    def trans_mult(self, rd, idt, rf, ifl, rres, ires):
        cs = self.bw * self.bw
        batch = self._batch(rd, cs)
        ptrs = list(map(_ptr, (rd, idt, rf, ifl, rres, ires)))
        where = _same_where(*[w for _, w in ptrs])
        _check(lib().s2kit_cuda_trans_mult(self.h, *[p for p, _ in ptrs], batch, cs, where), "trans_mult")

    def dlt_semi(self, data, m):
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, self.n)
        out = np.zeros((data.shape[0], self.bw - m))
        _check(lib().s2kit_cuda_dlt_semi(self.h, data.ctypes.data, m, out.ctypes.data, data.shape[0], HOST),
               "dlt_semi")
        return out

    def inv_dlt_semi(self, coeffs, m):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64).reshape(-1, self.bw - m)
        out = np.zeros((coeffs.shape[0], self.n))
        _check(lib().s2kit_cuda_inv_dlt_semi(self.h, coeffs.ctypes.data, m, out.ctypes.data, coeffs.shape[0], HOST),
               "inv_dlt_semi")
        return out

    def table(self, m):
        """Order m's cosine table in the reference's packed layout (GenerateCosPmlTable)."""
        out = np.zeros(table_size(m, self.bw))
        _check(lib().s2kit_cuda_table_export(self.h, m, out.ctypes.data), "table_export")
        return out

    # -- numpy conveniences (host path)
    def forward(self, rdata, idata, data_format=COMPLEX):
        rdata, idata = (np.ascontiguousarray(a, dtype=np.float64) for a in (rdata, idata))
        batch = rdata.size // (self.n * self.n)
        rc, ic = np.zeros((batch, self.bw * self.bw)), np.zeros((batch, self.bw * self.bw))
        self.fst(rdata, idata, rc, ic, data_format)
        return (rc[0], ic[0]) if rdata.ndim == 2 else (rc, ic)

    def inverse(self, rco, ico, data_format=COMPLEX):
        rco, ico = (np.ascontiguousarray(a, dtype=np.float64) for a in (rco, ico))
        batch = rco.size // (self.bw * self.bw)
        rd, idt = np.zeros((batch, self.n, self.n)), np.zeros((batch, self.n, self.n))
        self.inv_fst(rco, ico, rd, idt, data_format)
        return (rd[0], idt[0]) if rco.ndim == 1 else (rd, idt)


def shard_layout(bw, nranks, rank):
    """(orders, rows) owned by `rank` -- host-side arithmetic only, no GPU needed (s2kit_cuda_shard_layout)."""
    orders = (ctypes.c_int * bw)()
    rows = (ctypes.c_int * (2 * bw))()
    cnt = lib().s2kit_cuda_shard_layout(bw, nranks, rank, orders, rows)
    if cnt < 0:
        raise S2kitCudaError(f"unsupported split: bw={bw} over {nranks} ranks")
    return list(orders[:cnt]), list(rows[: 2 * bw // nranks])


class ShardedPlan:
    """One rank's share of a single-field transform split over `nranks` GPUs (s2kit_cuda_plan_create_sharded).

    forward : fst_rings(local rings) -> all_to_all(sendbuf -> recvbuf) -> fst_orders(recvbuf) -> own coefficients
    inverse : inv_fst_orders(coefficients) -> all_to_all -> inv_fst_rings(recvbuf) -> local rings
    All arrays are torch.float64 CUDA tensors; the exchange buffers have nranks * block_doubles entries.
    """

    def __init__(self, bw, rank, nranks, device=0):
        self.bw, self.n, self.rank, self.nranks = bw, 2 * bw, rank, nranks
        h = ctypes.c_void_p()
        _check(lib().s2kit_cuda_plan_create_sharded(ctypes.byref(h), bw, MEMO, device, rank, nranks),
               "plan_create_sharded")
        self.h = h
        blk, nr, rows = ctypes.c_long(), ctypes.c_int(), ctypes.c_int()
        _check(lib().s2kit_cuda_shard_info(h, ctypes.byref(blk), ctypes.byref(nr), ctypes.byref(rows)), "shard_info")
        self.block_doubles, self.rings = blk.value, nr.value
        self.orders, self.rows = shard_layout(bw, nranks, rank)

    def close(self):
        if getattr(self, "h", None):
            lib().s2kit_cuda_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(lib().s2kit_cuda_plan_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        _check(lib().s2kit_cuda_synchronize(self.h), "synchronize")

    def table_bytes(self):
        return lib().s2kit_cuda_plan_table_bytes(self.h)

    def table_stream_bytes(self):
        return lib().s2kit_cuda_plan_table_stream_bytes(self.h)

    def fst_rings(self, rdata, idata, sendbuf):
        _check(lib().s2kit_cuda_fst_rings(self.h, rdata.data_ptr(), idata.data_ptr(), sendbuf.data_ptr()), "fst_rings")

    def fst_orders(self, recvbuf, rcoeffs, icoeffs):
        _check(lib().s2kit_cuda_fst_orders(self.h, recvbuf.data_ptr(), rcoeffs.data_ptr(), icoeffs.data_ptr()),
               "fst_orders")

    def inv_fst_orders(self, rcoeffs, icoeffs, sendbuf):
        _check(lib().s2kit_cuda_inv_fst_orders(self.h, rcoeffs.data_ptr(), icoeffs.data_ptr(), sendbuf.data_ptr()),
               "inv_fst_orders")

    def inv_fst_rings(self, recvbuf, rdata, idata):
        _check(lib().s2kit_cuda_inv_fst_rings(self.h, recvbuf.data_ptr(), rdata.data_ptr(), idata.data_ptr()),
               "inv_fst_rings")

    def owned_coefficient_mask(self):
        """Boolean mask over the bw*bw coefficient positions this rank produces / consumes."""
        mask = np.zeros(self.bw * self.bw, dtype=bool)
        for m in self.orders:
            for sm in ((m,) if m == 0 else (m, -m)):
                a = index_of_harmonic_coeff(sm, m, self.bw)
                mask[a:a + self.bw - m] = True
        return mask


class MultiPlan:
    """One single-field transform on `ngpu` GPUs of this process (s2kit_cuda_multi_*, csrc/multi.cu): rings and orders
    split over the devices, the ring <-> order exchange done by the DCT kernels on peer-mapped memory."""

    def __init__(self, bw, ngpu, devices=None):
        self.bw, self.n, self.ngpu = bw, 2 * bw, ngpu
        h = ctypes.c_void_p()
        devs = (ctypes.c_int * ngpu)(*devices) if devices is not None else None
        _check(lib().s2kit_cuda_multi_create(ctypes.byref(h), bw, ngpu, devs), "multi_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().s2kit_cuda_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def table_bytes_per_gpu(self):
        return lib().s2kit_cuda_multi_table_bytes_per_gpu(self.h)

    def forward(self, rdata, idata):
        rd, idt = (np.ascontiguousarray(a, dtype=np.float64) for a in (rdata, idata))
        rc, ic = np.full(self.bw * self.bw, np.nan), np.full(self.bw * self.bw, np.nan)
        _check(lib().s2kit_cuda_multi_fst(self.h, rd.ctypes.data, idt.ctypes.data, rc.ctypes.data, ic.ctypes.data),
               "multi_fst")
        return rc, ic

    def inverse(self, rco, ico):
        rc, ic = (np.ascontiguousarray(a, dtype=np.float64) for a in (rco, ico))
        rd, idt = np.full((self.n, self.n), np.nan), np.full((self.n, self.n), np.nan)
        _check(lib().s2kit_cuda_multi_inv_fst(self.h, rc.ctypes.data, ic.ctypes.data, rd.ctypes.data, idt.ctypes.data),
               "multi_inv_fst")
        return rd, idt

    def run(self, inverse=False, iters=1):
        """`iters` device-resident transforms on the plan's own buffers; returns ms per transform (device time)."""
        ms = ctypes.c_double()
        _check(lib().s2kit_cuda_multi_run(self.h, 1 if inverse else 0, iters, ctypes.byref(ms)), "multi_run")
        return ms.value


def measure_fp64_peak(device=0):
    a, b = ctypes.c_double(), ctypes.c_double()
    _check(lib().s2kit_cuda_measure_fp64_peak(device, ctypes.byref(a), ctypes.byref(b)), "measure_fp64_peak")
    return {"fma_tflops": a.value, "dmma_tflops": b.value}


def measure_copy_bw(device=0, nbytes=1 << 30):
    g = ctypes.c_double()
    _check(lib().s2kit_cuda_measure_copy_bw(device, nbytes, ctypes.byref(g)), "measure_copy_bw")
    return g.value


# ---------------------------------------------------------------------------------------------------------
# The reference's own entry points, called through the drop-in symbols exactly as a C caller would
# (workspace / FFTW plan / host table arguments are passed as NULL: the GPU layer ignores them).
def _np(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_P)


def _fst_like(symbol, has_table):
    def call(rdata, idata, bw, data_format=COMPLEX, cutoff=None):
        rd, prd = _np(rdata)
        idt, pid = _np(idata)
        rc, ic = np.zeros(bw * bw), np.zeros(bw * bw)
        fn = getattr(lib(), symbol)
        args = [prd, pid, rc.ctypes.data_as(_P), ic.ctypes.data_as(_P), ctypes.c_int(bw)]
        if has_table:
            args.append(None)
        args += [None, ctypes.c_int(data_format), ctypes.c_int(bw if cutoff is None else cutoff), None, None, None]
        fn.restype = None
        fn(*args)
        return rc, ic

    return call


def _inv_like(symbol, has_table):
    def call(rcoeffs, icoeffs, bw, data_format=COMPLEX, cutoff=None):
        rc, prc = _np(rcoeffs)
        ic, pic = _np(icoeffs)
        n = 2 * bw
        rd, idt = np.zeros((n, n)), np.zeros((n, n))
        fn = getattr(lib(), symbol)
        args = [prc, pic, rd.ctypes.data_as(_P), idt.ctypes.data_as(_P), ctypes.c_int(bw)]
        if has_table:
            args.append(None)
        args += [None, ctypes.c_int(data_format), ctypes.c_int(bw if cutoff is None else cutoff), None, None]
        fn.restype = None
        fn(*args)
        return rd, idt

    return call


def _fzt_like(symbol, has_table):
    def call(rdata, idata, bw, data_format=REAL):
        rd, prd = _np(rdata)
        idt, pid = _np(idata)
        rr, ir = np.zeros(2 * bw), np.zeros(2 * bw)
        fn = getattr(lib(), symbol)
        args = [prd, pid, rr.ctypes.data_as(_P), ir.ctypes.data_as(_P), ctypes.c_int(bw)]
        if has_table:
            args.append(None)
        args += [None, ctypes.c_int(data_format), None, None]
        fn.restype = None
        fn(*args)
        return rr[:bw].copy(), ir[:bw].copy()

    return call


def _conv_like(symbol):
    def call(rdata, idata, rfilter, ifilter, bw):
        arrs = [_np(a) for a in (rdata, idata, rfilter, ifilter)]
        n = 2 * bw
        rr, ir = np.zeros((n, n)), np.zeros((n, n))
        fn = getattr(lib(), symbol)
        fn.restype = None
        fn(*[p for _, p in arrs], rr.ctypes.data_as(_P), ir.ctypes.data_as(_P), ctypes.c_int(bw), None)
        return rr, ir

    return call


FSTSemiMemo = _fst_like("FSTSemiMemo", True)
FSTSemiFly = _fst_like("FSTSemiFly", False)
InvFSTSemiMemo = _inv_like("InvFSTSemiMemo", True)
InvFSTSemiFly = _inv_like("InvFSTSemiFly", False)
FZTSemiMemo = _fzt_like("FZTSemiMemo", True)
FZTSemiFly = _fzt_like("FZTSemiFly", False)
ConvOn2SphereSemiMemo = _conv_like("ConvOn2SphereSemiMemo")
ConvOn2SphereSemiFly = _conv_like("ConvOn2SphereSemiFly")


def GenerateWeightsForDLT(bw):
    w = np.zeros(4 * bw)
    fn = lib().GenerateWeightsForDLT
    fn.restype = None
    fn(ctypes.c_int(bw), w.ctypes.data_as(_P))
    return w


def GenerateCosPmlTable(bw, m):
    out = np.zeros(table_size(m, bw))
    fn = lib().GenerateCosPmlTable
    fn.restype = None
    fn(ctypes.c_int(bw), ctypes.c_int(m), out.ctypes.data_as(_P), None)
    return out


def GeneratePmlTable(bw, m):
    """Theta-space table of order m, [bw - m][2bw] (pml.c:41-79); host-side setup."""
    out = np.zeros((bw - m) * 2 * bw)
    fn = lib().GeneratePmlTable
    fn.restype = None
    fn(ctypes.c_int(bw), ctypes.c_int(m), out.ctypes.data_as(_P), None)
    return out


def Pmm_L2(m, eval_points):
    pts, pp = _np(eval_points)
    out = np.zeros(pts.size)
    fn = lib().Pmm_L2
    fn.restype = None
    fn(ctypes.c_int(m), pp, ctypes.c_int(pts.size), out.ctypes.data_as(_P))
    return out


def DLTNaive(data, bw, m, weights, pml_table):
    """naive.c:35-60 through the relinkable C API: the dense product runs on the GPU."""
    (d, pd), (w, pw), (t, pt) = _np(data), _np(weights), _np(pml_table)
    out = np.zeros(bw - m)
    fn = lib().DLTNaive
    fn.restype = None
    fn(pd, ctypes.c_int(bw), ctypes.c_int(m), pw, out.ctypes.data_as(_P), pt, None)
    return out


def InvDLTNaive(coeffs, bw, m, pml_table):
    """naive.c:77-95."""
    (c, pc), (t, pt) = _np(coeffs), _np(pml_table)
    out = np.zeros(2 * bw)
    fn = lib().InvDLTNaive
    fn.restype = None
    fn(pc, ctypes.c_int(bw), ctypes.c_int(m), out.ctypes.data_as(_P), pt)
    return out


def TransposeCosPmlTable(bw, m, table):
    t, pt = _np(table)
    out = np.zeros(table_size(m, bw))
    fn = lib().TransposeCosPmlTable
    fn.restype = None
    fn(ctypes.c_int(bw), ctypes.c_int(m), pt, out.ctypes.data_as(_P))
    return out


def release():
    """Frees the plans cached by the reference-API entry points."""
    fn = lib().s2kit_compat_release
    fn.restype = None
    fn()
