"""Stage-by-stage parity diagnostics on a GPU box (not a test: prints a table, never aborts early).

Usage: python tools/gpu_diag.py [bw ...]   -> relative max-abs errors of every entry point vs the CPU oracle.
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import s2kit_b200 as s2  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if not np.all(np.isfinite(a)):
        return float("nan")
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def line(name, val, extra=""):
    flag = "ok " if (val == val and val < 1e-10) else "BAD"
    print(f"  [{flag}] {name:42s} {val:.3e} {extra}", flush=True)


def attempt(name, fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        print(f"  [EXC] {name}: {e}")
        traceback.print_exc()


def diag(bw, variant=s2.MEMO):
    kind = oracle.best_kind()
    print(f"== bw {bw} variant {'memo' if variant == s2.MEMO else 'fly'} (oracle: {kind})", flush=True)
    t0 = time.time()
    O = oracle.Oracle(bw, kind)
    print(f"   oracle tables {time.time() - t0:.2f}s", flush=True)
    t0 = time.time()
    P = s2.Plan(bw, variant, max_batch=4)
    P.synchronize()
    print(f"   plan create {time.time() - t0:.2f}s, table {P.table_bytes() / 1e6:.1f} MB", flush=True)
    n = 2 * bw

    def tables():
        ms = sorted(set([0, 1, 2, 3, bw // 2, bw // 2 + 1, bw - 2, bw - 1]))
        for m in ms:
            if 0 <= m < bw:
                line(f"table m={m}", rel(P.table(m), O.table(m)))

    attempt("tables", tables)
    rc, ic = O.gen_coeffs(1000)
    rng = np.random.RandomState(7)

    def dlt():
        for m in sorted(set([0, 1, 2, bw // 2, bw - 1])):
            col = rng.uniform(-1, 1, n)
            want = np.zeros(bw)
            O.L.ref_dlt_semi(O.h, oracle._p(col), m, oracle._p(want)) if kind == "ref" else None
            if kind != "ref":
                return
            got = P.dlt_semi(col, m)[0]
            line(f"dlt_semi m={m}", rel(got, want[: bw - m]))
            co = rng.uniform(-1, 1, bw - m)
            want2 = np.zeros(n)
            O.L.ref_inv_dlt_semi(O.h, oracle._p(co), m, oracle._p(want2))
            got2 = P.inv_dlt_semi(co, m)[0]
            line(f"inv_dlt_semi m={m}", rel(got2, want2))

    attempt("dlt", dlt)

    def transforms():
        for fmt, tag in ((s2.COMPLEX, "complex"), (s2.REAL, "real")):
            rd, idt = O.inverse(rc, ic, fmt)
            g = P.inverse(rc, ic, fmt)
            line(f"inv_fst {tag}", rel(np.stack(g), np.stack([rd, idt])))
            fr, fi = O.forward(rd, idt, fmt)
            c = P.forward(rd, idt, fmt)
            line(f"fst {tag}", rel(np.concatenate(c), np.concatenate([fr, fi])),
                 f"(round trip vs seeded coeffs {rel(np.concatenate(c), np.concatenate([rc, ic])):.2e})")
        # fully complex field
        r2, i2 = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)
        rd, idt = O.inverse(r2, i2, 0)
        line("inv_fst full-complex", rel(np.stack(P.inverse(r2, i2, 0)), np.stack([rd, idt])))
        fr, fi = O.forward(rd, idt, 0)
        line("fst full-complex", rel(np.concatenate(P.forward(rd, idt, 0)), np.concatenate([fr, fi])))
        # batch of 3 through the device-pointer path is covered by tests; here host batch
        rdb = np.stack([rd, 2 * rd, idt])
        idb = np.stack([idt, rd, -idt])
        cb = P.forward(rdb, idb, 0)
        w0 = O.forward(rdb[2], idb[2], 0)
        line("fst batch[2]", rel(np.concatenate([cb[0][2], cb[1][2]]), np.concatenate(w0)))

    attempt("transforms", transforms)

    def zonal_conv():
        sig = rng.uniform(-1, 1, (n, n))
        fil = rng.uniform(-1, 1, (n, n))
        z = np.zeros((n, n))
        zr, zi = O.zonal(fil, z, 1)
        rr, ir = np.zeros(bw), np.zeros(bw)
        P.fzt(fil, z, rr, ir, s2.REAL)
        line("fzt real", rel(rr, zr))
        zr, zi = O.zonal(fil, sig, 0)
        P.fzt(fil, sig, rr, ir, s2.COMPLEX)
        line("fzt complex", rel(np.concatenate([rr, ir]), np.concatenate([zr, zi])))
        want = O.conv(sig, z, fil, z)
        gr, gi = np.zeros((n, n)), np.zeros((n, n))
        P.conv(sig, z, fil, z, gr, gi)
        line("conv", rel(gr, want[0]), f"imag max {np.abs(gi).max():.1e}")

    attempt("zonal/conv", zonal_conv)
    P.close()
    O.close()


if __name__ == "__main__":
    print(s2.lib().s2kit_cuda_version().decode())
    try:
        print("fp64 peak:", s2.measure_fp64_peak())
        print("copy GB/s:", s2.measure_copy_bw(0, 1 << 30))
    except Exception as e:  # noqa: BLE001
        print("peak measurement failed:", e)
    bws = [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 64, 8, 17, 256]
    for bw in bws:
        attempt(f"bw {bw}", lambda: diag(bw))
    if "--fly" in sys.argv:
        for bw in bws[:2]:
            attempt(f"fly bw {bw}", lambda: diag(bw, s2.FLY))
