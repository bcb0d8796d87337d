"""Round trip coefficients -> grid -> coefficients on the GPU and report the error per order m (diagnostic)."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import s2kit_b200 as s2

bw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mmax = int(sys.argv[2]) if len(sys.argv) > 2 else bw - 1
rng = np.random.RandomState(5)
rc = np.zeros(bw * bw)
ic = np.zeros(bw * bw)
for m in range(-mmax, mmax + 1):
    a = s2.index_of_harmonic_coeff(m, abs(m), bw)
    cnt = bw - abs(m)
    rc[a:a + cnt] = rng.uniform(-1, 1, cnt)
    ic[a:a + cnt] = rng.uniform(-1, 1, cnt)
P = s2.Plan(bw, s2.MEMO, max_batch=1)
g = P.inverse(rc, ic, 0)
c = P.forward(g[0], g[1], 0)
err = np.maximum(np.abs(c[0] - rc), np.abs(c[1] - ic))
print("bw", bw, "orders up to", mmax, "finite:", np.isfinite(c[0]).all(), "max err", np.nanmax(err))
rows = []
for m in range(-(bw - 1), bw):
    a = s2.index_of_harmonic_coeff(m, abs(m), bw)
    rows.append((float(np.nanmax(err[a:a + bw - abs(m)])), m))
rows.sort(reverse=True)
print("worst orders:", [(m, f"{e:.2e}") for e, m in rows[:12]])
print("median order error:", f"{np.median([e for e, _ in rows]):.2e}")
