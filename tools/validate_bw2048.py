"""One-off parity check of BASELINE.json configs[4] (bw = 2048) against the compiled reference (oracle/_ref).

The reference needs ~4-5 minutes of CPU and ~23 GB of RAM to build its bw = 2048 tables, so this is not part of the
test suite; run it on the GPU box (`gpurun -- python tools/validate_bw2048.py`) and keep the printed summary under
profiles/.  The reference returns NaN for orders |m| >= 2044 (P_m^m overflow, pmm.c:22-30): parity is evaluated on
|m| <= 2043, and reported separately for the band of orders whose seeds underflow.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import s2kit_b200 as s2

bw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = 2 * bw
rng = np.random.RandomState(2048)
rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
t0 = time.time()
P = s2.Plan(bw, s2.MEMO, max_batch=1)
got = P.forward(rd, idt, s2.COMPLEX)
t_gpu = time.time() - t0
P.close()
t0 = time.time()
O = oracle.Oracle(bw, "ref")
t_tab = time.time() - t0
t0 = time.time()
want = O.forward(rd, idt, s2.COMPLEX)
t_fwd = time.time() - t0
O.close()
res = {"bw": bw, "gpu_plan_plus_forward_s": t_gpu, "ref_table_build_s": t_tab, "ref_forward_s": t_fwd}
scale = max(np.nanmax(np.abs(want[0][np.isfinite(want[0])])), np.nanmax(np.abs(want[1][np.isfinite(want[1])])))
per_order = {}
worst_all = 0.0
nan_orders = []
for m in range(-(bw - 1), bw):
    a = s2.index_of_harmonic_coeff(m, abs(m), bw)
    sl = slice(a, a + bw - abs(m))
    if not (np.isfinite(want[0][sl]).all() and np.isfinite(want[1][sl]).all()):
        nan_orders.append(m)
        continue
    e = max(np.abs(got[0][sl] - want[0][sl]).max(), np.abs(got[1][sl] - want[1][sl]).max()) / scale
    per_order[m] = e
    worst_all = max(worst_all, e)
res["reference_nan_orders"] = [min(nan_orders, default=None), max(nan_orders, default=None), len(nan_orders)]
res["rel_err_all_finite_orders"] = worst_all
res["worst_orders"] = sorted(((e, m) for m, e in per_order.items()), reverse=True)[:8]
res["ours_finite_everywhere"] = bool(np.isfinite(got[0]).all() and np.isfinite(got[1]).all())
print(json.dumps(res))
