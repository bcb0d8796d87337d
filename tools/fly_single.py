#!/usr/bin/env python
"""One single-field FSTSemiFly + InvFSTSemiFly at bandwidth --bw on cuda:0 (BASELINE.json configs[3]): times per direction
with CUDA events and the per-kernel split from the plan's profiler.  Also the command to put under ncu for K7."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import s2kit_b200 as s2  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bw", type=int, default=1024)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
bw, n = a.bw, 2 * a.bw
dev = torch.device("cuda", 0)
P = s2.Plan(bw, s2.FLY, max_batch=1, device=0)
P.set_stream(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev)
g.manual_seed(7)
rd = torch.rand(1, n, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
idt = torch.rand(1, n, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
rc = torch.zeros(1, bw * bw, device=dev, dtype=torch.float64)
ic = torch.zeros_like(rc)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


ms_f = timeit(lambda: P.fst(rd, idt, rc, ic, s2.COMPLEX))
ms_i = timeit(lambda: P.inv_fst(rc, ic, rd, idt, s2.COMPLEX))
P.profile(True)
P.fst(rd, idt, rc, ic, s2.COMPLEX)
P.inv_fst(rc, ic, rd, idt, s2.COMPLEX)
P.synchronize()
prof = P.profile_get()
print(json.dumps({"bw": bw, "variant": "fly", "ms_forward": ms_f, "ms_inverse": ms_i,
                  "kernel_ms": {k: v[0] for k, v in prof.items() if v[1]}}))
P.close()
