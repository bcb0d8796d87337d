#!/usr/bin/env python
"""Short batched workload for ncu captures: `iters` x (InvFST + FST) of `nfun` functions at bandwidth `bw`, device resident.

  ncu --set full --clock-control none --import-source on -k regex:k_fwd_uni -s 1 -c 1 -o gpurun_out/prof \
      python tools/prof_batch.py --nfun 256 --iters 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import s2kit_b200 as s2  # noqa: E402
from bench import synth_coeffs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bw", type=int, default=256)
ap.add_argument("--nfun", type=int, default=256)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--format", default="complex")
a = ap.parse_args()
dev = torch.device("cuda", 0)
fmt = s2.COMPLEX if a.format == "complex" else s2.REAL
P = s2.Plan(a.bw, s2.MEMO, max_batch=a.nfun, device=0)
P.set_stream(torch.cuda.current_stream().cuda_stream)
n = 2 * a.bw
rc, ic = synth_coeffs(torch, a.bw, a.nfun, dev, 1000)
rd = torch.empty(a.nfun, n, n, device=dev, dtype=torch.float64)
idt = torch.empty_like(rd)
rc2, ic2 = torch.empty_like(rc), torch.empty_like(ic)
for _ in range(a.iters):
    P.inv_fst(rc, ic, rd, idt, fmt)
    P.fst(rd, idt, rc2, ic2, fmt)
torch.cuda.synchronize()
print("roundtrip err", float((rc2 - rc).abs().max()), float((ic2 - ic).abs().max()))
P.close()
