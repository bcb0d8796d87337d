"""Per-kernel CUDA-event times of one forward + inverse call (device-resident): python tools/stage_times.py bw [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import s2kit_b200 as s2

bw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
variant = s2.FLY if "--fly" in sys.argv else s2.MEMO
n = 2 * bw
P = s2.Plan(bw, variant, max_batch=min(batch, 256))
P.set_stream(torch.cuda.current_stream().cuda_stream)
rd = torch.rand(batch, n, n, device="cuda", dtype=torch.float64)
idt = torch.rand(batch, n, n, device="cuda", dtype=torch.float64)
rc = torch.zeros(batch, bw * bw, device="cuda", dtype=torch.float64)
ic = torch.zeros_like(rc)
for _ in range(3):
    P.fst(rd, idt, rc, ic, 0)
    P.inv_fst(rc, ic, rd, idt, 0)
torch.cuda.synchronize()
P.profile(True)
reps = 5
for _ in range(reps):
    P.fst(rd, idt, rc, ic, 0)
    P.inv_fst(rc, ic, rd, idt, 0)
prof = P.profile_get()
print(f"bw {bw} batch {batch} table {P.table_bytes()/1e9:.2f} GB")
for k, (ms, cnt) in prof.items():
    if cnt:
        print(f"  {k:14s} {ms/reps:9.3f} ms per call ({cnt//reps} launches)")
