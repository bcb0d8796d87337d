import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import s2kit_b200 as s2
bw = int(sys.argv[1]); n = 2*bw
dev = torch.device("cuda", 0)
P = s2.Plan(bw, s2.MEMO, max_batch=1, device=0)
P.set_stream(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(7)
rd = torch.rand(n, n, generator=g, device=dev, dtype=torch.float64)*2-1
idt = torch.rand(n, n, generator=g, device=dev, dtype=torch.float64)*2-1
cr = torch.zeros(bw*bw, device=dev, dtype=torch.float64); ci = torch.zeros_like(cr)
def timeit(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/k
f = timeit(lambda: P.fst(rd, idt, cr, ci, 0)); i = timeit(lambda: P.inv_fst(cr, ci, rd, idt, 0))
P.profile(True); P.fst(rd, idt, cr, ci, 0); P.inv_fst(cr, ci, rd, idt, 0); P.synchronize()
print(json.dumps({"bw": bw, "ms_forward": f, "ms_inverse": i, "kernels": {k: round(v[0],4) for k, v in P.profile_get().items() if v[1]}}))
