import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import s2kit_b200 as s2
bw, G = 2048, 8
dev = torch.device("cuda", 0)
P = s2.ShardedPlan(bw, 0, G, device=0)
P.set_stream(torch.cuda.current_stream().cuda_stream)
blk = P.block_doubles
recv = torch.rand(G * blk, device=dev, dtype=torch.float64)
send = torch.zeros_like(recv)
cr = torch.zeros(bw * bw, device=dev, dtype=torch.float64); ci = torch.zeros_like(cr)
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print(json.dumps({"rowsplit_ctas": os.environ.get("S2KIT_CUDA_ROWSPLIT_CTAS"), "fst_orders_ms": timeit(lambda: P.fst_orders(recv, cr, ci)), "inv_fst_orders_ms": timeit(lambda: P.inv_fst_orders(cr, ci, send))}))
