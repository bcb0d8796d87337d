#!/usr/bin/env python
"""Single large-bandwidth field split over the GPUs of one box (BASELINE.json configs[4]).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
      tools/bench_single_field.py --bw 2048 [--check]

Each rank owns 2bw/G latitude rings and bw/G orders (m paired with bw-1-m) with their tables only; forward =
K1 on the rings -> NCCL all_to_all (ring-major -> order-major) -> K2+K3 on the orders; inverse mirrors it.
--check compares with the unsharded single-GPU plan on rank 0 (needs the whole table on one GPU).
Prints one JSON line on rank 0; times are CUDA-event times, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import s2kit_b200 as s2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bw", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bw, n = a.bw, 2 * a.bw
    P = s2.ShardedPlan(bw, rank, world, device=local)
    P.set_stream(torch.cuda.current_stream().cuda_stream)
    nr, blk = P.rings, P.block_doubles
    g = torch.Generator(device=dev)
    g.manual_seed(1234)  # same field on every rank; each keeps its rings
    full_r = torch.rand(n, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    full_i = torch.rand(n, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    if world > 1:  # identical field on every rank regardless of per-device generator state
        dist.broadcast(full_r, 0)
        dist.broadcast(full_i, 0)
    my_r, my_i = full_r[rank * nr:(rank + 1) * nr].clone(), full_i[rank * nr:(rank + 1) * nr].clone()  # not views
    if not a.check:
        del full_r, full_i
    send = torch.zeros(world * blk, device=dev, dtype=torch.float64)
    recv = torch.zeros_like(send)
    cr = torch.zeros(bw * bw, device=dev, dtype=torch.float64)
    ci = torch.zeros_like(cr)
    out_r, out_i = torch.zeros_like(my_r), torch.zeros_like(my_i)

    def exchange():
        if world > 1:
            dist.all_to_all_single(recv, send)
        else:
            recv.copy_(send)

    def forward():
        P.fst_rings(my_r, my_i, send)
        exchange()
        P.fst_orders(recv, cr, ci)

    def inverse():
        P.inv_fst_orders(cr, ci, send)
        exchange()
        P.inv_fst_rings(recv, out_r, out_i)

    def timeit(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_fwd = timeit(forward)
    ms_inv = timeit(inverse)
    ms_x = timeit(exchange)
    # round trip on this rank's rings (the field is not band-limited: compare inverse(forward(x)) twice instead)
    forward()
    inverse()
    a1_r = out_r.clone()
    my_r.copy_(out_r)
    my_i.copy_(out_i)
    forward()
    inverse()
    proj_err = float((out_r - a1_r).abs().max() / a1_r.abs().max())  # projection is idempotent
    res = {"bw": bw, "n_gpus": world, "ms_forward": ms_fwd, "ms_inverse": ms_inv, "ms_exchange_only": ms_x,
           "table_bytes_per_gpu": P.table_bytes(),
           "table_stream_gbs_per_gpu_fwd": P.table_stream_bytes() / (ms_fwd * 1e-3) / 1e9,
           "exchange_bytes_per_gpu": 8 * blk * (world - 1), "idempotence_rel_err": proj_err}
    if a.check:
        # gather the sharded coefficients and compare with the unsharded plan on rank 0
        my_r.copy_(full_r[rank * nr:(rank + 1) * nr])
        my_i.copy_(full_i[rank * nr:(rank + 1) * nr])
        cr.zero_()
        ci.zero_()
        forward()
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(cr)  # owned positions are disjoint, the rest is zero
            dist.all_reduce(ci)
        if rank == 0:
            Q = s2.Plan(bw, s2.MEMO, max_batch=1, device=local)
            Q.set_stream(torch.cuda.current_stream().cuda_stream)  # same stream as the fills below
            wr, wi = torch.zeros_like(cr), torch.zeros_like(ci)
            Q.fst(full_r, full_i, wr, wi, s2.COMPLEX)
            torch.cuda.synchronize()
            ok = torch.isfinite(wr) & torch.isfinite(wi)  # bw = 2048: orders >= 2044 are unpinned (reference NaN)
            scale = float(torch.maximum(wr[ok].abs().max(), wi[ok].abs().max()))
            res["sharded_vs_single_gpu_rel_err"] = float(
                torch.maximum((cr - wr)[ok].abs().max(), (ci - wi)[ok].abs().max())) / scale
            if res["sharded_vs_single_gpu_rel_err"] > 1e-9:  # diagnostics: which rank's orders are off
                import numpy as np
                err = torch.maximum((cr - wr).abs(), (ci - wi).abs()).cpu().numpy() / scale
                for r in range(world):
                    orders, _ = s2.shard_layout(bw, world, r)
                    worst = []
                    for m in orders:
                        for sm in ((m,) if m == 0 else (m, -m)):
                            a0 = s2.index_of_harmonic_coeff(sm, m, bw)
                            worst.append((float(np.nanmax(err[a0:a0 + bw - m])), sm))
                    worst.sort(reverse=True)
                    res[f"debug_rank{r}_worst"] = worst[:4]
                    res[f"debug_rank{r}_median"] = float(np.median([w[0] for w in worst]))
            Q.close()
    if rank == 0:
        print(json.dumps(res))
    P.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
