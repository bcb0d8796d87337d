"""Host-side sharding logic on CPU: the order / ring partition of the single-field path (s2kit_cuda_shard_layout)
and the ring<->order all-to-all pattern, exercised with world_size-2 gloo processes; plus the batched path's
function split used by bench.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_properties():
    import s2kit_b200 as s2

    for bw, g in ((64, 2), (64, 4), (256, 8), (2048, 8), (1024, 4)):
        n = 2 * bw
        seen_orders, seen_rows, work = [], [], []
        for r in range(g):
            orders, rows = s2.shard_layout(bw, g, r)
            assert len(orders) == bw // g and len(rows) == n // g
            seen_orders += orders
            seen_rows += [x for x in rows if x >= 0]
            work.append(sum(bw * bw - m * m for m in orders))  # contraction work ~ bw^2 - m^2
            for m in orders:  # both spectral rows of an owned order live on the owner
                assert m in rows and (m == 0 or n - m in rows)
        assert sorted(seen_orders) == list(range(bw))
        assert sorted(seen_rows) == [x for x in range(n) if x != bw]  # Nyquist row belongs to nobody
        assert max(work) / min(work) < (1.005 if bw >= 1024 else 1.05)  # pairs (m, bw-1-m) dealt round-robin balance the ranks
    with pytest.raises(s2.S2kitCudaError):
        s2.shard_layout(64, 3, 0)


def _worker(rank, world, port, bw, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import s2kit_b200 as s2

    n, nr = 2 * bw, 2 * bw // world
    # every rank fills its send blocks [dst][part][local row of dst][local ring] with a tag of (row, ring, part)
    send = torch.zeros(world, 2, nr, nr, dtype=torch.float64)
    for dst in range(world):
        _, rows = s2.shard_layout(bw, world, dst)
        for i, row in enumerate(rows):
            if row < 0:
                continue
            j = torch.arange(rank * nr, (rank + 1) * nr, dtype=torch.float64)
            send[dst, 0, i] = row * 10000 + j
            send[dst, 1, i] = -(row * 10000 + j)
    recv = torch.zeros_like(send)
    outs = list(recv.unbind(0))
    dist.all_to_all(outs, list(send.unbind(0))) if dist.get_backend() != "gloo" else None
    if dist.get_backend() == "gloo":  # gloo has no all_to_all: pairwise exchange
        for peer in range(world):
            if peer == rank:
                recv[peer] = send[peer]
            else:
                ops = [dist.P2POp(dist.isend, send[peer].contiguous(), peer), dist.P2POp(dist.irecv, outs[peer], peer)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
    # after the exchange rank `rank` holds, for each of ITS rows, all 2bw latitudes (segment s from peer s)
    _, rows = s2.shard_layout(bw, world, rank)
    ok = True
    for i, row in enumerate(rows):
        if row < 0:
            continue
        full = torch.cat([recv[s, 0, i] for s in range(world)])
        ok &= bool(torch.equal(full, row * 10000 + torch.arange(n, dtype=torch.float64)))
        ok &= bool(torch.equal(torch.cat([recv[s, 1, i] for s in range(world)]), -full))
    # batched path: bench.py gives every rank its own functions and only reduces the timing (MAX)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok &= float(t) == world
    ret[rank] = ok
    dist.destroy_process_group()


def test_ring_order_exchange_world2_gloo():
    world, bw = 2, 32
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, bw, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
