#!/usr/bin/env python
"""bench.py -- FP64 forward+inverse spherical harmonic transforms per second on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (InvFST then FST, COMPLEX format, Memo tables) over one batch of
synthetic band-limited functions.  Default workload = BASELINE.json configs[2]: bandwidth 256, 1024 functions
per GPU (one process per GPU, functions are independent: no data-path collective, weak scaling).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same metric through the
C-ABI with pinned HOST buffers (H2D/D2H inside the timed region), `roofline` = the dominant kernel against
its bound, `cpu_baseline` = the reference CPU path on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64_fwd_inv_sht_pairs_per_sec"
UNIT = "transform pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--bw", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1024, help="functions per GPU per step")
    ap.add_argument("--chunk", type=int, default=1024, help="functions per internal launch group (256: 8.2 ms per step, 512: 8.0, 1024: 7.9)")
    ap.add_argument("--format", default="complex", choices=["complex", "real"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="functions in the CPU sample (0 = auto)")
    ap.add_argument("--single-bw", type=int, default=2048, help="bandwidth of the single-field block (0 = skip)")
    ap.add_argument("--single-steps", type=int, default=10)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling pass (batch / n_gpus functions per GPU)")
    return ap.parse_args()


def workload_name(a):
    return f"batched InvFSTSemiMemo+FSTSemiMemo, bw={a.bw}, {a.batch} synthetic band-limited functions per GPU, {a.format.upper()} format"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pairs_per_sec(bw, fmt, sample, threads):
    """The reference's CPU path (oracle/_ref = its unmodified sources + FFTW-API stub; falls back to the port)."""
    import oracle

    kind = oracle.best_kind()
    if kind == "ref":
        O = oracle.Oracle(bw, "ref")
        O.bench_pairs(min(threads, sample), min(threads, sample), 999, fmt)  # warm caches / page in
        wall, busy = O.bench_pairs(sample, threads, 1000, fmt)
        O.close()
        return sample / wall, "reference", threads
    # port: single thread, python-driven
    O = oracle.Oracle(bw, "port")
    rc, ic = O.gen_coeffs(1000)
    n = max(2, min(sample, 8))
    t0 = time.time()
    for _ in range(n):
        g = O.inverse(rc, ic, fmt)
        O.forward(g[0], g[1], fmt)
    wall = time.time() - t0
    return n / wall, "port", 1


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fmt = 0 if a.format == "complex" else 1
    threads = os.cpu_count() or 1
    sample = a.cpu_sample or max(threads, 64)
    vals = []
    kind, cores = "reference", threads
    for i in range(a.warmup + a.steps):
        v, kind, cores = cpu_pairs_per_sec(a.bw, fmt, sample, threads)
        if i >= a.warmup:
            vals.append(v)
        if i == 0 and sample / v > 20:  # keep the whole run within minutes
            sample = max(cores, int(v * 10))
    value = len(vals) / sum(1.0 / v for v in vals)
    sample_txt = (f"{sample} functions (seeds 1000..) per step on {cores} host threads, InvFSTSemiMemo+FSTSemiMemo, "
                  f"tables excluded; FFTW replaced by oracle/fftw_stub (FFTW not installed)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "bw": a.bw, "format": a.format},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; start() before the warm-up so the tool is up by the
    time the timed region begins, mark() at its two ends; samples inside the marks are the ones reported."""

    def __init__(self, dev):
        self.dev, self.rows, self.proc, self.t0, self.t1 = dev, [], None, None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in ln.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.05]
        rows = inside or [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm),
                "samples_in_timed_region": len(inside)}


# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per function and kernel from the committed
# `ncu --set full` capture of this command at bw = 256, COMPLEX, 256 functions per launch
# (profiles/r1_ncu_full_metrics_final.csv); scaled to the launch size for `roofline.traffic`.
NCU_DRAM_BYTES_PER_FUNCTION_BW256_COMPLEX = {
    "phi_fft_fwd": (1.074085e9 + 1.027033e9) / 256, "phi_fft_inv": (1.07166e9 + 1.028037e9) / 256,
    "dct_fwd": (1.071721e9 + 0.508014e9) / 256, "dct_inv": (0.545191e9 + 1.017913e9) / 256,
    "legendre_fwd": (0.570337e9 + 0.271074e9) / 256, "legendre_inv": (0.294773e9 + 0.484361e9) / 256,
    # persistent K2+K3 kernel (kernels_pipe.cu), profiles/r1_ncu_pipe_summary.md
    "fused_fwd": (1.099542e9 + 0.270152e9) / 256,
}


def measured_traffic(kind, bw, fmt_real, functions_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of kernel kind `kind`, from the committed
    `ncu --set full` capture of THIS command at the SAME launch size (profiles/r2_dram_traffic.json, written by
    tools/ncu_traffic.py); falls back to the round-1 capture at 256 functions per launch, scaled."""
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if os.path.exists(path):
        try:
            rec = json.load(open(path))
            if rec.get("bw") == bw and bool(rec.get("format_real")) == bool(fmt_real):
                ent = rec.get("kernels", {}).get(kind)
                if ent and abs(ent["functions_per_launch"] - functions_per_launch) < 0.5:
                    return ent["dram_bytes_per_launch"], "profiles/r2_dram_traffic.json (ncu --set full, same launch size)"
        except Exception:
            pass
    if bw == 256 and not fmt_real and kind in NCU_DRAM_BYTES_PER_FUNCTION_BW256_COMPLEX:
        return (NCU_DRAM_BYTES_PER_FUNCTION_BW256_COMPLEX[kind] * functions_per_launch,
                "profiles/r1_ncu_full_metrics_final.csv (256 functions per launch, scaled)")
    return None, None


def table_doubles(bw):
    tot = 0
    for m in range(bw):
        tot += sum((l - 1) // 2 + 1 if m % 2 else l // 2 + 1 for l in range(m, bw))
    return tot


def algorithmic_per_function(bw, fmt_real):
    """algorithmic bytes / flops per function and stage."""
    B2 = bw * bw
    S = table_doubles(bw)
    T0 = sum(l // 2 + 1 for l in range(bw))
    lg = math.log2(2 * bw)
    return {
        # REAL format: both grid arrays are still moved (64 B^2), but only the order rows m' < bw of the spectral planes
        # (32 B^2 instead of 64 B^2), and the contraction reads half of the cosine planes
        "phi_fft": {"bytes": (96 if fmt_real else 128) * B2, "flops": 20 * B2 * lg},
        "dct": {"bytes": (48 if fmt_real else 96) * B2, "flops": (10 if fmt_real else 20) * B2 * lg},
        "legendre": {"bytes": (32 if fmt_real else 48) * B2, "table_bytes": 8 * S,
                     "flops": (4 * S) if fmt_real else (8 * S - 4 * T0)},
    }


def synth_coeffs(torch, bw, batch, device, seed):
    """Random coefficients of real-valued band-limited fields, symmetry of test_s2_semi_memo.c:156-172."""
    import numpy as np

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rc = torch.rand(batch, bw * bw, generator=g, device=device, dtype=torch.float64) * 2 - 1
    ic = torch.rand(batch, bw * bw, generator=g, device=device, dtype=torch.float64) * 2 - 1
    pos, neg, sgn = [], [], []
    for m in range(1, bw):
        l = np.arange(m, bw)
        pos.append(m * bw - (m * (m - 1)) // 2 + (l - m))
        big = bw - 1
        neg.append((big * (big + 3)) // 2 + 1 + ((big - m) * (big - m + 1)) // 2 + (l - m))
        sgn.append(np.full(l.shape, -1.0 if m % 2 else 1.0))
    pos = torch.from_numpy(np.concatenate(pos)).to(device)
    neg = torch.from_numpy(np.concatenate(neg)).to(device)
    sgn = torch.from_numpy(np.concatenate(sgn)).to(device)
    rc[:, neg] = rc[:, pos] * sgn
    ic[:, neg] = -ic[:, pos] * sgn
    ic[:, :bw] = 0.0
    return rc, ic


def _event_ms(torch, fn, warmup, steps, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def single_field_block(a, torch, dist, s2, world, rank, local, dev, hbm_peak, cpu_group=None):
    """The second half of BASELINE.json's metric: ONE field at bw = 2048 (configs[4], FSTSemiMemo with m-sharded tables).
    N = 1: the whole 11.7 GB table streams through one GPU per transform.  N > 1 (torchrun): (a) every rank runs its
    share, ring -> order exchange by NCCL all_to_all; (b) rank 0 alone drives all N GPUs through the in-library
    s2kit_cuda_multi_* path (exchange inside the DCT kernels over peer memory).  Forward results are checked against
    the committed sample of the reference's own output (tests/golden/oracle_vectors_large.npz)."""
    import numpy as np

    bw = a.single_bw
    n = 2 * bw
    steps = a.single_steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    golden = None
    gpath = os.path.join(ROOT, "tests", "golden", "oracle_vectors_large.npz")
    if bw == 2048 and os.path.exists(gpath):
        golden = np.load(gpath)
    rng = np.random.RandomState(2048)  # the grid the golden sample was computed from
    rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))

    def sample_err(fr, fi):
        if golden is None:
            return None
        wr, wi = golden["bw2048_fwd_sample_r"], golden["bw2048_fwd_sample_i"]
        ok = np.isfinite(wr) & np.isfinite(wi)  # the reference is NaN for |m| >= 2044 (pmm.c:22-30)
        scale = max(np.abs(wr[ok]).max(), np.abs(wi[ok]).max())
        return float(max(np.abs(fr[::509][ok] - wr[ok]).max(), np.abs(fi[::509][ok] - wi[ok]).max()) / scale)

    out = {"workload": f"single field, FSTSemiMemo + InvFSTSemiMemo at bw={bw}, COMPLEX format, Memo tables, "
                       f"{'sharded by order over ' + str(world) + ' GPUs' if world > 1 else 'one GPU'}",
           "bw": bw, "n_gpus": world, "steps": steps,
           "note": "inverse parity at bw = 2048 is pinned per order (tests/test_gpu_large.py); the reference's own 2-D "
                   "inverse is NaN at this size and its tables lose orders ~550-950 to seed underflow (DESIGN.md section 4)"}
    if world == 1:
        P = s2.Plan(bw, s2.MEMO, max_batch=1, device=local)
        P.set_stream(torch.cuda.current_stream().cuda_stream)
        gr, gi = torch.tensor(rd, device=dev), torch.tensor(idt, device=dev)
        cr = torch.zeros(bw * bw, device=dev, dtype=torch.float64)
        ci = torch.zeros_like(cr)
        og_r, og_i = torch.empty_like(gr), torch.empty_like(gi)
        ms_f = _event_ms(torch, lambda: P.fst(gr, gi, cr, ci, s2.COMPLEX), 3, steps, barrier)
        ms_i = _event_ms(torch, lambda: P.inv_fst(cr, ci, og_r, og_i, s2.COMPLEX), 3, steps, barrier)
        P.profile(True)
        P.fst(gr, gi, cr, ci, s2.COMPLEX)
        P.inv_fst(cr, ci, og_r, og_i, s2.COMPLEX)
        prof = {k: v[0] for k, v in P.profile_get().items() if v[1]}
        P.profile(False)
        one_copy = P.table_stream_bytes()  # what one transform reads; the plan holds P.table_bytes()
        out.update({"ms_forward": ms_f, "ms_inverse": ms_i, "pairs_per_s": 1e3 / (ms_f + ms_i),
                    "table_bytes_streamed_per_transform": one_copy, "table_bytes_resident": P.table_bytes(),
                    "kernel_ms_one_pair": prof,
                    "table_stream_gbs_forward": one_copy / (ms_f * 1e-3) / 1e9,
                    "table_stream_gbs_inverse": one_copy / (ms_i * 1e-3) / 1e9,
                    "frac_of_hbm_peak_forward": one_copy / (ms_f * 1e-3) / 1e9 / hbm_peak,
                    "frac_of_hbm_peak_inverse": one_copy / (ms_i * 1e-3) / 1e9 / hbm_peak,
                    "legendre_fwd_table_stream_frac": (one_copy / (prof["legendre_fwd"] * 1e-3) / 1e9 / hbm_peak
                                                       if "legendre_fwd" in prof else None),
                    "legendre_inv_table_stream_frac": (one_copy / (prof["legendre_inv"] * 1e-3) / 1e9 / hbm_peak
                                                       if "legendre_inv" in prof else None),
                    "forward_rel_err_vs_reference_sample": sample_err(cr.cpu().numpy(), ci.cpu().numpy())})
        P.close()
        return out
    # ---- (a) one process per GPU, NCCL all_to_all between the two halves of the transform
    P = s2.ShardedPlan(bw, rank, world, device=local)
    P.set_stream(torch.cuda.current_stream().cuda_stream)
    nr, blk = P.rings, P.block_doubles
    my_r = torch.tensor(rd[rank * nr:(rank + 1) * nr], device=dev)
    my_i = torch.tensor(idt[rank * nr:(rank + 1) * nr], device=dev)
    send = torch.zeros(world * blk, device=dev, dtype=torch.float64)
    recv = torch.zeros_like(send)
    cr = torch.zeros(bw * bw, device=dev, dtype=torch.float64)
    ci = torch.zeros_like(cr)
    out_r, out_i = torch.zeros_like(my_r), torch.zeros_like(my_i)

    def forward():
        P.fst_rings(my_r, my_i, send)
        dist.all_to_all_single(recv, send)
        P.fst_orders(recv, cr, ci)

    def inverse():
        P.inv_fst_orders(cr, ci, send)
        dist.all_to_all_single(recv, send)
        P.inv_fst_rings(recv, out_r, out_i)

    ms_f = allmax(_event_ms(torch, forward, 3, steps, barrier))
    ms_i = allmax(_event_ms(torch, inverse, 3, steps, barrier))
    ms_x = allmax(_event_ms(torch, lambda: dist.all_to_all_single(recv, send), 3, steps, barrier))
    forward()
    torch.cuda.synchronize()
    full_r, full_i = cr.clone(), ci.clone()
    dist.all_reduce(full_r)  # owned positions are disjoint, everything else is zero
    dist.all_reduce(full_i)
    x_bytes = 8 * blk * (world - 1)
    nccl = {"ms_forward": ms_f, "ms_inverse": ms_i, "pairs_per_s": 1e3 / (ms_f + ms_i), "ms_exchange_only": ms_x,
            "exchange_share_of_forward": ms_x / ms_f, "exchange_bytes_per_gpu": x_bytes,
            "exchange_gbs_per_gpu": x_bytes / (ms_x * 1e-3) / 1e9, "table_bytes_per_gpu": P.table_bytes(),
            "table_stream_gbs_per_gpu_forward": P.table_stream_bytes() / (ms_f * 1e-3) / 1e9,
            "forward_rel_err_vs_reference_sample": sample_err(full_r.cpu().numpy(), full_i.cpu().numpy())}
    out["nccl_all_to_all"] = nccl
    P.close()
    del send, recv, my_r, my_i, out_r, out_i
    torch.cuda.empty_cache()
    barrier()
    # ---- (b) rank 0 drives all GPUs through the C-ABI (s2kit_cuda_multi_*): no NCCL, the DCT kernels do the exchange.
    # The other ranks must leave their GPUs idle meanwhile: they wait on a CPU (gloo) barrier -- an NCCL barrier would keep
    # a spinning kernel of another process on every GPU, and the time-slicing between the two contexts doubled the measured
    # time of this path (2.67 vs 1.22 ms forward on two GPUs, profiles/r2_multi_stage_times.txt)
    if rank == 0:
        try:
            M = s2.MultiPlan(bw, world)
            got_r, got_i = M.forward(rd, idt)  # host-pointer call: also loads the device-resident buffers
            ms_mf = M.run(inverse=False, iters=steps)
            ms_mf = M.run(inverse=False, iters=steps)
            ms_mi = M.run(inverse=True, iters=steps)
            ms_mi = M.run(inverse=True, iters=steps)
            t0 = time.perf_counter()
            M.forward(rd, idt)
            wall = time.perf_counter() - t0
            fr, fi = full_r.cpu().numpy(), full_i.cpu().numpy()
            ok = np.isfinite(fr) & np.isfinite(got_r)
            out["in_library_p2p"] = {
                "ms_forward": ms_mf, "ms_inverse": ms_mi, "pairs_per_s": 1e3 / (ms_mf + ms_mi),
                "table_bytes_per_gpu": M.table_bytes_per_gpu(),
                "table_stream_gbs_per_gpu_forward": M.table_bytes_per_gpu() / (ms_mf * 1e-3) / 1e9,
                "host_pointer_forward_wall_ms": wall * 1e3,
                "forward_rel_err_vs_reference_sample": sample_err(got_r, got_i),
                "rel_err_vs_nccl_path": float(max(np.abs(got_r - fr)[ok].max(), np.abs(got_i - fi)[ok].max()) /
                                              np.abs(fr[ok]).max()),
                "note": "one process, s2kit_cuda_multi_run: device-resident rings/coefficients, CUDA-event time per "
                        "transform, maximum over the GPUs; ring<->order exchange = NVLink loads/stores inside K2/K5"}
            M.close()
        except Exception as ex:  # noqa: BLE001 -- e.g. no peer access on this box: report, do not lose the line
            out["in_library_p2p"] = {"unavailable": repr(ex)[:300]}
    if cpu_group is not None:
        dist.barrier(group=cpu_group)
    barrier()
    best = out["nccl_all_to_all"]
    if rank == 0 and "ms_forward" in out.get("in_library_p2p", {}):
        if out["in_library_p2p"]["pairs_per_s"] > best["pairs_per_s"]:
            best = out["in_library_p2p"]
    out.update({"ms_forward": best["ms_forward"], "ms_inverse": best["ms_inverse"], "pairs_per_s": best["pairs_per_s"]})
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist

    import s2kit_b200 as s2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None  # host-side barrier (single_field_block)
    fmt = s2.COMPLEX if a.format == "complex" else s2.REAL
    bw, n, batch = a.bw, 2 * a.bw, a.batch

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    fp64 = s2.measure_fp64_peak(local)  # FP64 DFMA / DMMA peaks are not in MEASURED_PEAKS.json: measured here

    plan = s2.Plan(bw, s2.MEMO, max_batch=a.chunk, device=local)
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    rc, ic = synth_coeffs(torch, bw, batch, dev, 1000 + rank)
    rd = torch.empty(batch, n, n, device=dev, dtype=torch.float64)
    idt = torch.empty_like(rd)
    rc2, ic2 = torch.empty_like(rc), torch.empty_like(ic)

    def step():
        plan.inv_fst(rc, ic, rd, idt, fmt)
        plan.fst(rd, idt, rc2, ic2, fmt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    err = float(((rc2 - rc).abs().max().item() + (ic2 - ic).abs().max().item()))  # round-trip sanity

    def timed(profile):
        plan.profile(profile)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # the timed region proper: K steps, no per-kernel instrumentation
    sampler.mark_begin()
    ms = timed(False)
    sampler.mark_end()
    clocks = sampler.stop()
    # the same K steps again with a CUDA-event pair around every kernel launch (per-stage roofline numbers)
    ms_profiled = timed(True)
    prof = plan.profile_get()
    plan.profile(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * batch * a.steps / (ms * 1e-3)

    # ---- strong scaling: BASELINE configs[2] as written -- the SAME 1024 functions sharded over the N GPUs
    strong = None
    if not a.no_strong:
        sb = max(1, batch // world)

        def sstep():
            plan.inv_fst(rc[:sb], ic[:sb], rd[:sb], idt[:sb], fmt)
            plan.fst(rd[:sb], idt[:sb], rc2[:sb], ic2[:sb], fmt)

        if world == 1:
            strong = {"total_functions": sb, "functions_per_gpu": sb, "value": value, "unit": UNIT,
                      "ms_per_step": ms / a.steps, "note": "N = 1: identical to the weak-scaling run above"}
        else:
            for _ in range(3):
                sstep()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                sstep()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms = float(t.item())
            strong = {"total_functions": sb * world, "functions_per_gpu": sb, "value": sb * world * a.steps / (sms * 1e-3),
                      "unit": UNIT, "ms_per_step": sms / a.steps,
                      "note": "same total work as the 1-GPU run, independent functions per rank, no collective"}

    # ---- end to end through the C-ABI with pinned host buffers
    e2e = None
    if not a.no_e2e:
        hc_r = torch.empty(batch, bw * bw, dtype=torch.float64).pin_memory()
        hc_i = torch.empty_like(hc_r).pin_memory()
        hc_r.copy_(rc.cpu())
        hc_i.copy_(ic.cpu())
        hg_r = torch.empty(batch, n, n, dtype=torch.float64).pin_memory()
        hg_i = torch.empty(batch, n, n, dtype=torch.float64).pin_memory()
        ho_r = torch.empty(batch, bw * bw, dtype=torch.float64).pin_memory()
        ho_i = torch.empty(batch, bw * bw, dtype=torch.float64).pin_memory()

        def e2e_step():
            plan.inv_fst(hc_r, hc_i, hg_r, hg_i, fmt)   # H2D coefficients, D2H grids
            plan.fst(hg_r, hg_i, ho_r, ho_i, fmt)       # H2D grids, D2H coefficients

        # One host thread, the two calls back to back.  (Measured: running the inverse of one half of the batch and the
        # forward of the other half concurrently from two threads / two plans, so that both PCIe directions are busy at
        # once, gives the same throughput -- 6318 vs 6303 pairs/s: the box moves ~66 GB/s host<->device in total,
        # whatever the mix of directions.  profiles/r1_ncu_summary.md)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_err = float((ho_r - hc_r).abs().max().item())
        per = 8 * (2 * bw * bw + 2 * n * n)
        e2e = {"value": world * batch * a.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": per * batch,
               "d2h_bytes_per_step": per * batch, "steps": a.e2e_steps, "roundtrip_max_abs_err": e2e_err,
               "note": "s2kit_cuda_inv_fst + s2kit_cuda_fst with pinned HOST buffers, copies inside the timed region"}

    # ---- roofline per kernel kind; the dominant one (largest share of the step) is reported as `roofline`
    alg = algorithmic_per_function(bw, fmt == s2.REAL)
    B2 = bw * bw
    per_fn = {  # algorithmic (bytes, flops, table bytes per launch) -- SURVEY.md section 8(d)
        "phi_fft_fwd": (alg["phi_fft"]["bytes"], alg["phi_fft"]["flops"], 0),
        "phi_fft_inv": (alg["phi_fft"]["bytes"], alg["phi_fft"]["flops"], 0),
        "dct_fwd": (alg["dct"]["bytes"], alg["dct"]["flops"], 0),
        "dct_inv": (alg["dct"]["bytes"], alg["dct"]["flops"], 0),
        "legendre_fwd": (alg["legendre"]["bytes"], alg["legendre"]["flops"], alg["legendre"]["table_bytes"]),
        "legendre_inv": (alg["legendre"]["bytes"], alg["legendre"]["flops"], alg["legendre"]["table_bytes"]),
        # fused DCT+Legendre: reads the spectral rows (2/3 of the DCT stage's bytes), writes coefficients
        "fused_fwd": (alg["dct"]["bytes"] * 2 // 3 + 16 * B2, alg["legendre"]["flops"] + alg["dct"]["flops"],
                      alg["legendre"]["table_bytes"]),
        "fused_inv": (alg["dct"]["bytes"] * 2 // 3 + 16 * B2, alg["legendre"]["flops"] + alg["dct"]["flops"],
                      alg["legendre"]["table_bytes"]),
    }
    stages = {}
    for kind, (kms, cnt) in prof.items():
        if cnt == 0 or kind not in per_fn:
            continue
        fn_per_launch = batch * a.steps / cnt
        avg_s = kms * 1e-3 / cnt
        byts = per_fn[kind][0] * fn_per_launch + per_fn[kind][2]
        flops = per_fn[kind][1] * fn_per_launch
        t_hbm, t_fp = byts / (hbm_peak * 1e9), flops / (fp64["dmma_tflops"] * 1e12)
        ent = {"ms_total": kms, "launches": cnt, "avg_launch_ms": kms / cnt, "functions_per_launch": fn_per_launch,
               "alg_bytes_per_launch": byts, "alg_flops_per_launch": flops,
               "hbm_gbs": byts / avg_s / 1e9, "fp64_tflops": flops / avg_s / 1e12}
        if t_fp > t_hbm:
            ent.update({"bound": "tensor", "achieved": ent["fp64_tflops"], "peak": fp64["dmma_tflops"], "unit": "TFLOP/s"})
        else:
            ent.update({"bound": "hbm", "achieved": ent["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s"})
        ent["frac"] = ent["achieved"] / ent["peak"]
        stages[kind] = ent
    dom = max(stages, key=lambda k: stages[k]["ms_total"]) if stages else None
    roofline = None
    if dom:
        d = stages[dom]
        traffic, traffic_src = measured_traffic(dom, bw, fmt == s2.REAL, d["functions_per_launch"])
        roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                    "frac": d["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "alg_bytes_per_launch": d["alg_bytes_per_launch"], "alg_flops_per_launch": d["alg_flops_per_launch"],
                    "peak_source": (hbm_src if d["bound"] == "hbm" else
                                    "FP64 tensor (DMMA mma.sync.m8n8k4.f64) micro-benchmark measured in this run; "
                                    "MEASURED_PEAKS.json has no FP64 figure"),
                    "share_of_step": d["ms_total"] / ms_profiled}
    launches = int(sum(c for _, c in prof.values()))

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        threads = os.cpu_count() or 1
        sample = a.cpu_sample or max(threads, 64)
        v, kind, cores = cpu_pairs_per_sec(bw, 0 if fmt == s2.COMPLEX else 1, sample, threads)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{sample} functions on {cores} host threads, InvFSTSemiMemo+FSTSemiMemo of the reference "
                         f"(FFTW replaced by oracle/fftw_stub), tables excluded"}

    single = None
    if a.single_bw:
        plan.close()
        del rd, idt, rc2, ic2
        torch.cuda.empty_cache()
        try:
            single = single_field_block(a, torch, dist, s2, world, rank, local, dev, hbm_peak, cpu_group)
        except Exception as ex:  # noqa: BLE001 -- never lose the headline line to the second block
            single = {"error": repr(ex)[:400]}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "ms_per_step_with_kernel_events": ms_profiled / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "bw": bw, "functions_per_gpu": batch, "chunk": a.chunk,
                       "format": a.format, "variant": "memo",
                       "l2": "inputs larger than L2 (per step 1 GiB coefficients -> 4 GiB grids -> 1 GiB coefficients)",
                       "sharding": "independent functions per rank, no collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "stages": stages,
            "fp64_peak_measured": fp64, "cpu_baseline": cpu, "roundtrip_max_abs_err": err,
            "strong_scaling": strong, "single_field": single,
        }
        print(json.dumps(out))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
