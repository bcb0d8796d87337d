// multi.cu -- ONE spherical harmonic transform of a single large-bandwidth field on the G GPUs of one box, driven by
// one process through the C-ABI (s2kit_cuda_multi_*, include/s2kit_cuda.h).
//
// What it replaces: the serial m-loop of FSTSemiMemo / InvFSTSemiMemo (src/FST_semi_memo.c:96-201, 262-342) at
// bandwidths whose tables (11.7 GB at bw = 2048) are streamed once per transform -- the stream is divided over the
// GPUs.  Partition as in shard.cu: latitude rings for the longitude FFT (K1 / K6), orders (m paired with bw-1-m) with
// their tables for the DCT + Legendre stages (K2-K5).
//
// The ring <-> order exchange is not a separate collective: every GPU maps its peers' ring buffers (peer access over
// NVLink / NVSwitch) and the DCT kernels do the exchange as part of their own memory accesses --
//   forward : K2 on the order owner LOADS the 2bw/G-latitude run of each of its rows from the ring owner's K1 output
//             (contiguous 4 KiB runs at bw = 2048, G = 8), so the transfer of one row overlaps the transforms of others;
//   inverse : K5 STORES each run of its result rows into the ring owner's buffer, K6 then reads local memory.
// Ordering between GPUs: one event per device and direction (K1 done / K5 done) that every peer's stream waits on;
// the ring buffers are double-buffered so back-to-back transforms need no further synchronisation (see the comments
// in run_forward / run_inverse).  One host thread per GPU enqueues its device's work, so launch overhead does not
// serialise over the devices.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "s2k_shard.cuh"

int s2k_fail_msg(const char* what);
int s2k_fail_cuda(const char* what, cudaError_t e);

namespace {

// sense-reversing spin barrier for the G enqueue threads (they meet within microseconds of each other)
struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 1;
    void wait() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_release);
        } else {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) == s)
                if (++spins > 2000) std::this_thread::yield();
        }
    }
};

// one owned run of coefficients: `len` doubles at `at` of the full bw*bw arrays, `packed` in the compact buffer
struct Run {
    long at, packed;
    int len;
};

__global__ void k_runs_copy(const Run* __restrict__ runs, const double* __restrict__ src_r,
                            const double* __restrict__ src_i, double* __restrict__ dst_r, double* __restrict__ dst_i,
                            int pack) {
    const Run r = runs[blockIdx.x];
    for (int i = threadIdx.x; i < r.len; i += blockDim.x) {
        const long a = pack ? r.at + i : r.packed + i, b = pack ? r.packed + i : r.at + i;
        dst_r[b] = src_r[a];
        dst_i[b] = src_i[a];
    }
}

enum JobKind { JOB_NONE = 0, JOB_FWD, JOB_INV, JOB_EXIT };

struct Job {
    int kind = JOB_NONE;
    int iters = 1;
    bool host_io = false;
    const double *h_in_r = nullptr, *h_in_i = nullptr;  // host arrays of a host-pointer call
    double *h_out_r = nullptr, *h_out_i = nullptr;
};

}  // namespace

struct s2kit_cuda_multi {
    int bw = 0, n = 0, G = 0, nr = 0;
    long block = 0;
    struct Dev {
        int device = 0;
        s2kit_cuda_plan* plan = nullptr;
        double *ring_r = nullptr, *ring_i = nullptr;  // this GPU's latitude rings [nr][2bw]
        double *coef_r = nullptr, *coef_i = nullptr;  // full bw*bw arrays, owned positions valid
        double* ringbuf[2] = {nullptr, nullptr};      // exchange blocks [peer][part][local row][local ring], two generations
        double *pack_r = nullptr, *pack_i = nullptr;  // owned coefficients, compact
        double *hpack_r = nullptr, *hpack_i = nullptr;  // pinned host mirrors
        Run* d_runs = nullptr;
        std::vector<Run> runs;
        long packed_len = 0;
        cudaEvent_t ev_x[2] = {nullptr, nullptr};  // "my exchange-side kernel of generation g is done"
        cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
        std::thread th;
        int status = 0;
        std::string error;
        float ms = 0.f;
    };
    std::vector<Dev> dev;
    // job hand-off
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    Job job;
    long job_seq = 0;
    int done = 0;
    SpinBarrier bar;
    unsigned gen = 0;  // ring-buffer generation, advances once per transform
    std::mutex api_mu;  // one call at a time
    double last_ms = 0.0;
};

namespace {

#define MCK(d, call)                                                      \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            (d).status = 1;                                               \
            (d).error = std::string(#call) + ": " + cudaGetErrorString(e__); \
        }                                                                 \
    } while (0)

// S2KIT_CUDA_MULTI_PROF=1: CUDA events between the stages of the LAST transform of a job, printed per device (diagnostics)
static bool multi_prof() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_MULTI_PROF");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}
struct StageProf {
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    void mark(cudaStream_t s) {
        if (n >= 6) return;
        if (!ev[n]) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n++], s);
    }
    void report(int g, const char* const* names) {
        char line[256];
        int at = snprintf(line, sizeof line, "[s2kit multi] gpu %d:", g);
        for (int i = 0; i + 1 < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            at += snprintf(line + at, sizeof line - at, " %s %.3f ms", names[i], ms);
        }
        fprintf(stderr, "%s\n", line);
        n = 0;
    }
};
static StageProf g_prof[s2k::S2K_MAX_PEERS];

// forward on device g, generation `gen` of the ring buffers.
// Safety of reusing ringbuf[gen & 1] without extra events: K1_s(i+2) follows K2_s(i+1) in s's stream, K2_s(i+1) waited
// for every K1_d(i+1), and K1_d(i+1) follows K2_d(i) -- the last reader of generation i on any device -- in d's stream.
void run_forward(s2kit_cuda_multi* mp, int g, unsigned gen, bool first, bool last, const Job& job) {
    s2kit_cuda_multi::Dev& d = mp->dev[g];
    s2kit_cuda_plan* p = d.plan;
    ShardState* st = shard_of(p);
    const int b = gen & 1, n = mp->n, nr = mp->nr, bw = mp->bw;
    cudaStream_t s = p->stream;
    if (first) {
        if (job.host_io) {
            MCK(d, cudaMemcpyAsync(d.ring_r, job.h_in_r + (long)g * nr * n, sizeof(double) * nr * n, cudaMemcpyHostToDevice, s));
            MCK(d, cudaMemcpyAsync(d.ring_i, job.h_in_i + (long)g * nr * n, sizeof(double) * nr * n, cudaMemcpyHostToDevice, s));
        }
        MCK(d, cudaEventRecord(d.ev_start, s));
    }
    const bool prof = last && multi_prof();
    if (prof) g_prof[g].mark(s);
    MCK(d, s2k::launch_phi_fft_fwd(p, d.ring_r, d.ring_i, 0, d.ringbuf[b], 1, S2KIT_COMPLEX, &st->ring_view));
    MCK(d, cudaEventRecord(d.ev_x[b], s));
    if (prof) g_prof[g].mark(s);
    mp->bar.wait();  // every device has recorded its K1 event of this generation
    for (int q = 0; q < mp->G; ++q)
        if (q != g) MCK(d, cudaStreamWaitEvent(s, mp->dev[q].ev_x[b], 0));
    if (prof) g_prof[g].mark(s);
    s2k::PlaneView ov = st->order_view;
    s2k::PeerSegs peers;
    memset(&peers, 0, sizeof(peers));
    for (int q = 0; q < mp->G; ++q) peers.ptr[q] = mp->dev[q].ringbuf[b] + (long)g * mp->block;
    ov.peers = &peers;
    // K2 pulls its rows out of the peers' K1 output: the exchange rides on the kernel's own loads
    MCK(d, s2k::launch_dct_fwd(p, d.ringbuf[b], p->d_X, 1, 0, st->nrows_real, S2KIT_COMPLEX, &ov));
    if (prof) g_prof[g].mark(s);
    MCK(d, s2k::launch_legendre_fwd(p, p->d_table, 0, p->d_X, d.coef_r, d.coef_i, (long)bw * bw, 1, 0, st->norders,
                                    S2KIT_COMPLEX, st->d_orders));
    if (prof) g_prof[g].mark(s);
    // (ev_x[b] is recorded again two generations later; by then every peer has passed the barrier of the generation in
    // between, i.e. has long enqueued its wait on this record)
    if (last) {
        MCK(d, cudaEventRecord(d.ev_stop, s));
        if (job.host_io) {
            k_runs_copy<<<(unsigned)d.runs.size(), 128, 0, s>>>(d.d_runs, d.coef_r, d.coef_i, d.pack_r, d.pack_i, 1);
            MCK(d, cudaGetLastError());
            MCK(d, cudaMemcpyAsync(d.hpack_r, d.pack_r, sizeof(double) * d.packed_len, cudaMemcpyDeviceToHost, s));
            MCK(d, cudaMemcpyAsync(d.hpack_i, d.pack_i, sizeof(double) * d.packed_len, cudaMemcpyDeviceToHost, s));
        }
    }
}

// inverse on device g.  K5 pushes its result runs into the ring owners' buffers; K6 reads local memory.
// Reuse of ringbuf[gen & 1]: K5_d(i+2) follows K6_d(i+1), which waited for K5_s(i+1) of every s, which follows
// K6_s(i) -- the last reader of generation i on device s.
void run_inverse(s2kit_cuda_multi* mp, int g, unsigned gen, bool first, bool last, const Job& job) {
    s2kit_cuda_multi::Dev& d = mp->dev[g];
    s2kit_cuda_plan* p = d.plan;
    ShardState* st = shard_of(p);
    const int b = gen & 1, n = mp->n, nr = mp->nr, bw = mp->bw;
    cudaStream_t s = p->stream;
    if (first) {
        if (job.host_io) {
            MCK(d, cudaMemcpyAsync(d.pack_r, d.hpack_r, sizeof(double) * d.packed_len, cudaMemcpyHostToDevice, s));
            MCK(d, cudaMemcpyAsync(d.pack_i, d.hpack_i, sizeof(double) * d.packed_len, cudaMemcpyHostToDevice, s));
            k_runs_copy<<<(unsigned)d.runs.size(), 128, 0, s>>>(d.d_runs, d.pack_r, d.pack_i, d.coef_r, d.coef_i, 0);
            MCK(d, cudaGetLastError());
        }
        MCK(d, cudaEventRecord(d.ev_start, s));
    }
    const bool prof = last && multi_prof();
    if (prof) g_prof[g].mark(s);
    MCK(d, s2k::launch_legendre_inv(p, p->d_table_t, 0, d.coef_r, d.coef_i, (long)bw * bw, p->d_X, 1, 0, st->norders,
                                    S2KIT_COMPLEX, st->d_orders));
    if (prof) g_prof[g].mark(s);
    s2k::PlaneView ov = st->order_view;
    s2k::PeerSegs peers;
    memset(&peers, 0, sizeof(peers));
    for (int q = 0; q < mp->G; ++q) peers.ptr[q] = mp->dev[q].ringbuf[b] + (long)g * mp->block;
    ov.peers = &peers;
    MCK(d, s2k::launch_dct_inv(p, p->d_X, d.ringbuf[b], 1, 0, st->nrows_real, S2KIT_COMPLEX, &ov));
    MCK(d, cudaEventRecord(d.ev_x[b], s));
    if (prof) g_prof[g].mark(s);
    mp->bar.wait();
    for (int q = 0; q < mp->G; ++q)
        if (q != g) MCK(d, cudaStreamWaitEvent(s, mp->dev[q].ev_x[b], 0));
    if (prof) g_prof[g].mark(s);
    MCK(d, s2k::launch_phi_fft_inv(p, d.ringbuf[b], d.ring_r, d.ring_i, 0, 1, S2KIT_COMPLEX, &st->ring_view));
    if (prof) g_prof[g].mark(s);
    if (last) {
        MCK(d, cudaEventRecord(d.ev_stop, s));
        if (job.host_io) {
            MCK(d, cudaMemcpyAsync(job.h_out_r + (long)g * nr * n, d.ring_r, sizeof(double) * nr * n, cudaMemcpyDeviceToHost, s));
            MCK(d, cudaMemcpyAsync(job.h_out_i + (long)g * nr * n, d.ring_i, sizeof(double) * nr * n, cudaMemcpyDeviceToHost, s));
        }
    }
}

void worker(s2kit_cuda_multi* mp, int g) {
    s2kit_cuda_multi::Dev& d = mp->dev[g];
    cudaSetDevice(d.device);
    long seen = 0;
    for (;;) {
        Job job;
        unsigned gen0;
        {
            std::unique_lock<std::mutex> lk(mp->mu);
            mp->cv_go.wait(lk, [&] { return mp->job_seq != seen; });
            seen = mp->job_seq;
            job = mp->job;
            gen0 = mp->gen;
        }
        if (job.kind == JOB_EXIT) return;
        d.status = 0;
        for (int it = 0; it < job.iters; ++it) {
            if (job.kind == JOB_FWD)
                run_forward(mp, g, gen0 + it, it == 0, it == job.iters - 1, job);
            else
                run_inverse(mp, g, gen0 + it, it == 0, it == job.iters - 1, job);
        }
        MCK(d, cudaStreamSynchronize(d.plan->stream));
        if (multi_prof() && g_prof[g].n) {
            static const char* const fwd_names[] = {"K1", "peer wait", "K2 (pull)", "K3", ""};
            static const char* const inv_names[] = {"K4", "K5 (push)", "peer wait", "K6", ""};
            g_prof[g].report(g, job.kind == JOB_FWD ? fwd_names : inv_names);
        }
        d.ms = 0.f;
        if (!d.status) cudaEventElapsedTime(&d.ms, d.ev_start, d.ev_stop);
        {
            std::lock_guard<std::mutex> lk(mp->mu);
            if (++mp->done == mp->G) mp->cv_done.notify_all();
        }
    }
}

int submit(s2kit_cuda_multi* mp, const Job& job) {
    {
        std::lock_guard<std::mutex> lk(mp->mu);
        mp->job = job;
        mp->done = 0;
        ++mp->job_seq;
    }
    mp->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lk(mp->mu);
        mp->cv_done.wait(lk, [&] { return mp->done == mp->G; });
        mp->gen += (unsigned)job.iters;
    }
    double worst = 0.0;
    for (auto& d : mp->dev) {
        if (d.status) return s2k_fail_msg(d.error.c_str());
        worst = d.ms > worst ? d.ms : worst;
    }
    mp->last_ms = worst / job.iters;
    return 0;
}

// owned runs of rank g in ascending position order
std::vector<Run> owned_runs(int bw, int G, int g, long* total) {
    std::vector<int> orders(bw);
    int cnt = s2kit_cuda_shard_layout(bw, G, g, orders.data(), nullptr);
    std::vector<Run> runs;
    long packed = 0;
    for (int i = 0; i < cnt; ++i) {
        const int m = orders[i];
        for (int sgn = 0; sgn < (m ? 2 : 1); ++sgn) {
            long at;
            if (!sgn)
                at = (long)m * bw - ((long)m * (m - 1)) / 2;
            else {
                const long big = bw - 1;
                at = (big * (big + 3)) / 2 + 1 + ((big - m) * (big - m + 1)) / 2;
            }
            runs.push_back({at, packed, bw - m});
            packed += bw - m;
        }
    }
    *total = packed;
    return runs;
}

}  // namespace

extern "C" int s2kit_cuda_multi_destroy(s2kit_cuda_multi* mp) {
    if (!mp) return 0;
    if (!mp->dev.empty() && mp->dev[0].th.joinable()) {
        {
            std::lock_guard<std::mutex> lk(mp->mu);
            mp->job.kind = JOB_EXIT;
            ++mp->job_seq;
        }
        mp->cv_go.notify_all();
        for (auto& d : mp->dev)
            if (d.th.joinable()) d.th.join();
    }
    int saved = -1;
    cudaGetDevice(&saved);
    for (auto& d : mp->dev) {
        cudaSetDevice(d.device);
        void* ptrs[] = {d.ring_r, d.ring_i, d.coef_r, d.coef_i, d.ringbuf[0], d.ringbuf[1], d.pack_r, d.pack_i, d.d_runs};
        for (void* q : ptrs)
            if (q) cudaFree(q);
        if (d.hpack_r) cudaFreeHost(d.hpack_r);
        if (d.hpack_i) cudaFreeHost(d.hpack_i);
        for (cudaEvent_t e : {d.ev_x[0], d.ev_x[1], d.ev_start, d.ev_stop})
            if (e) cudaEventDestroy(e);
        if (d.plan) s2kit_cuda_plan_destroy(d.plan);
    }
    if (saved >= 0) cudaSetDevice(saved);
    delete mp;
    return 0;
}

extern "C" int s2kit_cuda_multi_create(s2kit_cuda_multi** out, int bw, int ngpu, const int* devices) {
    if (!out) return s2k_fail_msg("null output pointer");
    *out = nullptr;
    if (ngpu < 1 || ngpu > s2k::S2K_MAX_PEERS) return s2k_fail_msg("ngpu must be in [1, 8]");
    if (s2kit_cuda_shard_layout(bw, ngpu, 0, nullptr, nullptr) < 0)
        return s2k_fail_msg("multi-GPU plans need power-of-two bw >= 16 and ngpu with (bw/2) % ngpu == 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return s2k_fail_msg("no CUDA device available: s2kit_cuda has no CPU fallback");
    int saved = -1;
    cudaGetDevice(&saved);
    s2kit_cuda_multi* mp = new s2kit_cuda_multi();
    mp->bw = bw;
    mp->n = 2 * bw;
    mp->G = ngpu;
    mp->nr = 2 * bw / ngpu;
    mp->block = 2L * mp->nr * mp->nr;
    mp->bar.n = ngpu;
    mp->dev.resize(ngpu);
    int rc = 0;
    auto fail = [&](int code) {
        s2kit_cuda_multi_destroy(mp);
        if (saved >= 0) cudaSetDevice(saved);
        return code;
    };
    for (int g = 0; g < ngpu; ++g) {
        mp->dev[g].device = devices ? devices[g] : g;
        if (mp->dev[g].device < 0 || mp->dev[g].device >= ndev) return fail(s2k_fail_msg("invalid device index"));
        for (int h = 0; h < g; ++h)
            if (mp->dev[h].device == mp->dev[g].device) return fail(s2k_fail_msg("devices must be distinct"));
    }
    // peer access between every pair
    for (int g = 0; g < ngpu; ++g) {
        cudaSetDevice(mp->dev[g].device);
        for (int h = 0; h < ngpu; ++h) {
            if (h == g) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, mp->dev[g].device, mp->dev[h].device);
            if (!can) return fail(s2k_fail_msg("the GPUs cannot access each other's memory (no peer access)"));
            cudaError_t e = cudaDeviceEnablePeerAccess(mp->dev[h].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else if (e != cudaSuccess)
                return fail(s2k_fail_cuda("cudaDeviceEnablePeerAccess", e));
        }
    }
    const long n = mp->n, nr = mp->nr, cs = (long)bw * bw;
    for (int g = 0; g < ngpu && !rc; ++g) {
        s2kit_cuda_multi::Dev& d = mp->dev[g];
        rc = s2kit_cuda_plan_create_sharded(&d.plan, bw, S2KIT_CUDA_MEMO, d.device, g, ngpu);
        if (rc) break;
        cudaSetDevice(d.device);
        d.runs = owned_runs(bw, ngpu, g, &d.packed_len);
        cudaError_t e = cudaSuccess;
        auto dm = [&](double** q, size_t doubles) {
            if (e == cudaSuccess) e = cudaMalloc((void**)q, doubles * sizeof(double));
            if (e == cudaSuccess) e = cudaMemset(*q, 0, doubles * sizeof(double));
        };
        dm(&d.ring_r, (size_t)nr * n);
        dm(&d.ring_i, (size_t)nr * n);
        dm(&d.coef_r, (size_t)cs);
        dm(&d.coef_i, (size_t)cs);
        dm(&d.ringbuf[0], (size_t)ngpu * mp->block);
        dm(&d.ringbuf[1], (size_t)ngpu * mp->block);
        dm(&d.pack_r, (size_t)d.packed_len);
        dm(&d.pack_i, (size_t)d.packed_len);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&d.hpack_r, sizeof(double) * d.packed_len);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&d.hpack_i, sizeof(double) * d.packed_len);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d.d_runs, sizeof(Run) * d.runs.size());
        if (e == cudaSuccess)
            e = cudaMemcpy(d.d_runs, d.runs.data(), sizeof(Run) * d.runs.size(), cudaMemcpyHostToDevice);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&d.ev_x[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev_start);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev_stop);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = s2k_fail_cuda("multi-GPU plan setup", e);
    }
    if (rc) return fail(rc);
    for (int g = 0; g < ngpu; ++g) mp->dev[g].th = std::thread(worker, mp, g);
    if (saved >= 0) cudaSetDevice(saved);
    *out = mp;
    return 0;
}

extern "C" int s2kit_cuda_multi_ngpu(const s2kit_cuda_multi* mp) { return mp ? mp->G : 0; }
extern "C" size_t s2kit_cuda_multi_table_bytes_per_gpu(const s2kit_cuda_multi* mp) {
    return mp ? s2kit_cuda_plan_table_bytes(mp->dev[0].plan) : 0;
}

// forward SHT of one field, host pointers: rdata / idata full 2bw x 2bw grids, rcoeffs / icoeffs full bw*bw arrays
extern "C" int s2kit_cuda_multi_fst(s2kit_cuda_multi* mp, const double* rdata, const double* idata, double* rcoeffs,
                                    double* icoeffs) {
    if (!mp) return s2k_fail_msg("null multi-GPU plan");
    if (!rdata || !idata || !rcoeffs || !icoeffs) return s2k_fail_msg("null pointer");
    std::lock_guard<std::mutex> lock(mp->api_mu);
    Job job;
    job.kind = JOB_FWD;
    job.host_io = true;
    job.h_in_r = rdata;
    job.h_in_i = idata;
    if (int rc = submit(mp, job)) return rc;
    for (auto& d : mp->dev)
        for (const Run& r : d.runs) {
            memcpy(rcoeffs + r.at, d.hpack_r + r.packed, sizeof(double) * r.len);
            memcpy(icoeffs + r.at, d.hpack_i + r.packed, sizeof(double) * r.len);
        }
    return 0;
}

extern "C" int s2kit_cuda_multi_inv_fst(s2kit_cuda_multi* mp, const double* rcoeffs, const double* icoeffs,
                                        double* rdata, double* idata) {
    if (!mp) return s2k_fail_msg("null multi-GPU plan");
    if (!rdata || !idata || !rcoeffs || !icoeffs) return s2k_fail_msg("null pointer");
    std::lock_guard<std::mutex> lock(mp->api_mu);
    for (auto& d : mp->dev)
        for (const Run& r : d.runs) {
            memcpy(d.hpack_r + r.packed, rcoeffs + r.at, sizeof(double) * r.len);
            memcpy(d.hpack_i + r.packed, icoeffs + r.at, sizeof(double) * r.len);
        }
    Job job;
    job.kind = JOB_INV;
    job.host_io = true;
    job.h_out_r = rdata;
    job.h_out_i = idata;
    return submit(mp, job);
}

// Device-resident form: the data stays in the plan's own per-GPU buffers (s2kit_cuda_multi_buffers); `iters` transforms
// run back to back.  *ms_per_transform = device time (CUDA events on each GPU's stream, maximum over the GPUs).
extern "C" int s2kit_cuda_multi_run(s2kit_cuda_multi* mp, int inverse, int iters, double* ms_per_transform) {
    if (!mp) return s2k_fail_msg("null multi-GPU plan");
    if (iters < 1) return s2k_fail_msg("iters must be positive");
    std::lock_guard<std::mutex> lock(mp->api_mu);
    Job job;
    job.kind = inverse ? JOB_INV : JOB_FWD;
    job.iters = iters;
    if (int rc = submit(mp, job)) return rc;
    if (ms_per_transform) *ms_per_transform = mp->last_ms;
    return 0;
}

// GPU g's buffers: its latitude rings [2bw/G][2bw] (re, im) and its full-size coefficient arrays, of which only the
// owned orders are read / written.  Device pointers on device *device.
extern "C" int s2kit_cuda_multi_buffers(s2kit_cuda_multi* mp, int g, int* device, double** ring_r, double** ring_i,
                                        double** coef_r, double** coef_i) {
    if (!mp || g < 0 || g >= mp->G) return s2k_fail_msg("bad GPU index");
    const s2kit_cuda_multi::Dev& d = mp->dev[g];
    if (device) *device = d.device;
    if (ring_r) *ring_r = d.ring_r;
    if (ring_i) *ring_i = d.ring_i;
    if (coef_r) *coef_r = d.coef_r;
    if (coef_i) *coef_i = d.coef_i;
    return 0;
}
