// shard.cu -- one transform split over the GPUs of one box (single large-bandwidth field).
//
// The reference has no parallelism of any kind; this is the B200-side answer to bw = 2048, whose 11.5 GB table
// is streamed once per transform: the stream is divided over G GPUs.  Work is partitioned twice:
//   * latitude rings  j in [r*2bw/G, (r+1)*2bw/G)  for the longitude FFT (K1 / K6), and
//   * orders, dealt in pairs (q, bw-1-q) -> rank q mod G (work of order m ~ bw^2 - m^2, so a pair is balanced),
//     for the DCT + Legendre stages (K2-K5); a rank keeps only its own orders' tables.
// Between the two sits ONE exchange step, an all-to-all of equal blocks [peer][part][local row][local ring]
// (2bw/G x 2bw/G complex values per pair of ranks: 4 MiB at bw = 2048, G = 8).  The exchange itself is the caller's
// (NCCL all_to_all over NVLink, see bench_single_field.py / tests); the kernels here read and write the blocks in
// place through PlaneView addressing, so no pack / unpack pass exists.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "s2k_internal.cuh"

int s2k_plan_create_impl(s2kit_cuda_plan** out, int bw, int variant, int max_batch, int device, int rank, int nranks);
int s2k_fail_msg(const char* what);
int s2k_fail_cuda(const char* what, cudaError_t e);

namespace {

int ilog2i(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// orders owned by `rank`, ascending
std::vector<int> orders_of(int bw, int nranks, int rank) {
    std::vector<int> o;
    for (int q = 0; q < (bw + 1) / 2; ++q) {
        if (q % nranks != rank) continue;
        o.push_back(q);
        if (bw - 1 - q != q) o.push_back(bw - 1 - q);
    }
    std::sort(o.begin(), o.end());
    return o;
}

// spectral rows of `rank` in local order: for each owned order m its row m and (m > 0) the row 2bw - m of order -m
std::vector<int> rows_of(int bw, int nranks, int rank) {
    std::vector<int> r;
    for (int m : orders_of(bw, nranks, rank)) {
        r.push_back(m);
        if (m > 0) r.push_back(2 * bw - m);
    }
    return r;
}

bool shard_ok(int bw, int nranks) {
    if (nranks < 1 || (nranks & (nranks - 1))) return false;
    if (bw < 16 || (bw & (bw - 1))) return false;
    return (bw / 2) % nranks == 0 && (2 * bw / nranks) >= 8;
}

}  // namespace

#include "s2k_shard.cuh"


extern "C" int s2kit_cuda_shard_layout(int bw, int nranks, int rank, int* orders_out, int* rows_out) {
    if (!shard_ok(bw, nranks) || rank < 0 || rank >= nranks) return -1;
    std::vector<int> o = orders_of(bw, nranks, rank), r = rows_of(bw, nranks, rank);
    if (orders_out) std::copy(o.begin(), o.end(), orders_out);
    if (rows_out) {
        std::copy(r.begin(), r.end(), rows_out);
        for (int i = (int)r.size(); i < 2 * bw / nranks; ++i) rows_out[i] = -1;
    }
    return (int)o.size();
}

extern "C" int s2kit_cuda_plan_create_sharded(s2kit_cuda_plan** out, int bw, int variant, int device, int rank,
                                              int nranks) {
    if (out) *out = nullptr;
    if (!shard_ok(bw, nranks) || rank < 0 || rank >= nranks)
        return s2k_fail_msg("sharded plans need power-of-two bw >= 16 and nranks with (bw/2) % nranks == 0");
    if (variant != S2KIT_CUDA_MEMO) return s2k_fail_msg("sharded plans keep resident (Memo) tables");
    if (int rc = s2k_plan_create_impl(out, bw, variant, 1, device, rank, nranks)) return rc;
    s2kit_cuda_plan* p = *out;
    ShardState* st = new ShardState();
    p->shard = st;
    const int n = 2 * bw, nr = n / nranks;
    st->nr = nr;
    st->block = 2L * nr * nr;
    // row -> (owner, local row) for every spectral row; row bw (Nyquist) belongs to nobody
    std::vector<long> rowbase(n, 0);
    for (int r = 0; r < nranks; ++r) {
        std::vector<int> rows = rows_of(bw, nranks, r);
        for (size_t i = 0; i < rows.size(); ++i) rowbase[rows[i]] = ((long)r * 2 * nr + (long)i) * nr;
    }
    std::vector<int> mine = rows_of(bw, nranks, rank), orders = orders_of(bw, nranks, rank);
    st->nrows_real = (int)mine.size();
    st->norders = (int)orders.size();
    mine.resize(nr, bw);  // pad slot points at the dead Nyquist row; never launched
    cudaError_t e = cudaMalloc((void**)&st->d_rowbase, sizeof(long) * n);
    if (e == cudaSuccess) e = cudaMemcpy(st->d_rowbase, rowbase.data(), sizeof(long) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_rowlist, sizeof(int) * nr);
    if (e == cudaSuccess) e = cudaMemcpy(st->d_rowlist, mine.data(), sizeof(int) * nr, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_orders, sizeof(int) * orders.size());
    if (e == cudaSuccess)
        e = cudaMemcpy(st->d_orders, orders.data(), sizeof(int) * orders.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // pageable uploads have landed before any stream reads them
    if (e != cudaSuccess) {
        s2kit_cuda_plan_destroy(p);
        *out = nullptr;
        return s2k_fail_cuda("sharded plan setup", e);
    }
    // K1 / K6 address send / receive blocks by order row
    st->ring_view.rowbase = st->d_rowbase;
    st->ring_view.rowlist = nullptr;
    st->ring_view.part_stride = (long)nr * nr;
    st->ring_view.lrow_stride = nr;
    st->ring_view.seg_stride = 0;
    st->ring_view.seg_shift = 30;
    st->ring_view.seg_mask = 0x3fffffff;
    st->ring_view.nrings = nr;
    // K2 / K5 walk one local row across the blocks of all peers
    st->order_view.rowbase = nullptr;
    st->order_view.rowlist = st->d_rowlist;
    st->order_view.part_stride = (long)nr * nr;
    st->order_view.lrow_stride = nr;
    st->order_view.seg_stride = 2L * nr * nr;
    st->order_view.seg_shift = ilog2i(nr);
    st->order_view.seg_mask = nr - 1;
    st->order_view.nrings = nr;
    return 0;
}

void s2k_shard_destroy(s2kit_cuda_plan* p) {
    ShardState* st = shard_of(p);
    if (!st) return;
    if (st->d_rowbase) cudaFree(st->d_rowbase);
    if (st->d_rowlist) cudaFree(st->d_rowlist);
    if (st->d_orders) cudaFree(st->d_orders);
    delete st;
    p->shard = nullptr;
}

extern "C" int s2kit_cuda_shard_info(const s2kit_cuda_plan* p, long* block_doubles, int* rings_per_rank,
                                     int* rows_per_rank) {
    if (!p || !shard_of(p)) return s2k_fail_msg("not a sharded plan");
    if (block_doubles) *block_doubles = shard_of(p)->block;
    if (rings_per_rank) *rings_per_rank = shard_of(p)->nr;
    if (rows_per_rank) *rows_per_rank = shard_of(p)->nr;
    return 0;
}

#define CKS(call)                                              \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return s2k_fail_cuda(#call, e__); \
    } while (0)

// forward, stage 1: local rings (rdata/idata: [nr][2bw]) -> send blocks
extern "C" int s2kit_cuda_fst_rings(s2kit_cuda_plan* p, const double* rdata, const double* idata, double* sendbuf) {
    if (!p || !shard_of(p)) return s2k_fail_msg("not a sharded plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CKS(cudaSetDevice(p->device));
    ShardState* st = shard_of(p);
    CKS(s2k::launch_phi_fft_fwd(p, rdata, idata, 0, sendbuf, 1, S2KIT_COMPLEX, &st->ring_view));
    return 0;
}

// forward, stage 2: receive blocks -> coefficients of this rank's orders (positions of the full bw*bw arrays)
extern "C" int s2kit_cuda_fst_orders(s2kit_cuda_plan* p, const double* recvbuf, double* rcoeffs, double* icoeffs) {
    if (!p || !shard_of(p)) return s2k_fail_msg("not a sharded plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CKS(cudaSetDevice(p->device));
    ShardState* st = shard_of(p);
    CKS(s2k::launch_dct_fwd(p, recvbuf, p->d_X, 1, 0, st->nrows_real, S2KIT_COMPLEX, &st->order_view));
    CKS(s2k::launch_legendre_fwd(p, p->d_table, 0, p->d_X, rcoeffs, icoeffs, (long)p->bw * p->bw, 1, 0, st->norders,
                                 S2KIT_COMPLEX, st->d_orders));
    return 0;
}

// inverse, stage 1: coefficients of this rank's orders -> send blocks
extern "C" int s2kit_cuda_inv_fst_orders(s2kit_cuda_plan* p, const double* rcoeffs, const double* icoeffs,
                                         double* sendbuf) {
    if (!p || !shard_of(p)) return s2k_fail_msg("not a sharded plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CKS(cudaSetDevice(p->device));
    ShardState* st = shard_of(p);
    CKS(s2k::launch_legendre_inv(p, p->d_table_t, 0, rcoeffs, icoeffs, (long)p->bw * p->bw, p->d_X, 1, 0, st->norders,
                                 S2KIT_COMPLEX, st->d_orders));
    CKS(s2k::launch_dct_inv(p, p->d_X, sendbuf, 1, 0, st->nrows_real, S2KIT_COMPLEX, &st->order_view));
    return 0;
}

// inverse, stage 2: receive blocks -> local rings
extern "C" int s2kit_cuda_inv_fst_rings(s2kit_cuda_plan* p, const double* recvbuf, double* rdata, double* idata) {
    if (!p || !shard_of(p)) return s2k_fail_msg("not a sharded plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CKS(cudaSetDevice(p->device));
    ShardState* st = shard_of(p);
    CKS(s2k::launch_phi_fft_inv(p, recvbuf, rdata, idata, 0, 1, S2KIT_COMPLEX, &st->ring_view));
    return 0;
}
