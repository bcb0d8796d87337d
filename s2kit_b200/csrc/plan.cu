// plan.cu -- plan object, table/workspace management and the s2kit_cuda_* C-ABI (include/s2kit_cuda.h).
//
// A plan owns everything the reference makes the caller carry around: the quadrature weights
// (GenerateWeightsForDLT), the cosine tables (Spharmonic_Pml_Table & co., src/legendre_polynomials/cospml.c:387-518),
// the FFT/DCT "plans" (FFTW descriptors in the reference) and the workspaces.  Device tables are generated
// once per plan (Memo) or per call into a bounded scratch ring (Fly, the reference's O(bw^2)-memory
// variant, src/FST_semi_fly.c:96,259-261).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "host_setup.h"
#include "s2k_internal.cuh"

static thread_local std::string g_last_error;

static int fail(const char* what, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    g_last_error = buf;
    return 1;
}
static int fail_msg(const char* what) {
    g_last_error = what;
    return 2;
}

#define CK(call)                                                 \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return fail(#call, e__);         \
    } while (0)

extern "C" const char* s2kit_cuda_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* s2kit_cuda_version(void) { return "s2kit_b200 0.1 (sm_100a, FP64 DMMA)"; }

// ------------------------------------------------------------------------------------------------ profiling
#include <map>
#include <mutex>
namespace s2k {
cudaError_t ensure_smem(const void* kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[key] = bytes;
    return e;
}

void prof_collect_slots(s2kit_cuda_plan* p);
int prof_begin(s2kit_cuda_plan* p, int kind) {
    if (!p->profiling) return -1;
    if (p->prof_used >= 4096) {  // bounded: fold the finished brackets into the totals and recycle the slots
        cudaStreamSynchronize(p->stream);
        prof_collect_slots(p);
    }
    if (p->prof_used == p->prof_slots.size()) {
        ProfileSlot s;
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        s.kind = kind;
        p->prof_slots.push_back(s);
    }
    int idx = (int)p->prof_used++;
    p->prof_slots[idx].kind = kind;
    cudaEventRecord(p->prof_slots[idx].a, p->stream);
    return idx;
}
void prof_end(s2kit_cuda_plan* p, int slot) {
    if (slot >= 0) cudaEventRecord(p->prof_slots[slot].b, p->stream);
}
}  // namespace s2k

void s2k::prof_collect_slots(s2kit_cuda_plan* p) {
    for (size_t i = 0; i < p->prof_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p->prof_slots[i].a, p->prof_slots[i].b) == cudaSuccess) {
            p->prof_ms[p->prof_slots[i].kind] += ms;
            p->prof_launches[p->prof_slots[i].kind] += 1;
        }
    }
    p->prof_used = 0;
}

extern "C" int s2kit_cuda_profile_enable(s2kit_cuda_plan* p, int on) {
    if (!p) return fail_msg("null plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    p->profiling = on != 0;
    return 0;
}
extern "C" int s2kit_cuda_profile_reset(s2kit_cuda_plan* p) {
    if (!p) return fail_msg("null plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CK(cudaStreamSynchronize(p->stream));
    p->prof_used = 0;
    for (int k = 0; k < S2KIT_K_COUNT; ++k) {
        p->prof_ms[k] = 0;
        p->prof_launches[k] = 0;
    }
    return 0;
}
extern "C" int s2kit_cuda_profile_get(s2kit_cuda_plan* p, double* ms, long* launches) {
    if (!p) return fail_msg("null plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    CK(cudaStreamSynchronize(p->stream));
    s2k::prof_collect_slots(p);
    for (int k = 0; k < S2KIT_K_COUNT; ++k) {
        if (ms) ms[k] = p->prof_ms[k];
        if (launches) launches[k] = p->prof_launches[k];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ layout
static int row_size(int m, int l) {  // RowSize, cospml.c:250-258
    if (l < m) return 0;
    return (m & 1) ? (l - 1) / 2 + 1 : l / 2 + 1;
}

static void build_layout(s2kit_cuda_plan* p, const std::vector<char>& owned) {
    const int bw = p->bw;
    p->h_meta.assign(2 * bw, s2k::BlockMeta{0, 0, 0, 0});
    p->h_rt_start.clear();
    p->h_order_start.assign(bw + 1, 0);
    for (int m = 0; m < bw; ++m) {
        uint32_t cur = 0;
        for (int par = 0; par < 2; ++par) {
            s2k::BlockMeta& mb = p->h_meta[2 * m + par];
            int rows = (bw - m - par + 1) / 2;
            if (rows < 0) rows = 0;
            mb.rows = rows;
            mb.len0 = row_size(m, m + par);
            mb.nrt = (rows + 7) / 8;
            mb.rt_base = (int)p->h_rt_start.size();
            for (int rt = 0; rt < mb.nrt; ++rt) {
                p->h_rt_start.push_back(cur);
                int last_row = std::min(8 * rt + 7, rows - 1);
                cur += (uint32_t)((mb.len0 + last_row + 7) >> 3);
            }
        }
        p->h_order_start[m + 1] = p->h_order_start[m] + (owned[m] ? cur : 0);
    }
    const int lch = s2k::table_unit_rows(bw);
    p->h_units.clear();
    p->h_unit_first.assign(bw + 1, 0);
    for (int m = 0; m < bw; ++m) {
        p->h_unit_first[m] = (int)p->h_units.size() / 2;
        // longest recurrence roll-up first: a launch ends with its slowest CTA
        int last = m + ((bw - 1 - m) / lch) * lch;
        for (int l0 = last; l0 >= m; l0 -= lch) {
            p->h_units.push_back(m);
            p->h_units.push_back(l0);
        }
    }
    p->h_unit_first[bw] = (int)p->h_units.size() / 2;
}

// Work lists of the uniform-warp kernels (kernels_uni.cu).  An SM sub-partition has its own FP64 / DMMA pipe and a warp
// never leaves its sub-partition, so the units of one order are split into FOUR queues of equal cost (longest
// processing time first into the lightest queue), one per sub-partition; the four warps of a sub-partition take their
// queue's units heaviest first.  off[4 m + s] .. off[4 m + s + 1]: queue s of order m inside `list`.
static void lpt_queues(std::vector<std::pair<int, int>>& items, std::vector<int>& off, std::vector<unsigned short>& list) {
    std::stable_sort(items.begin(), items.end(),
                     [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
    std::vector<unsigned short> q[4];
    long load[4] = {0, 0, 0, 0};
    for (auto& it : items) {
        int best = 0;
        for (int s = 1; s < 4; ++s)
            if (load[s] < load[best]) best = s;
        q[best].push_back((unsigned short)it.second);
        load[best] += it.first + 1;  // + 1: a unit costs its set-up even when it has no tiles
    }
    for (int s = 0; s < 4; ++s) {
        list.insert(list.end(), q[s].begin(), q[s].end());
        off.push_back((int)list.size());
    }
}

static int uni_cap() {
    int cap = 16;
    if (const char* e = getenv("S2KIT_CUDA_UNI_CAP")) cap = std::max(0, atoi(e));
    return cap;
}

// forward: (parity, row tile) units, long rows alone, short rows in adjacent pairs of at most `cap` column-tile steps
static void build_subitems(s2kit_cuda_plan* p, std::vector<int>& off, std::vector<unsigned short>& list) {
    const int bw = p->bw, cap = uni_cap();
    off.assign(1, 0);
    list.clear();
    for (int m = 0; m < bw; ++m) {
        std::vector<std::pair<int, int>> items;  // (cost, code = parity | row tile << 1 | pair << 12)
        for (int par = 0; par < 2; ++par) {
            const s2k::BlockMeta& mb = p->h_meta[2 * m + par];
            auto tiles = [&](int rt) { return (mb.len0 + std::min(8 * rt + 7, mb.rows - 1) + 7) >> 3; };
            int rt = mb.nrt - 1;
            while (rt >= 0) {
                if (rt >= 1 && tiles(rt) + tiles(rt - 1) <= cap) {
                    items.push_back({tiles(rt) + tiles(rt - 1), par | (rt << 1) | (1 << 12)});
                    rt -= 2;
                } else {
                    items.push_back({tiles(rt), par | (rt << 1)});
                    rt -= 1;
                }
            }
        }
        lpt_queues(items, off, list);
    }
}

// inverse: (parity, column tile) units, every column tile of the panel included (a tile no row reaches still has to be
// zeroed), tiles many rows reach alone, the others in adjacent pairs
static void build_inv_subitems(s2kit_cuda_plan* p, std::vector<int>& off, std::vector<unsigned short>& list) {
    const int bw = p->bw, nct = (((bw + 1) / 2) + 7) >> 3, cap = uni_cap();
    off.assign(1, 0);
    list.clear();
    for (int m = 0; m < bw; ++m) {
        std::vector<std::pair<int, int>> items;  // code = parity | column tile << 1 | pair << 12
        for (int par = 0; par < 2; ++par) {
            const s2k::BlockMeta& mb = p->h_meta[2 * m + par];
            auto tiles = [&](int rt) { return (mb.len0 + std::min(8 * rt + 7, mb.rows - 1) + 7) >> 3; };
            auto reach = [&](int ct) {  // row tiles that reach column tile ct (first_row_tile_reaching, s2k_legendre.cuh)
                int rt_min = 0;
                if (8 * ct >= mb.len0 + 7) rt_min = (8 * ct - mb.len0 - 7) / 8 + 1;
                while (rt_min < mb.nrt && ct >= tiles(rt_min)) ++rt_min;
                return std::max(0, mb.nrt - rt_min);
            };
            int ct = 0;
            while (ct < nct) {
                if (ct + 1 < nct && reach(ct) + reach(ct + 1) <= cap) {
                    items.push_back({reach(ct) + reach(ct + 1), par | (ct << 1) | (1 << 12)});
                    ct += 2;
                } else {
                    items.push_back({reach(ct), par | (ct << 1)});
                    ct += 1;
                }
            }
        }
        lpt_queues(items, off, list);
    }
}

// inverse, persistent contraction kernel (kernels_flow.cu): units of FOUR adjacent column tiles of one parity block
static void build_invq_subitems(s2kit_cuda_plan* p, std::vector<int>& off, std::vector<unsigned short>& list) {
    const int bw = p->bw, nct = (((bw + 1) / 2) + 7) >> 3;
    off.assign(1, 0);
    list.clear();
    for (int m = 0; m < bw; ++m) {
        std::vector<std::pair<int, int>> items;  // code = parity | quad << 1
        for (int par = 0; par < 2; ++par) {
            const s2k::BlockMeta& mb = p->h_meta[2 * m + par];
            auto tiles = [&](int rt) { return (mb.len0 + std::min(8 * rt + 7, mb.rows - 1) + 7) >> 3; };
            auto reach = [&](int ct) {
                int rt_min = 0;
                if (8 * ct >= mb.len0 + 7) rt_min = (8 * ct - mb.len0 - 7) / 8 + 1;
                while (rt_min < mb.nrt && ct >= tiles(rt_min)) ++rt_min;
                return std::max(0, mb.nrt - rt_min);
            };
            for (int q = 0; 4 * q < nct; ++q) {
                int cost = 0;  // DMMA column-tile steps of the quad
                for (int c = 4 * q; c < std::min(nct, 4 * q + 4); ++c) cost += reach(c);
                items.push_back({cost, par | (q << 1)});
            }
        }
        lpt_queues(items, off, list);
    }
}

// Keep a Memo table that fits comfortably in L2 resident across launches: the batch streams hundreds of MB through
// L2 between two uses of the same order's tiles (persisting-L2 access policy window on the plan's stream).
static void apply_table_l2_policy(s2kit_cuda_plan* p) {
    if (p->variant != S2KIT_CUDA_MEMO || !p->d_table || !p->l2_persist) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess) return;
    size_t want = p->table_bytes;
    if (prop.persistingL2CacheMaxSize <= 0 || want > (size_t)prop.accessPolicyMaxWindowSize) return;
    size_t carve = std::min(want, (size_t)prop.persistingL2CacheMaxSize);
    if (carve * 4 < want) return;  // far larger than the carve-out: streamed anyway
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = p->d_table;
    attr.accessPolicyWindow.num_bytes = want;
    attr.accessPolicyWindow.hitRatio = (float)((double)carve / (double)want);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(p->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

// Tensor map of the spectral-plane workspace for the TMA stores of K1 (kernels_fft.cu).  The driver entry point is
// looked up at run time so the library does not link libcuda; on failure K1 simply keeps its LSU store path.
static void make_plane_tensor_map(s2kit_cuda_plan* p) {
    p->tma_S_ok = false;
    const char* off = getenv("S2KIT_CUDA_NO_TMA");
    if (off && off[0] == '1') return;
    if (!p->fast || p->n > 512 || p->n < 64) return;
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return;
    }
    const cuuint64_t n = (cuuint64_t)p->n;
    cuuint64_t gdim[3] = {n, n, (cuuint64_t)p->chunk * 2};
    cuuint64_t gstride[2] = {n * sizeof(double), n * n * sizeof(double)};
    cuuint32_t box[3] = {8, (cuuint32_t)(p->n < 256 ? p->n : 256), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<encode_fn>(fn)(&p->tma_S, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p->d_S, gdim, gstride, box,
                                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p->tma_S_ok = (r == CUDA_SUCCESS);
}

// Uploads go through the plan's own (non-blocking) stream, so the kernels launched on it afterwards are ordered behind
// them; the host buffers are temporaries, hence the synchronisation before returning.
template <typename T>
static cudaError_t upload(cudaStream_t st, T** dptr, const void* host, size_t count) {
    cudaError_t e = cudaMalloc((void**)dptr, count * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

void s2k_shard_destroy(s2kit_cuda_plan* p);
int s2k_fail_msg(const char* what) { return fail_msg(what); }
int s2k_fail_cuda(const char* what, cudaError_t e) { return fail(what, e); }

// The caller's current device is left as it was found (plan creation and every entry point switch to the plan's).
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() {
        if (cudaGetDevice(&saved) != cudaSuccess) saved = -1;
    }
    ~DeviceGuard() {
        if (saved >= 0) cudaSetDevice(saved);
    }
};

static int plan_build(s2kit_cuda_plan* p, int bw, int variant, int max_batch, int device, int rank, int nranks);

int s2k_plan_create_impl(s2kit_cuda_plan** out, int bw, int variant, int max_batch, int device, int rank,
                         int nranks) {
    if (!out) return fail_msg("null output pointer");
    *out = nullptr;
    if (bw < 2 || bw > 2048) return fail_msg("bandwidth must be in [2, 2048]");
    if (variant != S2KIT_CUDA_MEMO && variant != S2KIT_CUDA_FLY) return fail_msg("unknown variant");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0)
        return fail_msg("no CUDA device available: s2kit_cuda has no CPU fallback");
    if (device < 0 || device >= ndev) return fail_msg("invalid device index");
    DeviceGuard guard;
    CK(cudaSetDevice(device));
    s2kit_cuda_plan* p = new s2kit_cuda_plan();
    p->mu = new std::mutex();
    int rc = plan_build(p, bw, variant, max_batch, device, rank, nranks);
    if (rc) {
        // every partial allocation (stream, GBs of tables at large bw) goes back; the error text survives
        std::string keep = g_last_error;
        s2kit_cuda_plan_destroy(p);
        g_last_error = keep;
        return rc;
    }
    *out = p;
    return 0;
}

static int plan_build(s2kit_cuda_plan* p, int bw, int variant, int max_batch, int device, int rank, int nranks) {
    p->bw = bw;
    p->n = 2 * bw;
    p->variant = variant;
    p->device = device;
    p->rank = rank;
    p->nranks = nranks;
    p->fast = is_pow2(bw) && bw >= 16;
    {
        // persisting-L2 carve-out for the tables: measured neutral with one table copy and harmful with two (it takes
        // L2 away from the streaming kernels), the per-CTA bulk prefetch already does the job -- opt-in
        const char* np = getenv("S2KIT_CUDA_L2PERSIST");
        p->l2_persist = (np && np[0] == '1');
        const char* sp = getenv("S2KIT_CUDA_SPLIT");
        p->nsplit = sp ? std::max(1, std::min(16, atoi(sp))) : 1;
    }
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        p->sm_count = prop.multiProcessorCount;
    }
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    const int n = p->n;

    // ---- host-computed seeds (libm), uploaded
    {
        std::vector<double> w(4 * bw), s(n), x(bw), tw(2 * (size_t)n), tb(2 * (size_t)bw), qn(8 * (size_t)n),
            qb(8 * (size_t)bw);
        s2k_host_weights(bw, w.data());
        s2k_host_sines(bw, s.data());
        s2k_host_nodes(bw, x.data());
        s2k_host_twiddles(n, tw.data());
        s2k_host_twiddles(bw, tb.data());
        s2k_host_quarter(n, qn.data());
        s2k_host_quarter(bw, qb.data());
        CK(upload(p->stream, &p->d_weights, w.data(), w.size()));
        CK(upload(p->stream, &p->d_sin, s.data(), s.size()));
        std::vector<double> wv(4 * bw), sv(n);
        s2k_host_reordered(bw, w.data(), s.data(), wv.data(), sv.data());
        CK(upload(p->stream, &p->d_wv, wv.data(), wv.size()));
        CK(upload(p->stream, &p->d_sv, sv.data(), sv.size()));
        CK(upload(p->stream, &p->d_nodes, x.data(), x.size()));
        CK(upload(p->stream, &p->d_tw_n, tw.data(), (size_t)n));
        CK(upload(p->stream, &p->d_tw_b, tb.data(), (size_t)bw));
        CK(upload(p->stream, &p->d_q_n, qn.data(), 4 * (size_t)n));
        CK(upload(p->stream, &p->d_q_b, qb.data(), 4 * (size_t)bw));
        std::vector<double> seeds((size_t)bw * bw);
        s2k_host_seeds(bw, 0, bw, seeds.data());
        CK(upload(p->stream, &p->d_seeds, seeds.data(), seeds.size()));
    }
    CK(cudaMalloc((void**)&p->d_rec, sizeof(double2) * (size_t)bw * bw));

    // ---- which orders this plan owns
    std::vector<char> owned(bw, 1);
    if (nranks > 1) {
        std::fill(owned.begin(), owned.end(), 0);
        // pair m with bw-1-m (work ~ bw^2 - m^2): pair q goes to rank q % nranks
        for (int q = 0; q < (bw + 1) / 2; ++q) {
            if (q % nranks != rank) continue;
            owned[q] = 1;
            owned[bw - 1 - q] = 1;
        }
    }
    for (int m = 0; m < bw; ++m)
        if (owned[m]) p->my_orders.push_back(m);
    build_layout(p, owned);
    CK(upload(p->stream, &p->d_meta, p->h_meta.data(), p->h_meta.size()));
    CK(upload(p->stream, &p->d_rt_start, p->h_rt_start.data(), p->h_rt_start.size()));
    CK(upload(p->stream, &p->d_order_start, p->h_order_start.data(), p->h_order_start.size()));
    CK(upload(p->stream, &p->d_units, p->h_units.data(), p->h_units.size()));
    if (p->fast && p->n == 512 && nranks == 1) {
        std::vector<int> off;
        std::vector<unsigned short> list;
        build_subitems(p, off, list);
        p->n_sub_list = (int)list.size();
        CK(upload(p->stream, &p->d_sub_off, off.data(), off.size()));
        CK(upload(p->stream, &p->d_sub_list, list.data(), list.size()));
        build_inv_subitems(p, off, list);
        p->n_isub_list = (int)list.size();
        CK(upload(p->stream, &p->d_isub_off, off.data(), off.size()));
        CK(upload(p->stream, &p->d_isub_list, list.data(), list.size()));
        build_invq_subitems(p, off, list);
        p->n_iq_list = (int)list.size();
        CK(upload(p->stream, &p->d_iq_off, off.data(), off.size()));
        CK(upload(p->stream, &p->d_iq_list, list.data(), list.size()));
    }
    CK(s2k::launch_rec_coeffs(p));

    // ---- tables
    const uint64_t total_tiles = p->h_order_start[bw];
    if (variant == S2KIT_CUDA_MEMO) {
        // bw < 512: two copies of the tiles in one allocation, A-fragment order for the forward contraction and
        // B-fragment (tile-transposed) order for the wide batched inverse kernels (26 MB at bw = 256).  bw >= 512 (where
        // the table is what fills the memory: 11.7 GB at bw = 2048): ONE copy, the inverse gathers its fragments from it
        p->table_tiles = total_tiles;
        {
            const char* tc = getenv("S2KIT_CUDA_TABLE_COPIES");
            p->table_single = bw >= 512 && !(tc && tc[0] == '2');
        }
        p->table_bytes = (p->table_single ? 1 : 2) * total_tiles * 64 * sizeof(double);
        CK(cudaMalloc((void**)&p->d_table, p->table_bytes));
        p->d_table_t = p->table_single ? p->d_table : p->d_table + total_tiles * 64;
        // generate every run of consecutive owned orders
        int m = 0;
        while (m < bw) {
            if (!owned[m]) {
                ++m;
                continue;
            }
            int hi = m;
            while (hi < bw && owned[hi]) ++hi;
            CK(s2k::launch_table_gen(p, p->d_table, 0, m, hi, 0));
            if (!p->table_single) CK(s2k::launch_table_gen(p, p->d_table_t, 0, m, hi, 1));
            m = hi;
        }
    } else {
        // Fly: scratch table for a group of orders.  A group must be large enough to fill the GPU with generator CTAs
        // (two 256-thread CTAs per SM, one per 2 x 64 degrees of one order): at bw = 1024 a 64 MiB group gave 221 CTAs
        // per launch, 17 % of the warp slots, and 23 launches per direction each waiting for its slowest CTA (ncu,
        // profiles/r1_ncu_summary.md section 6).  A 1 GiB group no longer stays in L2, but writing and re-reading the
        // tiles through HBM costs far less than the idle SMs did: bw = 1024 forward 5.7 -> 3.1 ms (64 MiB -> 1 GiB).
        uint64_t biggest = 0;
        for (int m = 0; m < bw; ++m) biggest = std::max(biggest, p->h_order_start[m + 1] - p->h_order_start[m]);
        // bw >= 512: the inverse contraction reads A-order tiles too (the narrow-panel reader of the one-copy Memo plans), so
        // both directions generate the same layout and the generator's stores stay 64-byte runs (in B-fragment order a
        // table row is scattered over eight sectors: bw = 1024 inverse generation 1.9 vs 1.4 ms)
        p->table_single = bw >= 512;
        uint64_t ring_mb = 2048;
        if (const char* e = getenv("S2KIT_CUDA_FLY_RING_MB")) ring_mb = std::max(1L, atol(e));
        uint64_t cap = std::max<uint64_t>(biggest, (ring_mb << 20) / 512);
        p->fly_tiles = std::min<uint64_t>(cap, total_tiles);
        p->table_bytes = p->fly_tiles * 64 * sizeof(double);
        CK(cudaMalloc((void**)&p->d_table, p->table_bytes));
    }

    // ---- workspace
    size_t per_fn = sizeof(double) * ((size_t)2 * n * n + (size_t)n * 2 * bw);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    int chunk = std::max(1, max_batch);
    size_t budget = free_b / 3;  // leave room for the caller's data
    if ((size_t)chunk * per_fn > budget) chunk = (int)std::max<size_t>(1, budget / per_fn);
    p->chunk = chunk;
    CK(cudaMalloc((void**)&p->d_S, sizeof(double) * (size_t)chunk * 2 * n * n));
    CK(cudaMalloc((void**)&p->d_X, sizeof(double) * (size_t)chunk * n * 2 * bw));
    make_plane_tensor_map(p);
    CK(cudaStreamSynchronize(p->stream));
    apply_table_l2_policy(p);
    return 0;
}

extern "C" int s2kit_cuda_plan_create(s2kit_cuda_plan** out, int bw, int variant, int max_batch, int device) {
    return s2k_plan_create_impl(out, bw, variant, max_batch, device, 0, 1);
}

struct HostPipe;
static void host_pipe_destroy(void* hp);

extern "C" int s2kit_cuda_plan_destroy(s2kit_cuda_plan* p) {
    if (!p) return 0;
    DeviceGuard guard;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    s2k_shard_destroy(p);
    // tables and constants of a clone belong to the plan it was cloned from
    void* shared[] = {p->d_wv, p->d_sv, p->d_weights, p->d_sin, p->d_tw_n, p->d_tw_b, p->d_q_n, p->d_q_b, p->d_nodes,
                      p->d_seeds, p->d_rec, p->d_meta, p->d_rt_start, p->d_order_start, p->d_units,
                      p->d_sub_off,  p->d_sub_list, p->d_isub_off, p->d_isub_list, p->d_iq_off, p->d_iq_list};
    void* own[] = {p->d_S, p->d_T, p->d_X, p->d_coef, p->d_coef2, p->d_filt, p->d_stage};
    if (!p->shares_tables)
        for (void* q : shared)
            if (q) cudaFree(q);
    for (void* q : own)
        if (q) cudaFree(q);
    if (p->own_table && p->d_table) cudaFree(p->d_table);
    if (p->own_ckpt) {  // created on first use by whichever plan object generated tables first (kernels_table.cu)
        if (p->d_ckpt) cudaFree(p->d_ckpt);
        if (p->d_unit_first) cudaFree(p->d_unit_first);
    }
    for (auto& s : p->prof_slots) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    host_pipe_destroy(p->host_pipe);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    if (p->aux_stream) cudaStreamDestroy(p->aux_stream);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p->mu;
    delete p;
    return 0;
}

// A second plan object for the same bandwidth that SHARES the device tables and constants of `src` (read-only after
// creation) and owns its stream and workspaces: what concurrent callers need, one clone per thread.  `src` must
// outlive its clones.
extern "C" int s2kit_cuda_plan_clone(s2kit_cuda_plan** out, const s2kit_cuda_plan* src, int max_batch) {
    if (!out) return fail_msg("null output pointer");
    *out = nullptr;
    if (!src) return fail_msg("null plan");
    if (src->shard) return fail_msg("sharded plans cannot be cloned");
    DeviceGuard guard;
    CK(cudaSetDevice(src->device));
    s2kit_cuda_plan* p = new s2kit_cuda_plan(*src);
    p->mu = new std::mutex();
    p->shares_tables = true;
    p->own_table = false;
    p->own_ckpt = false;
    p->stream = nullptr;
    p->own_stream = true;
    p->host_pipe = nullptr;
    p->aux_stream = nullptr;
    p->ev_fork = p->ev_join = nullptr;
    p->d_S = p->d_T = p->d_X = p->d_coef = p->d_coef2 = p->d_filt = p->d_stage = nullptr;
    p->stage_doubles = 0;
    p->prof_slots.clear();
    p->prof_used = 0;
    p->profiling = false;
    p->tma_S_ok = false;
    auto build = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        const int n = p->n, bw = p->bw;
        if (p->variant == S2KIT_CUDA_FLY) {
            // the Fly scratch ring is written per call: private
            p->d_table = nullptr;
            CK(cudaMalloc((void**)&p->d_table, p->table_bytes));
            p->own_table = true;
        }
        p->chunk = std::max(1, std::min(max_batch, src->chunk));
        CK(cudaMalloc((void**)&p->d_S, sizeof(double) * (size_t)p->chunk * 2 * n * n));
        CK(cudaMalloc((void**)&p->d_X, sizeof(double) * (size_t)p->chunk * n * 2 * bw));
        make_plane_tensor_map(p);
        apply_table_l2_policy(p);
        return 0;
    };
    if (int rc = build()) {
        std::string keep = g_last_error;
        s2kit_cuda_plan_destroy(p);
        g_last_error = keep;
        return rc;
    }
    *out = p;
    return 0;
}

extern "C" int s2kit_cuda_plan_set_stream(s2kit_cuda_plan* p, void* stream) {
    if (!p) return fail_msg("null plan");
    std::lock_guard<std::mutex> lock(*p->mu);
    DeviceGuard guard;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    p->stream = (cudaStream_t)stream;
    p->own_stream = false;
    apply_table_l2_policy(p);
    return 0;
}
extern "C" void* s2kit_cuda_plan_stream(s2kit_cuda_plan* p) { return p ? (void*)p->stream : nullptr; }
extern "C" int s2kit_cuda_synchronize(s2kit_cuda_plan* p) {
    if (!p) return fail_msg("null plan");
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}
extern "C" int s2kit_cuda_plan_bw(const s2kit_cuda_plan* p) { return p ? p->bw : 0; }
extern "C" size_t s2kit_cuda_plan_table_bytes(const s2kit_cuda_plan* p) { return p ? p->table_bytes : 0; }
extern "C" size_t s2kit_cuda_plan_table_stream_bytes(const s2kit_cuda_plan* p) {
    if (!p) return 0;
    return p->variant == S2KIT_CUDA_MEMO ? p->table_tiles * 64 * sizeof(double) : 0;
}

// ------------------------------------------------------------------------------------------------ order groups
// Memo: one group covering [0, bw) on the resident table.  Fly: groups that fit the scratch ring; the table
// of each group is generated right before it is consumed.
struct OrderGroup {
    int lo, hi;
    uint64_t shift;
};

static std::vector<OrderGroup> order_groups(const s2kit_cuda_plan* p, int m_lo, int m_hi) {
    std::vector<OrderGroup> g;
    if (p->variant == S2KIT_CUDA_MEMO) {
        g.push_back({m_lo, m_hi, 0});
        return g;
    }
    int m = m_lo;
    while (m < m_hi) {
        int hi = m + 1;
        while (hi < m_hi && p->h_order_start[hi + 1] - p->h_order_start[m] <= p->fly_tiles) ++hi;
        g.push_back({m, hi, p->h_order_start[m]});
        m = hi;
    }
    return g;
}

static int ensure(double** ptr, size_t doubles) {
    if (*ptr) return 0;
    CK(cudaMalloc((void**)ptr, doubles * sizeof(double)));
    return 0;
}

// ------------------------------------------------------------------------------------------------ device paths
// One sub-batch of nf functions through the forward kernels, on p->stream, in the workspace slice starting at function
// slot w0 of d_S / d_X.
static int fst_sub(s2kit_cuda_plan* p, const double* rd, const double* id, double* rc, double* ic, int nf,
                   long data_stride, long coef_stride, int fmt, int w0) {
    const int bw = p->bw, n = p->n;
    const int nrows = (fmt == S2KIT_REAL) ? bw : 2 * bw - 1;
    double* dS = p->d_S + (size_t)w0 * 2 * n * n;
    double* dX = p->d_X + (size_t)w0 * n * 2 * bw;
    {
        // TMA variant of K1 available: keep the planes' latitudes in the DCT's own load order (PlaneView::lat_perm)
        s2k::PlaneView pv = s2k::default_view(p->n);
        pv.lat_perm = s2k::tma_planes_ok(p, w0 + nf);
        CK(s2k::launch_phi_fft_fwd(p, rd, id, data_stride, dS, nf, fmt, &pv));
        // batched: one persistent kernel does the DCTs and the contraction (kernels_uni.cu at bw = 256, kernels_pipe.cu)
        const bool pipe = s2k::fwd_pipe_supported(p, nf, fmt);
        const bool uni = pipe && s2k::fwd_pipe_fused() && s2k::fwd_uni_supported(p, nf, fmt);
        const bool pipe_fused = pipe && s2k::fwd_pipe_fused();
        if (!pipe_fused) CK(s2k::launch_dct_fwd(p, dS, dX, nf, 0, nrows, fmt, &pv));
        for (const OrderGroup& g : order_groups(p, 0, bw)) {
            if (p->variant == S2KIT_CUDA_FLY) CK(s2k::launch_table_gen(p, p->d_table, g.shift, g.lo, g.hi));
            if (uni)
                CK(s2k::launch_fwd_uni(p, p->d_table, g.shift, dS, rc, ic, coef_stride, nf, g.lo, g.hi, fmt,
                                       pv.lat_perm));
            else if (pipe_fused)
                CK(s2k::launch_fwd_pipe(p, p->d_table, g.shift, dS, rc, ic, coef_stride, nf, g.lo, g.hi, fmt,
                                        pv.lat_perm));
            else if (pipe)
                CK(s2k::launch_leg_fwd_stream(p, p->d_table, g.shift, dX, rc, ic, coef_stride, nf, g.lo, g.hi, fmt));
            else
                CK(s2k::launch_legendre_fwd(p, p->d_table, g.shift, dX, rc, ic, coef_stride, nf, g.lo, g.hi, fmt));
        }
    }
    return 0;
}

static int inv_fst_sub(s2kit_cuda_plan* p, const double* rc, const double* ic, double* rd, double* id, int nf,
                       long coef_stride, long data_stride, int fmt, int w0) {
    const int bw = p->bw, n = p->n;
    const int nrows = (fmt == S2KIT_REAL) ? bw : 2 * bw - 1;
    double* dS = p->d_S + (size_t)w0 * 2 * n * n;
    double* dX = p->d_X + (size_t)w0 * n * 2 * bw;
    s2k::PlaneView pv = s2k::default_view(p->n);
    pv.lat_perm = s2k::tma_planes_ok(p, w0 + nf);
    // batched at bw = 256 (opt-in): contraction and DCT-III in one persistent kernel, the cosine planes stay in shared memory
    const bool uni = s2k::inv_uni_supported(p, nf, fmt);
    // batched at bw = 256: the contraction as a persistent kernel (kernels_flow.cu)
    const bool flow = !uni && s2k::inv_flow_supported(p, nf, fmt);
    for (const OrderGroup& g : order_groups(p, 0, bw)) {
        const double* tt = p->variant == S2KIT_CUDA_FLY ? p->d_table : p->d_table_t;
        if (p->variant == S2KIT_CUDA_FLY) CK(s2k::launch_table_gen(p, p->d_table, g.shift, g.lo, g.hi, p->table_single ? 0 : 1));
        if (uni)
            CK(s2k::launch_inv_uni(p, tt, g.shift, rc, ic, coef_stride, dS, nf, g.lo, g.hi, fmt, pv.lat_perm));
        else if (flow)
            CK(s2k::launch_inv_flow(p, tt, g.shift, rc, ic, coef_stride, dX, nf, g.lo, g.hi, fmt));
        else
            CK(s2k::launch_legendre_inv(p, tt, g.shift, rc, ic, coef_stride, dX, nf, g.lo, g.hi, fmt));
    }
    if (!uni) CK(s2k::launch_dct_inv(p, dX, dS, nf, 0, nrows, fmt, &pv));
    CK(s2k::launch_phi_fft_inv(p, dS, rd, id, data_stride, nf, fmt, &pv));
    return 0;
}

// A chunk of a device call is cut into p->nsplit sub-batches that alternate between the plan's stream and an auxiliary
// one (S2KIT_CUDA_SPLIT, Memo plans): the HBM-bound kernels (longitude FFTs, DCTs) of one sub-batch can then run beside
// the FP64-bound contraction of another instead of each kernel having the machine to itself.
template <typename F>
static int run_split(s2kit_cuda_plan* p, int nf, F sub) {
    int ns = (p->variant == S2KIT_CUDA_MEMO) ? p->nsplit : 1;
    if (ns > 1 && nf < 64 * ns) ns = 1;
    if (ns <= 1) return sub(0, nf);
    if (!p->aux_stream) {
        CK(cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    const int per = ((nf + ns - 1) / ns + 7) / 8 * 8;  // whole 32-column panels in both data formats
    cudaStream_t main_stream = p->stream;
    CK(cudaEventRecord(p->ev_fork, main_stream));
    CK(cudaStreamWaitEvent(p->aux_stream, p->ev_fork, 0));
    int rc = 0, i = 0;
    for (int f0 = 0; f0 < nf && !rc; f0 += per, ++i) {
        p->stream = (i & 1) ? p->aux_stream : main_stream;
        rc = sub(f0, std::min(per, nf - f0));
    }
    p->stream = main_stream;
    if (rc) return rc;
    CK(cudaEventRecord(p->ev_join, p->aux_stream));
    CK(cudaStreamWaitEvent(main_stream, p->ev_join, 0));
    return 0;
}

static int fst_device(s2kit_cuda_plan* p, const double* rdata, const double* idata, double* rco, double* ico,
                      int batch, long data_stride, long coef_stride, int fmt) {
    for (int c0 = 0; c0 < batch; c0 += p->chunk) {
        const int nf = std::min(p->chunk, batch - c0);
        const int rc = run_split(p, nf, [&](int f0, int n) {
            const long f = c0 + f0;
            return fst_sub(p, rdata + f * data_stride, idata + f * data_stride, rco + f * coef_stride, ico + f * coef_stride, n,
                           data_stride, coef_stride, fmt, f0);
        });
        if (rc) return rc;
    }
    return 0;
}

static int inv_fst_device(s2kit_cuda_plan* p, const double* rco, const double* ico, double* rdata, double* idata,
                          int batch, long coef_stride, long data_stride, int fmt) {
    for (int c0 = 0; c0 < batch; c0 += p->chunk) {
        const int nf = std::min(p->chunk, batch - c0);
        const int rc = run_split(p, nf, [&](int f0, int n) {
            const long f = c0 + f0;
            return inv_fst_sub(p, rco + f * coef_stride, ico + f * coef_stride, rdata + f * data_stride, idata + f * data_stride,
                               n, coef_stride, data_stride, fmt, f0);
        });
        if (rc) return rc;
    }
    return 0;
}

static int fzt_device(s2kit_cuda_plan* p, const double* rdata, const double* idata, double* rres, double* ires,
                      int batch, long data_stride, long res_stride, int fmt) {
    const int bw = p->bw;
    for (int c0 = 0; c0 < batch; c0 += p->chunk) {
        int nf = std::min(p->chunk, batch - c0);
        CK(s2k::launch_zonal_rowsum(p, rdata + (long)c0 * data_stride, idata + (long)c0 * data_stride, data_stride,
                                    p->d_S, nf));
        CK(s2k::launch_dct_fwd(p, p->d_S, p->d_X, nf, 0, 1, S2KIT_COMPLEX));
        double* rr = rres + (long)c0 * res_stride;
        double* ir = ires + (long)c0 * res_stride;
        for (const OrderGroup& g : order_groups(p, 0, 1)) {
            if (p->variant == S2KIT_CUDA_FLY) CK(s2k::launch_table_gen(p, p->d_table, g.shift, g.lo, g.hi));
            // order 0 lands at positions l of the output (IndexOfHarmonicCoeff(0,l) = l)
            CK(s2k::launch_legendre_fwd(p, p->d_table, g.shift, p->d_X, rr, ir, res_stride, nf, 0, 1, S2KIT_COMPLEX));
        }
        if (fmt == S2KIT_REAL)  // FST_semi_memo.c:405-406 (bw entries; the reference clears 2bw)
            CK(cudaMemset2DAsync(ir, res_stride * sizeof(double), 0, bw * sizeof(double), nf, p->stream));
    }
    return 0;
}

static int conv_device(s2kit_cuda_plan* p, const double* rdata, const double* idata, const double* rfilter,
                       const double* ifilter, double* rres, double* ires, int batch, long data_stride,
                       long filter_stride) {
    const int bw = p->bw;
    const long cs = (long)bw * bw;
    if (ensure(&p->d_coef, (size_t)p->chunk * 2 * cs)) return 1;
    if (ensure(&p->d_coef2, (size_t)p->chunk * 2 * cs)) return 1;
    if (ensure(&p->d_filt, (size_t)p->chunk * 2 * bw)) return 1;
    double* fr = p->d_coef;
    double* fi = p->d_coef + (size_t)p->chunk * cs;
    double* tr = p->d_coef2;
    double* ti = p->d_coef2 + (size_t)p->chunk * cs;
    double* hr = p->d_filt;
    double* hi = p->d_filt + (size_t)p->chunk * bw;
    for (int c0 = 0; c0 < batch; c0 += p->chunk) {
        int nf = std::min(p->chunk, batch - c0);
        // ConvOn2SphereSemiMemo, FST_semi_memo.c:502-508: everything in REAL format, cutoff = bw
        if (filter_stride != 0 || c0 == 0) {
            int nfilt = filter_stride ? nf : 1;
            if (fzt_device(p, rfilter + (long)c0 * filter_stride, ifilter + (long)c0 * filter_stride, hr, hi, nfilt,
                           filter_stride ? filter_stride : (long)p->n * p->n, bw, S2KIT_REAL))
                return 1;
        }
        if (fst_device(p, rdata + (long)c0 * data_stride, idata + (long)c0 * data_stride, fr, fi, nf, data_stride, cs,
                       S2KIT_REAL))
            return 1;
        CK(s2k::launch_spectral_mul(p, fr, fi, cs, hr, hi, filter_stride ? bw : 0, tr, ti, cs, nf));
        if (inv_fst_device(p, tr, ti, rres + (long)c0 * data_stride, ires + (long)c0 * data_stride, nf, cs,
                           data_stride, S2KIT_REAL))
            return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ host staging
static cudaError_t copy_in(double* dst, long dst_pitch, const double* src, long src_stride, long len, int count,
                           cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, dst_pitch * sizeof(double), src, src_stride * sizeof(double), len * sizeof(double),
                             count, cudaMemcpyHostToDevice, s);
}
static cudaError_t copy_out(double* dst, long dst_stride, const double* src, long src_pitch, long len, int count,
                            cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, dst_stride * sizeof(double), src, src_pitch * sizeof(double), len * sizeof(double),
                             count, cudaMemcpyDeviceToHost, s);
}

static int check_common(s2kit_cuda_plan* p, int batch, int fmt) {
    if (!p) return fail_msg("null plan");
    if (batch < 0) return fail_msg("negative batch");
    if (fmt != S2KIT_COMPLEX && fmt != S2KIT_REAL) return fail_msg("unknown data format");
    if (cudaSetDevice(p->device) != cudaSuccess) return fail_msg("cudaSetDevice failed");
    return 0;
}

// Every public entry point runs under the plan's lock with the plan's device current, and puts the caller's device
// back on return.
#define S2K_ENTER(p, batch, fmt)                       \
    if (!(p)) return fail_msg("null plan");            \
    DeviceGuard guard__;                               \
    std::lock_guard<std::mutex> lock__(*(p)->mu);      \
    if (int r__ = check_common((p), (batch), (fmt))) return r__

// Host-pointer calls: a three-stage pipeline over sub-chunks of the batch -- H2D of sub-chunk i+1, the kernels of
// sub-chunk i and D2H of sub-chunk i-1 run concurrently on three streams with double-buffered device staging, so
// PCIe runs full duplex and the GPU work hides behind the copies.  (Pinned host memory is needed for true overlap;
// pageable memory still works, the copies just serialise.)  The copy streams and events live as long as the plan.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    bool ok = false;
    HostPipe() {
        ok = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) == cudaSuccess;
        for (int b = 0; b < 2 && ok; ++b)
            ok = cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&comp_done[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming) == cudaSuccess;
    }
    ~HostPipe() {
        for (int b = 0; b < 2; ++b) {
            if (in_done[b]) cudaEventDestroy(in_done[b]);
            if (comp_done[b]) cudaEventDestroy(comp_done[b]);
            if (out_done[b]) cudaEventDestroy(out_done[b]);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    }
};
static void host_pipe_destroy(void* hp) { delete reinterpret_cast<HostPipe*>(hp); }

// one strided host array of a batched call: `len` doubles per function, functions `stride` doubles apart
// (stride 0 = one array shared by the whole batch: copied once per sub-chunk slot)
struct HostSpan {
    const double* in = nullptr;
    double* out = nullptr;
    long stride = 0, len = 0;
};

// ins / outs: the host arrays of the call.  compute(nf, din[], dout[]) enqueues the kernels for nf functions on
// p->stream; device arrays are dense (function f at f * len).
template <typename F>
static int host_pipeline(s2kit_cuda_plan* p, int batch, const HostSpan* ins, int nin, const HostSpan* outs, int nout,
                         F compute) {
    const int sub = std::max(1, std::min(p->chunk, 32));
    size_t per_slot = 0;
    for (int i = 0; i < nin; ++i) per_slot += (size_t)(ins[i].stride ? sub : 1) * ins[i].len;
    for (int i = 0; i < nout; ++i) per_slot += (size_t)sub * outs[i].len;
    const size_t need = 2 * per_slot;
    if (p->stage_doubles < need) {
        if (p->d_stage) cudaFree(p->d_stage);
        p->d_stage = nullptr;
        p->stage_doubles = 0;
        CK(cudaMalloc((void**)&p->d_stage, need * sizeof(double)));
        p->stage_doubles = need;
    }
    if (!p->host_pipe) p->host_pipe = new HostPipe();
    HostPipe& hp = *reinterpret_cast<HostPipe*>(p->host_pipe);
    if (!hp.ok) return fail_msg("could not create the copy streams");
    double* din[2][4];
    double* dout[2][4];
    {
        double* q = p->d_stage;
        for (int b = 0; b < 2; ++b) {
            for (int i = 0; i < nin; ++i) {
                din[b][i] = q;
                q += (size_t)(ins[i].stride ? sub : 1) * ins[i].len;
            }
            for (int i = 0; i < nout; ++i) {
                dout[b][i] = q;
                q += (size_t)sub * outs[i].len;
            }
        }
    }
    int i = 0;
    for (int c0 = 0; c0 < batch; c0 += sub, ++i) {
        const int b = i & 1, nf = std::min(sub, batch - c0);
        // H2D: the input buffers of slot b are free once the kernels of sub-chunk i-2 are done
        if (i >= 2) CK(cudaStreamWaitEvent(hp.s_in, hp.comp_done[b], 0));
        for (int k = 0; k < nin; ++k) {
            if (ins[k].stride)
                CK(copy_in(din[b][k], ins[k].len, ins[k].in + (long)c0 * ins[k].stride, ins[k].stride, ins[k].len, nf,
                           hp.s_in));
            else if (i < 2)  // shared array: once per slot
                CK(copy_in(din[b][k], ins[k].len, ins[k].in, ins[k].len, ins[k].len, 1, hp.s_in));
        }
        CK(cudaEventRecord(hp.in_done[b], hp.s_in));
        // kernels: need the inputs, and the output buffers of slot b drained (sub-chunk i-2)
        CK(cudaStreamWaitEvent(p->stream, hp.in_done[b], 0));
        if (i >= 2) CK(cudaStreamWaitEvent(p->stream, hp.out_done[b], 0));
        if (int rc = compute(nf, din[b], dout[b])) return rc;
        CK(cudaEventRecord(hp.comp_done[b], p->stream));
        // D2H
        CK(cudaStreamWaitEvent(hp.s_out, hp.comp_done[b], 0));
        for (int k = 0; k < nout; ++k)
            CK(copy_out(outs[k].out + (long)c0 * outs[k].stride, outs[k].stride, dout[b][k], outs[k].len, outs[k].len,
                        nf, hp.s_out));
        CK(cudaEventRecord(hp.out_done[b], hp.s_out));
    }
    CK(cudaStreamSynchronize(hp.s_out));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

static HostSpan span_in(const double* ptr, long stride, long len) {
    HostSpan s;
    s.in = ptr;
    s.stride = stride;
    s.len = len;
    return s;
}
static HostSpan span_out(double* ptr, long stride, long len) {
    HostSpan s;
    s.out = ptr;
    s.stride = stride;
    s.len = len;
    return s;
}

extern "C" int s2kit_cuda_fst(s2kit_cuda_plan* p, const double* rdata, const double* idata, double* rco, double* ico,
                              int batch, long data_stride, long coef_stride, int fmt, int where) {
    S2K_ENTER(p, batch, fmt);
    if (batch == 0) return 0;
    if (where == S2KIT_CUDA_DEVICE) return fst_device(p, rdata, idata, rco, ico, batch, data_stride, coef_stride, fmt);
    const long gs = (long)p->n * p->n, cs = (long)p->bw * p->bw;
    HostSpan ins[2] = {span_in(rdata, data_stride, gs), span_in(idata, data_stride, gs)};
    HostSpan outs[2] = {span_out(rco, coef_stride, cs), span_out(ico, coef_stride, cs)};
    return host_pipeline(p, batch, ins, 2, outs, 2, [&](int nf, double** di, double** dout) {
        return fst_device(p, di[0], di[1], dout[0], dout[1], nf, gs, cs, fmt);
    });
}

extern "C" int s2kit_cuda_inv_fst(s2kit_cuda_plan* p, const double* rco, const double* ico, double* rdata,
                                  double* idata, int batch, long coef_stride, long data_stride, int fmt, int where) {
    S2K_ENTER(p, batch, fmt);
    if (batch == 0) return 0;
    if (where == S2KIT_CUDA_DEVICE)
        return inv_fst_device(p, rco, ico, rdata, idata, batch, coef_stride, data_stride, fmt);
    const long gs = (long)p->n * p->n, cs = (long)p->bw * p->bw;
    HostSpan ins[2] = {span_in(rco, coef_stride, cs), span_in(ico, coef_stride, cs)};
    HostSpan outs[2] = {span_out(rdata, data_stride, gs), span_out(idata, data_stride, gs)};
    return host_pipeline(p, batch, ins, 2, outs, 2, [&](int nf, double** di, double** dout) {
        return inv_fst_device(p, di[0], di[1], dout[0], dout[1], nf, cs, gs, fmt);
    });
}

extern "C" int s2kit_cuda_fzt(s2kit_cuda_plan* p, const double* rdata, const double* idata, double* rres, double* ires,
                              int batch, long data_stride, long res_stride, int fmt, int where) {
    S2K_ENTER(p, batch, fmt);
    if (batch == 0) return 0;
    if (where == S2KIT_CUDA_DEVICE) return fzt_device(p, rdata, idata, rres, ires, batch, data_stride, res_stride, fmt);
    const long gs = (long)p->n * p->n;
    const int bw = p->bw;
    HostSpan ins[2] = {span_in(rdata, data_stride, gs), span_in(idata, data_stride, gs)};
    HostSpan outs[2] = {span_out(rres, res_stride, bw), span_out(ires, res_stride, bw)};
    return host_pipeline(p, batch, ins, 2, outs, 2, [&](int nf, double** di, double** dout) {
        return fzt_device(p, di[0], di[1], dout[0], dout[1], nf, gs, bw, fmt);
    });
}

extern "C" int s2kit_cuda_conv(s2kit_cuda_plan* p, const double* rdata, const double* idata, const double* rfilter,
                               const double* ifilter, double* rres, double* ires, int batch, long data_stride,
                               long filter_stride, int where) {
    S2K_ENTER(p, batch, S2KIT_REAL);
    if (batch == 0) return 0;
    if (where == S2KIT_CUDA_DEVICE)
        return conv_device(p, rdata, idata, rfilter, ifilter, rres, ires, batch, data_stride, filter_stride);
    // host pointers: signal and filter grids stream in, result grids stream out, through the same three-stage
    // pipeline and the plan's persistent staging as the transforms (no allocation per call)
    const long gs = (long)p->n * p->n;
    HostSpan ins[4] = {span_in(rdata, data_stride, gs), span_in(idata, data_stride, gs),
                       span_in(rfilter, filter_stride, gs), span_in(ifilter, filter_stride, gs)};
    HostSpan outs[2] = {span_out(rres, data_stride, gs), span_out(ires, data_stride, gs)};
    return host_pipeline(p, batch, ins, 4, outs, 2, [&](int nf, double** di, double** dout) {
        return conv_device(p, di[0], di[1], di[2], di[3], dout[0], dout[1], nf, gs, filter_stride ? gs : 0);
    });
}

extern "C" int s2kit_cuda_trans_mult(s2kit_cuda_plan* p, const double* rd, const double* id, const double* rf,
                                     const double* ifl, double* rres, double* ires, int batch, long coef_stride,
                                     int where) {
    S2K_ENTER(p, batch, S2KIT_COMPLEX);
    if (batch == 0) return 0;
    const int bw = p->bw;
    const long cs = (long)bw * bw;
    if (where == S2KIT_CUDA_DEVICE) {
        CK(s2k::launch_spectral_mul(p, rd, id, coef_stride, rf, ifl, bw, rres, ires, coef_stride, batch));
        return 0;
    }
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(double) * (4 * (size_t)cs + 2 * bw)));
    double *a = d, *b = d + cs, *c = d + 2 * cs, *e2 = d + 3 * cs, *hr = d + 4 * cs, *hi = hr + bw;
    int rc = 0;
    for (int f = 0; f < batch && !rc; ++f) {
        cudaMemcpyAsync(a, rd + (long)f * coef_stride, cs * 8, cudaMemcpyHostToDevice, p->stream);
        cudaMemcpyAsync(b, id + (long)f * coef_stride, cs * 8, cudaMemcpyHostToDevice, p->stream);
        cudaMemcpyAsync(hr, rf + (long)f * bw, bw * 8, cudaMemcpyHostToDevice, p->stream);
        cudaMemcpyAsync(hi, ifl + (long)f * bw, bw * 8, cudaMemcpyHostToDevice, p->stream);
        cudaError_t e = s2k::launch_spectral_mul(p, a, b, cs, hr, hi, bw, c, e2, cs, 1);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(rres + (long)f * coef_stride, c, cs * 8, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ires + (long)f * coef_stride, e2, cs * 8, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) rc = fail("trans_mult", e);
    }
    cudaFree(d);
    return rc;
}

// ------------------------------------------------------------------------------------------------ 1-D entry points
// DLTSemi / InvDLTSemi on ncols independent real columns of one order: columns are fed through the batched
// kernels as the real parts of ncols "functions" whose imaginary parts are zero.
extern "C" int s2kit_cuda_dlt_semi(s2kit_cuda_plan* p, const double* data, int m, double* result, int ncols, int where) {
    S2K_ENTER(p, ncols, S2KIT_COMPLEX);
    if (m < 0 || m >= p->bw) return fail_msg("order out of range");
    if (ncols == 0) return 0;
    const int bw = p->bw, n = p->n;
    const long cs = (long)bw * bw;
    double *dcoef = nullptr;
    CK(cudaMalloc((void**)&dcoef, sizeof(double) * 2 * (size_t)cs));
    int rc = 0;
    for (int c = 0; c < ncols && !rc; ++c) {
        // place the column as order row m of function 0 (real part), imaginary part zero
        cudaError_t e = cudaMemsetAsync(p->d_S, 0, sizeof(double) * 2 * (size_t)n * n, p->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(p->d_S + (size_t)m * n, data + (long)c * n, n * 8,
                                where == S2KIT_CUDA_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                p->stream);
        if (e == cudaSuccess) e = s2k::launch_dct_fwd(p, p->d_S, p->d_X, 1, m, m + 1, S2KIT_COMPLEX);
        for (const OrderGroup& g : order_groups(p, m, m + 1)) {
            if (e == cudaSuccess && p->variant == S2KIT_CUDA_FLY) e = s2k::launch_table_gen(p, p->d_table, g.shift, g.lo, g.hi);
            if (e == cudaSuccess)
                e = s2k::launch_legendre_fwd(p, p->d_table, g.shift, p->d_X, dcoef, dcoef + cs, cs, 1, m, m + 1,
                                             S2KIT_REAL);
        }
        long at = (long)m * bw - ((long)m * (m - 1)) / 2;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(result + (long)c * (bw - m), dcoef + at, (bw - m) * 8,
                                where == S2KIT_CUDA_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) rc = fail("dlt_semi", e);
    }
    cudaFree(dcoef);
    return rc;
}

extern "C" int s2kit_cuda_inv_dlt_semi(s2kit_cuda_plan* p, const double* coeffs, int m, double* result, int ncols,
                                       int where) {
    S2K_ENTER(p, ncols, S2KIT_COMPLEX);
    if (m < 0 || m >= p->bw) return fail_msg("order out of range");
    if (ncols == 0) return 0;
    const int bw = p->bw, n = p->n;
    const long cs = (long)bw * bw;
    double* dcoef = nullptr;
    CK(cudaMalloc((void**)&dcoef, sizeof(double) * 2 * (size_t)cs));
    int rc = 0;
    const double undo = sqrt(2.0 * M_PI);  // K5 folds in the 1/sqrt(2 pi) of InvFST; DLT alone has none
    for (int c = 0; c < ncols && !rc; ++c) {
        cudaError_t e = cudaMemsetAsync(dcoef, 0, sizeof(double) * 2 * (size_t)cs, p->stream);
        long at = (long)m * bw - ((long)m * (m - 1)) / 2;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dcoef + at, coeffs + (long)c * (bw - m), (bw - m) * 8,
                                where == S2KIT_CUDA_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                p->stream);
        for (const OrderGroup& g : order_groups(p, m, m + 1)) {
            if (e == cudaSuccess && p->variant == S2KIT_CUDA_FLY) e = s2k::launch_table_gen(p, p->d_table, g.shift, g.lo, g.hi, p->table_single ? 0 : 1);
            if (e == cudaSuccess)
                e = s2k::launch_legendre_inv(p, p->variant == S2KIT_CUDA_FLY ? p->d_table : p->d_table_t, g.shift, dcoef, dcoef + cs, cs, p->d_X, 1, m, m + 1,
                                             S2KIT_REAL);
        }
        if (e == cudaSuccess) e = s2k::launch_dct_inv(p, p->d_X, p->d_S, 1, m, m + 1, S2KIT_COMPLEX);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) {
            rc = fail("inv_dlt_semi", e);
            break;
        }
        std::vector<double> row(n);
        e = cudaMemcpy(row.data(), p->d_S + (size_t)m * n, n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            rc = fail("inv_dlt_semi D2H", e);
            break;
        }
        for (int j = 0; j < n; ++j) row[j] *= undo;
        e = cudaMemcpy(result + (long)c * n, row.data(), n * 8,
                       where == S2KIT_CUDA_DEVICE ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost);
        if (e != cudaSuccess) rc = fail("inv_dlt_semi out", e);
    }
    cudaFree(dcoef);
    return rc;
}

// naive algorithm (naive.c): no plan, the table comes from the caller.  Host pointers are staged through the device.
static int naive_common(const double* vec, long vec_len, const double* weights, const double* pml, double* result,
                        long res_len, int bw, int m, int where, bool forward) {
    if (bw < 1 || m < 0 || m >= bw) return fail_msg("order out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail_msg("no CUDA device available: s2kit_cuda has no CPU fallback");
    const int size = 2 * bw, rows = bw - m;
    const size_t tab = (size_t)size * rows;
    const double *dv = vec, *dw = weights, *dt = pml;
    double* dr = result;
    double* stage = nullptr;
    if (where != S2KIT_CUDA_DEVICE) {
        CK(cudaMalloc((void**)&stage, sizeof(double) * (tab + vec_len + res_len + size)));
        double *sv = stage + tab, *sw = sv + vec_len, *sr = sw + size;
        cudaError_t e = cudaMemcpy(stage, pml, tab * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(sv, vec, vec_len * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && weights) e = cudaMemcpy(sw, weights, (size_t)size * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(stage);
            return fail("naive DLT H2D", e);
        }
        dt = stage, dv = sv, dw = sw, dr = sr;
    }
    cudaError_t e = forward ? s2k::launch_naive_dlt(dv, dw, dt, dr, size, rows, 0)
                            : s2k::launch_naive_inv_dlt(dv, dt, dr, size, rows, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e == cudaSuccess && stage) e = cudaMemcpy(result, dr, res_len * 8, cudaMemcpyDeviceToHost);
    if (stage) cudaFree(stage);
    if (e != cudaSuccess) return fail("naive DLT", e);
    return 0;
}

extern "C" int s2kit_cuda_dlt_naive(const double* data, int bw, int m, const double* weights, double* result,
                                    const double* pml_table, int where) {
    if (!data || !weights || !result || !pml_table) return fail_msg("null pointer");
    return naive_common(data, 2L * bw, weights, pml_table, result, (long)bw - m, bw, m, where, true);
}

extern "C" int s2kit_cuda_inv_dlt_naive(const double* coeffs, int bw, int m, double* result, const double* pml_table,
                                        int where) {
    if (!coeffs || !result || !pml_table) return fail_msg("null pointer");
    return naive_common(coeffs, (long)bw - m, nullptr, pml_table, result, 2L * bw, bw, m, where, false);
}

// ------------------------------------------------------------------------------------------------ tables
static int table_to_host(s2kit_cuda_plan* p, const double* table, uint64_t shift, int m, double* host_out) {
    int size = 0;
    for (int l = m; l < p->bw; ++l) size += row_size(m, l);
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(double) * (size_t)size));
    cudaError_t e = s2k::launch_table_unpack(p, table, shift, m, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d, sizeof(double) * (size_t)size, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail("table export", e);
    return 0;
}

static int table_generate_locked(s2kit_cuda_plan* p, int m, double* host_out);

extern "C" int s2kit_cuda_table_export(s2kit_cuda_plan* p, int m, double* host_out) {
    S2K_ENTER(p, 0, S2KIT_COMPLEX);
    if (m < 0 || m >= p->bw) return fail_msg("order out of range");
    if (p->variant != S2KIT_CUDA_MEMO) return table_generate_locked(p, m, host_out);
    if (p->h_order_start[m + 1] == p->h_order_start[m]) return fail_msg("order not resident on this rank");
    return table_to_host(p, p->d_table, 0, m, host_out);
}

extern "C" int s2kit_cuda_table_generate(s2kit_cuda_plan* p, int m, double* host_out) {
    S2K_ENTER(p, 0, S2KIT_COMPLEX);
    return table_generate_locked(p, m, host_out);
}

static int table_generate_locked(s2kit_cuda_plan* p, int m, double* host_out) {
    if (m < 0 || m >= p->bw) return fail_msg("order out of range");
    uint64_t tiles = 0;
    {
        // size of order m in a full layout (independent of ownership)
        for (int par = 0; par < 2; ++par) {
            const s2k::BlockMeta& mb = p->h_meta[2 * m + par];
            for (int rt = 0; rt < mb.nrt; ++rt)
                tiles += (uint64_t)((mb.len0 + std::min(8 * rt + 7, mb.rows - 1) + 7) >> 3);
        }
    }
    if (p->h_order_start[m + 1] - p->h_order_start[m] != tiles)
        return fail_msg("order not owned by this rank");
    double* scratch = nullptr;
    CK(cudaMalloc((void**)&scratch, tiles * 64 * sizeof(double)));
    uint64_t shift = p->h_order_start[m];
    cudaError_t e = s2k::launch_table_gen(p, scratch, shift, m, m + 1);
    int rc = 0;
    if (e != cudaSuccess)
        rc = fail("table generate", e);
    else
        rc = table_to_host(p, scratch, shift, m, host_out);
    cudaFree(scratch);
    return rc;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int s2kit_cuda_measure_fp64_peak(int device, double* fma_tflops, double* dmma_tflops) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("no CUDA device available");
    CK(cudaSetDevice(device));
    double a = 0, b = 0;
    CK(s2k::measure_fp64(&a, &b));
    if (fma_tflops) *fma_tflops = a;
    if (dmma_tflops) *dmma_tflops = b;
    return 0;
}

extern "C" int s2kit_cuda_measure_copy_bw(int device, size_t bytes, double* gbs) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("no CUDA device available");
    CK(cudaSetDevice(device));
    double v = 0;
    CK(s2k::measure_copy(bytes, &v));
    if (gbs) *gbs = v;
    return 0;
}

extern "C" void* s2kit_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void s2kit_cuda_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
