// kernels_flow.cu -- K4 (inverse Legendre contraction) at bw = 256, batched, as ONE persistent kernel.
//
//   V[col, k] = sum_l c[col, l] T_m[l, k]   on DMMA
//   (the contraction loop of InvDLTSemi, src/legendre_transform/seminaive.c:76-96, inside the m-loops of InvFSTSemiMemo,
//    src/FST_semi_memo.c:262-342)
//
// What the profile of k_legendre_inv<32,32> (one CTA per (order, 32 columns), 3 CTAs per SM; profiles/r2_dram_traffic.json,
// ncu source view) showed: DMMA pipe 65 % busy of which 0.79 useful; 21 % of the warp samples in the CTA prologue (the
// coefficient panel is staged before any DMMA can issue and a CTA only has ~200 DMMAs per warp to amortise it), 7 % in the
// epilogue, 7 % waiting for table tiles on the two-deep register ring; the LSU data pipe 70 % busy (one 64-bit shared
// load per DMMA for the coefficient fragments, table tiles through LDG at 64 bytes per wavefront).  This kernel keeps the
// work items -- (order m, 32 columns) -- and changes the rest:
//   * persistent, one 16-warp CTA per SM, two coefficient panels: the panel of item k + 1 is staged by STAGE TASKS that
//     run beside the DMMA units of item k, taken by the same warps from the same four per-sub-partition ticket queues
//     (the machinery of k_fwd_uni, kernels_uni.cu: no CTA-wide barrier, release/acquire counters per panel buffer);
//   * a DMMA unit is FOUR adjacent column tiles of one parity block against the 32 panel columns: 16 accumulator
//     fragments, 32 DMMAs per row-tile step, and every coefficient fragment feeds four DMMAs instead of two -- half the
//     shared-memory traffic per DMMA;
//   * the four table tiles of a step are contiguous in memory (row-tile-major layout, 2 KB): ONE cp.async.bulk copy (TMA,
//     async proxy, no LSU wavefronts on the way in) by one elected lane per 32 DMMAs, completion counted in bytes on an
//     mbarrier, two stages per warp; the lanes read their B fragments with one 128-bit shared load per tile.
#include <stdlib.h>

#include "s2k_legendre.cuh"

namespace s2k {

constexpr int FLOW_NC = 32;
constexpr int FLOW_WARPS = 16;
constexpr int FLOW_THREADS = FLOW_WARPS * 32;
constexpr int FLOW_STAGES = 2;  // ring stages per warp, four tiles (2 KB) each
constexpr int FLOW_Q = 4;       // column tiles per unit

struct FlowInvArgs {
    const double* table;  // B-fragment-ordered tiles
    const uint64_t* order_start;
    uint64_t table_shift;
    const int* sub_off;              // [4 bw + 1]: queue s of order m = sub_list[sub_off[4m+s] .. sub_off[4m+s+1])
    const unsigned short* sub_list;  // packed units: parity | quad << 1  (column tiles 4 quad .. 4 quad + 3)
    int nlist;
    const double* rco;
    const double* ico;
    long coef_stride;
    double* V;  // cosine planes [f][order row][part][cos_slot]
    int nfun, m_lo, norders, ncoltiles, real_fmt;
    int sleep_ns;
};

__device__ __forceinline__ void flow_signal(int* ctr) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(ctr)))
                 : "memory");
}
__device__ __forceinline__ void flow_wait_ge(const int* ctr, int need, int sleep_ns) {
    if (need <= 0) return;
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(ctr));
    unsigned spins = 0;
    for (;;) {
        int v;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v >= need) return;
        if (sleep_ns) __nanosleep(sleep_ns);
        if (++spins > (1u << 26)) __trap();  // a lost signal must not hang the device
    }
}

__host__ __device__ constexpr size_t flow_meta_bytes(int B, int nlist) {
    return sizeof(int) * (4 * B + 4) + sizeof(unsigned) * (B + 4) + sizeof(unsigned short) * ((nlist + 7) & ~7);
}

__global__ void __launch_bounds__(FLOW_THREADS, 1) k_inv_flow(const FlowInvArgs a) {
    constexpr int N = 512, B = 256, NC = FLOW_NC, CS = 132, PS = NC * CS + 8, PANEL = 2 * PS;
    extern __shared__ __align__(128) double smem[];
    double* panels = smem;                                                     // [2][PANEL]: [parity][column][row r]
    double2* rings = reinterpret_cast<double2*>(smem + 2 * PANEL);             // [WARPS][STAGES][4 tiles][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rings + FLOW_WARPS * FLOW_STAGES * FLOW_Q * 32);  // [WARPS][STAGES]
    int* tick = reinterpret_cast<int*>(bars + FLOW_WARPS * FLOW_STAGES);        // [4] ticket counters
    int* done = tick + 4;  // [0..1] stage_done, [2..3] dmma_done, per panel-buffer parity
    int* qoff = tick + 16;
    unsigned* ost = reinterpret_cast<unsigned*>(qoff + 4 * B + 4);
    unsigned short* qlist = reinterpret_cast<unsigned short*>(ost + B + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const int sq = warp & 3;  // this warp's SM sub-partition = its task queue
    const int cols_per_fn = a.real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int nitems = a.norders * a.ncoltiles;
    const int nk = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    double2* ring = rings + warp * FLOW_STAGES * FLOW_Q * 32;
    const unsigned bar0 = static_cast<unsigned>(__cvta_generic_to_shared(bars + warp * FLOW_STAGES));
    unsigned phases = 0;

    if (tid < 8) tick[tid] = 0;
    for (int i = tid; i <= 4 * B; i += FLOW_THREADS) qoff[i] = a.sub_off[i];
    for (int i = tid; i <= B; i += FLOW_THREADS)
        ost[i] = (i >= a.m_lo && i <= a.m_lo + a.norders) ? (unsigned)(a.order_start[i] - a.table_shift) : 0u;
    for (int i = tid; i < a.nlist; i += FLOW_THREADS) qlist[i] = a.sub_list[i];
    // stale panel entries meet zero table padding: they must be finite
    for (int i = tid; i < 2 * PANEL; i += FLOW_THREADS) panels[i] = 0.0;
    if (lane == 0) {
        for (int s = 0; s < FLOW_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    auto item_of = [&](int t, int& m, int& f0) {
        const int oi = t / a.ncoltiles, x = t - oi * a.ncoltiles;
        m = a.m_lo + oi;
        f0 = x * NF;
    };
    auto order_prefetch = [&](int m) {
        prefetch_order_l2(a.table + (uint64_t)ost[m] * 64, a.order_start[m + 1] - a.order_start[m], lane, 32, 1u << 20);
    };
    if (warp == 0 && nk > 0) order_prefetch(a.m_lo + (int)blockIdx.x / a.ncoltiles);

    // stage task `q` of item `item`: panel columns 2q (real part) and 2q + 1 (imaginary part) of buffer `buf`,
    // de-interleaved by the parity of l - m
    auto stage_task = [&](int item, int q, double* buf) {
        int m, f0;
        item_of(item, m, f0);
        const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        if (f >= a.nfun || (sgn && m == 0)) return;  // dead columns: whatever they hold is finite and never stored
        const int cnt = B - m;
        const long at = (long)f * a.coef_stride + (sgn ? coef_base(-m, B) : coef_base(m, B));
        const double* sr = a.rco + at;
        const double* si = a.ico + at;
        double vr[8], vi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int o = lane + 32 * e;
            vr[e] = vi[e] = 0.0;
            if (o < cnt) {
                vr[e] = __ldg(sr + o);
                vi[e] = __ldg(si + o);
            }
        }
        // element o = l - m goes to parity o & 1 (= lane & 1), row o >> 1; the (at most 8) rows between the order's last
        // degree and the end of its last row tile are cleared: they meet zero table padding, not table values
        double* dst = buf + (2 * q) * CS + (lane & 1) * PS + (lane >> 1);
        const int lim = cnt + 16 < 8 * 32 ? cnt + 16 : 8 * 32;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int o = lane + 32 * e;
            if (o < lim) {
                dst[16 * e] = vr[e];
                dst[16 * e + CS] = vi[e];
            }
        }
    };

    // the coefficient runs the stage tasks of `item` will read: into L2 (one run per lane: 16 pairs x re / im)
    auto prefetch_cols = [&](int item) {
        int m, f0;
        item_of(item, m, f0);
        const int q = lane >> 1, part = lane & 1;
        const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        if (f >= a.nfun || (sgn && m == 0)) return;
        const double* run = (part ? a.ico : a.rco) + (long)f * a.coef_stride + (sgn ? coef_base(-m, B) : coef_base(m, B));
        const uint64_t lo = (reinterpret_cast<uint64_t>(run) + 15) & ~15ull;
        const uint64_t hi = reinterpret_cast<uint64_t>(run + (B - m)) & ~15ull;
        if (hi > lo)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
    };

    // DMMA unit `code` of the item (m, f0) whose coefficient panel is `Cp`
    auto dmma_task = [&](int m, int f0, int code, const double* Cp) {
        const int p = code & 1, ct0 = FLOW_Q * (code >> 1);
        const BlockMeta mb0 = block_meta_of(m, 0, B);
        const BlockMeta mb = p ? block_meta_of(m, 1, B) : mb0;
        const double* tblk = a.table + ((uint64_t)ost[m] + (p ? block_tiles_of(mb0) : 0u)) * 64;
        const int rt_min = first_row_tile_reaching(mb, ct0);
        const int cnt = mb.nrt - rt_min;
        double acc[FLOW_Q][NC / 8][2];
#pragma unroll
        for (int c = 0; c < FLOW_Q; ++c)
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc[c][j][0] = acc[c][j][1] = 0.0;
        const double* cp = Cp + p * PS + g * CS + q4;
        auto tiles_at = [&](int rt) {  // tiles of row tile rt inside this unit's four column tiles
            const int t = tiles_in_row(mb, rt) - ct0;
            return t < FLOW_Q ? t : FLOW_Q;
        };
        auto issue = [&](int i) {  // elected lane: the tiles of step i into stage i % STAGES
            const int rt = rt_min + i, s = i % FLOW_STAGES;
            bulk_tile_copy(ring + s * FLOW_Q * 32, tblk + ((uint64_t)row_tile_start_of(mb, rt) + ct0) * 64,
                           512u * (unsigned)tiles_at(rt), bar0 + 8 * s);
        };
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < FLOW_STAGES; ++s)
                if (s < cnt) issue(s);
        }
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const int rt = rt_min + i, s = i % FLOW_STAGES;
            const int nv = tiles_at(rt);
            double av[NC / 8][2];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
                av[j][0] = cp[j * 8 * CS + 8 * rt];
                av[j][1] = cp[j * 8 * CS + 8 * rt + 4];
            }
            bulk_wait(bar0 + 8 * s, (phases >> s) & 1u);
            phases ^= 1u << s;
            double2 bv[FLOW_Q];
#pragma unroll
            for (int c = 0; c < FLOW_Q; ++c) bv[c] = ring[(s * FLOW_Q + c) * 32 + lane];
            if (nv == FLOW_Q) {
#pragma unroll
                for (int c = 0; c < FLOW_Q; ++c)
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][0], bv[c].x);
#pragma unroll
                for (int c = 0; c < FLOW_Q; ++c)
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][1], bv[c].y);
            } else {
#pragma unroll
                for (int c = 0; c < FLOW_Q - 1; ++c)
                    if (c < nv) {
#pragma unroll
                        for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][0], bv[c].x);
#pragma unroll
                        for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][1], bv[c].y);
                    }
            }
            // every lane's DMMAs of this step have been issued, so its fragment loads from the stage have completed:
            // the stage may be overwritten by the copy for step i + STAGES
            if (i + FLOW_STAGES < cnt) {
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(i + FLOW_STAGES);
                }
            }
        }
        // ---- epilogue: lane holds column 8j + g, cosine slots 8 (ct0 + c) + 2 q4 + {0, 1} of parity p (adjacent in the
        // parity-split plane).  Column tiles no row reaches are written too (zeros): K5 reads every slot.
        const int sgn = a.real_fmt ? 0 : ((g >> 1) & 1), part = g & 1;
        const int fl0 = a.real_fmt ? (g >> 1) : (g >> 2), flstep = a.real_fmt ? 4 : 2;
        if (sgn && m == 0) return;
        const int mp = sgn ? N - m : m;
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            const int f = f0 + fl0 + j * flstep;
            if (f >= a.nfun) continue;
            double* d = a.V + (((long)f * N + mp) * 2 + part) * B + p * (B / 2) + 8 * ct0 + 2 * q4;
#pragma unroll
            for (int c = 0; c < FLOW_Q; ++c) *reinterpret_cast<double2*>(d + 8 * c) = make_double2(acc[c][j][0], acc[c][j][1]);
        }
    };

    // ---- the task streams (as in k_fwd_uni).  Queue sq hands out tickets; ticket numbers map to (stream, task): stream -1 =
    // the four stage tasks of item 0, stream k >= 0 = the stage tasks of item k + 1 interleaved with the DMMA units of item
    // k.  A task waits only for what it depends on:
    //   DMMA unit of item k      : the 16 stage tasks of item k            (stage_done[k & 1] >= 16 (k / 2 + 1))
    //   stage task of item k + 1 : every DMMA unit of item k - 1, the last reader of that panel buffer
    int k = -1, base = 0, nd = 0, nf = nk > 0 ? 4 : 0, ntask = nf, qb = 0, m = 0, f0 = 0;
    int cum[2] = {0, 0};  // DMMA units of the items of each buffer parity up to the current stream
    int tn = 0;
    if (lane == 0) tn = atomicAdd(&tick[sq], 1);
#pragma unroll 1
    for (;;) {
        const int t_abs = __shfl_sync(0xffffffffu, tn, 0);
        while (t_abs >= base + ntask) {  // the ticket belongs to a later stream
            base += ntask;
            if (++k >= nk) break;
            item_of((int)blockIdx.x + k * (int)gridDim.x, m, f0);
            qb = qoff[4 * m + sq];
            nd = qoff[4 * m + sq + 1] - qb;
            nf = (k + 1 < nk) ? 4 : 0;
            ntask = nd + nf;
            cum[k & 1] += qoff[4 * m + 4] - qoff[4 * m];
        }
        if (k >= nk) break;
        if (lane == 0) tn = atomicAdd(&tick[sq], 1);  // the task after this one: its latency hides behind this task
        const int t = t_abs - base;
        // stream order: stage and DMMA tasks alternate while both kinds last, then the rest
        bool is_stage;
        int idx;
        {
            const int mi = nf < nd ? nf : nd;
            if (t < 2 * mi) {
                is_stage = !(t & 1);
                idx = t >> 1;
            } else if (nf > nd) {
                is_stage = true;
                idx = t - nd;
            } else {
                is_stage = false;
                idx = t - nf;
            }
        }
        const int item = (int)blockIdx.x + k * (int)gridDim.x;  // (k = -1: only stage tasks, of item 0)
        if (sq == 0 && t == 0 && k >= 0) {  // once per item: pull what the coming streams read into L2
            if (k + 1 < nk) {
                const int m2 = a.m_lo + (item + (int)gridDim.x) / a.ncoltiles;
                if (m2 != m) order_prefetch(m2);
            }
            if (k + 2 < nk) prefetch_cols(item + 2 * (int)gridDim.x);
        }
        if (is_stage) {
            flow_wait_ge(&done[2 + ((k + 1) & 1)], cum[(k + 1) & 1], a.sleep_ns);  // the buffer's previous readers
            stage_task(item + (int)gridDim.x, 4 * sq + idx, panels + ((k + 1) & 1) * PANEL);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) flow_signal(&done[(k + 1) & 1]);
        } else {
            flow_wait_ge(&done[k & 1], 16 * (k / 2 + 1), a.sleep_ns);  // the item's panel is complete
            dmma_task(m, f0, qlist[qb + idx], panels + (k & 1) * PANEL);
            __syncwarp();
            if (lane == 0) flow_signal(&done[2 + (k & 1)]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ launcher
// Opt-in (S2KIT_CUDA_FLOW=1).  Measured at bw = 256, 1024 functions: 1.80 ms against 1.19 ms for k_legendre_inv<32,32>.
// The DMMA phase itself is lean (LSU data pipe 31 % busy against 70 %), but the pipe is only 43 % busy: an item has eight
// units of very unequal length (16 / 12 / 8 / 4 row-tile steps at m = 0) for sixteen warps, its critical path is the longest
// unit (16 steps x 32 DMMAs), and with two panel buffers the staging of item k + 1 cannot start before the last unit of
// item k - 1 has finished -- 40 % of the warp samples sit in the dependency polls (profiles/r2_ncu_flow_summary.md).
static bool flow_enabled() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_FLOW");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}

bool inv_flow_supported(const s2kit_cuda_plan* p, int nfun, int data_format) {
    if (!flow_enabled() || !p->fast || p->n != 512 || !p->d_iq_list || p->table_single) return false;
    return nfun * (data_format == S2KIT_REAL ? 2 : 4) >= FLOW_NC;
}

cudaError_t launch_inv_flow(s2kit_cuda_plan* p, const double* table_t, uint64_t shift, const double* rco, const double* ico,
                            long coef_stride, double* V, int nfun, int m_lo, int m_hi, int data_format) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    FlowInvArgs a;
    a.table = table_t;
    a.order_start = p->d_order_start;
    a.table_shift = shift;
    a.sub_off = p->d_iq_off;
    a.sub_list = p->d_iq_list;
    a.nlist = p->n_iq_list;
    a.rco = rco;
    a.ico = ico;
    a.coef_stride = coef_stride;
    a.V = V;
    a.nfun = nfun;
    a.m_lo = m_lo;
    a.norders = m_hi - m_lo;
    a.real_fmt = data_format == S2KIT_REAL;
    const int NF = FLOW_NC / (a.real_fmt ? 2 : 4);
    a.ncoltiles = (nfun + NF - 1) / NF;
    static const int sleep_ns = [] {
        const char* e = getenv("S2KIT_CUDA_UNI_SLEEP");
        return e ? atoi(e) : 32;
    }();
    a.sleep_ns = sleep_ns;
    constexpr int PANEL = 2 * (FLOW_NC * 132 + 8);
    const size_t smem = sizeof(double) * 2 * PANEL + sizeof(double2) * FLOW_WARPS * FLOW_STAGES * FLOW_Q * 32 +
                        8 * FLOW_WARPS * FLOW_STAGES + 64 + flow_meta_bytes(256, p->n_iq_list);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_inv_flow), smem);
    if (e != cudaSuccess) return e;
    const int nitems = a.norders * a.ncoltiles;
    int slot = prof_begin(p, S2KIT_K_LEGENDRE_INV);
    k_inv_flow<<<nitems < p->sm_count ? nitems : p->sm_count, FLOW_THREADS, smem, p->stream>>>(a);
    e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
