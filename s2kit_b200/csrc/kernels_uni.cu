// kernels_uni.cu -- the batched forward path at bw = 256 as ONE persistent kernel of UNIFORM warps: K2 + K3 fused.
//
//   weights . DCT-II(2bw) . triangular contraction . coefficient placement
//   (DLTSemi, src/legendre_transform/seminaive.c:153-198, inside the m-loops of FSTSemiMemo,
//    src/FST_semi_memo.c:96-108,131-145,175-201)
//
// Successor of the warp-specialised k_fwd_pipe (kernels_pipe.cu), which ran the FP64 pipe at 50 % (DMMA 28.8 % + DFMA
// 21.7 %, profiles/r1_ncu_pipe_summary.md): its 8 DCT warps and 8 DMMA warps share one pipe, the kernel cost the SUM of
// the two halves, and whenever the consumers waited for a panel only the latency-bound FFT warps were left to feed it.
// Here there are no roles.  Work items are still (order m, 32-column tile), dealt round-robin to one CTA per SM, but
// inside the CTA every item is cut into TASKS that any of the 16 warps takes from a shared-memory counter:
//   * 16 DCT tasks of the NEXT item: one warp turns the weighted real / imaginary spectral rows of one (function, +-m)
//     into two panel columns with the one-warp 512-point FFT (s2k_fft16.cuh / s2k_dct16.cuh: 16 points per lane, one
//     shared-memory exchange done in place inside the warp's own two panel columns, no named barriers, a third fewer
//     FP64 instructions per point than the 8-points-per-thread block FFT);
//   * the DMMA sub-items of the CURRENT item: single row tiles for the long rows, pairs of adjacent row tiles (which
//     share every panel fragment) for the short ones, heaviest first (host-built list, plan.cu), table tiles through
//     the lane-private cp.async ring that removed the table-latency stalls in k_leg_fwd_stream.
// DCT and DMMA tasks alternate in the task order, so at any time about half of the warps issue DFMAs and half DMMAs:
// the latency gaps of the FFT code are filled by DMMAs of other warps and the pipe stays busy.  One __syncthreads per
// item separates the panel generations (two panel buffers).
// What the first version's profile showed (profiles/r2_ncu_uni_summary.md: FP64 + DMMA pipes 46 % busy, 14 % of the warp
// samples on the task counter and the dependent metadata loads behind it, 14 % on the DCT's global loads, 11 % at the
// barrier) and what this version does about it:
//   * a warp never leaves its SM sub-partition and every sub-partition has its own FP64 / DMMA pipe: the units of an
//     order are split on the host into four queues of equal cost, one per sub-partition (plan.cu, lpt_queues);
//   * all per-order metadata (queue offsets, unit codes, table offsets) is copied into shared memory once per CTA;
//   * a warp takes its NEXT task from the counter before it starts the current one (and the first task of the next
//     item before the barrier), so the atomic's latency is never waited for;
//   * the spectral rows the DCT tasks will read are pulled into L2 two items ahead (cp.async.bulk.prefetch.L2).
#include <stdlib.h>

#include "s2k_dct16.cuh"
#include "s2k_legendre.cuh"

namespace s2k {

constexpr int UNI_NC = 32;
constexpr int UNI_WARPS = 16;  // 128 registers per thread.  20 warps at 96 registers (spills, 6-tile rings) were measured slower: 2.67 vs 2.37 ms
constexpr int UNI_THREADS = UNI_WARPS * 32;
constexpr int UNI_RING = 8;  // table tiles in flight per warp (8 x 512 B)

struct UniArgs {
    const double* table;
    const uint64_t* order_start;
    uint64_t table_shift;
    const int* sub_off;               // [4 bw + 1]: queue s of order m = sub_list[sub_off[4m+s] .. sub_off[4m+s+1])
    const unsigned short* sub_list;   // packed units: parity | row tile << 1 | pair << 12
    int nlist;
    const double* S;       // spectral planes [f][part][order row][latitude slot]
    const double* weights; // 4bw, load order (s2k_host_reordered)
    const double2* tw;
    const double2* qtab;
    double* rco;
    double* ico;
    long coef_stride;
    int nfun, m_lo, norders, ncoltiles, real_fmt, lat_perm;
    int debug_skip;  // measurement only (S2KIT_CUDA_UNI_SKIP): 1 = DCT tasks do nothing, 2 = DMMA units do nothing
    int lead;        // DMMA units at the head of a stream before the DCT tasks start to alternate with them
    int sleep_ns;    // back-off of the dependency polls (0 = plain polling)
};

// per-CTA copy of the per-order metadata (shared memory): queue offsets, unit codes, first tile of every order
struct UniMeta {
    const int* qoff;
    const unsigned short* qlist;
    const unsigned* ost;
};
template <int B>
__device__ __forceinline__ UniMeta uni_stage_meta(void* at, const int* sub_off, const unsigned short* sub_list, int nlist,
                                                  const uint64_t* order_start, uint64_t shift, int m_lo, int m_hi) {
    int* qoff = reinterpret_cast<int*>(at);
    unsigned* ost = reinterpret_cast<unsigned*>(qoff + 4 * B + 4);
    unsigned short* qlist = reinterpret_cast<unsigned short*>(ost + B + 4);
    for (int i = threadIdx.x; i <= 4 * B; i += blockDim.x) qoff[i] = sub_off[i];
    for (int i = threadIdx.x; i <= B; i += blockDim.x)  // only the launch's orders are resident (Fly groups): clamp the rest
        ost[i] = (i >= m_lo && i <= m_hi) ? (unsigned)(order_start[i] - shift) : 0u;
    for (int i = threadIdx.x; i < nlist; i += blockDim.x) qlist[i] = sub_list[i];
    UniMeta mt;
    mt.qoff = qoff;
    mt.qlist = qlist;
    mt.ost = ost;
    return mt;
}
__host__ __device__ constexpr size_t uni_meta_bytes(int B, int nlist) {
    return sizeof(int) * (4 * B + 4) + sizeof(unsigned) * (B + 4) + sizeof(unsigned short) * ((nlist + 7) & ~7);
}

// dependency counters in shared memory: release-increment by one lane after its warp's work (ordered by __syncwarp),
// acquire-poll by every lane of a waiting warp
__device__ __forceinline__ void uni_signal(int* ctr) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(ctr)))
                 : "memory");
}
__device__ __forceinline__ void uni_wait_ge(const int* ctr, int need, int sleep_ns) {
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

__device__ __forceinline__ void uni_item(const UniArgs& a, int t, int NF, int& m, int& f0) {
    const int oi = t / a.ncoltiles, x = t - oi * a.ncoltiles;
    m = a.m_lo + oi;
    f0 = x * NF;
}

// One DMMA sub-item: row tile rt1 (and, PAIR, rt1 - 1 = rt0) of one parity block against the panel.
// tp0 / tp1: this lane's 16 bytes of the first tile of row tile rt0 / rt1; skip0 / skip1: the second k-step of the last
// column tile multiplies only padding.
template <bool PAIR>
__device__ __forceinline__ void uni_rows(const double* __restrict__ tp0, int ctn0, bool skip0,
                                         const double* __restrict__ tp1, int ctn1, bool skip1, const double* xp,
                                         double (&acc0)[UNI_NC / 8][2], double (&acc1)[UNI_NC / 8][2], double2* ring) {
    constexpr int NC = UNI_NC, CS = 132;
    constexpr int STEPS = PAIR ? UNI_RING / 2 : UNI_RING;  // column-tile steps in flight
#pragma unroll
    for (int u = 0; u < STEPS; ++u) {
        if (u < ctn1) {
            if (PAIR && u < ctn0) cp_async16(reinterpret_cast<double*>(ring + ((2 * u) & (UNI_RING - 1)) * 32), tp0 + u * 64);
            cp_async16(reinterpret_cast<double*>(ring + ((PAIR ? 2 * u + 1 : u) & (UNI_RING - 1)) * 32), tp1 + u * 64);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int ct = 0; ct < ctn1; ++ct) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STEPS - 1) : "memory");
        double2* s1 = ring + ((PAIR ? 2 * ct + 1 : ct) & (UNI_RING - 1)) * 32;
        double2* s0 = ring + ((2 * ct) & (UNI_RING - 1)) * 32;
        const double2 a1 = *s1;
        double2 a0 = make_double2(0.0, 0.0);
        const bool have0 = PAIR && ct < ctn0;
        if (have0) a0 = *s0;
        double b[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            b[j][0] = xp[j * 8 * CS + 8 * ct];
            b[j][1] = xp[j * 8 * CS + 8 * ct + 4];
        }
        if (have0) {
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a0.x, b[j][0]);
        }
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a1.x, b[j][0]);
        if (have0 && !(skip0 && ct == ctn0 - 1)) {
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a0.y, b[j][1]);
        }
        if (!(skip1 && ct == ctn1 - 1)) {
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a1.y, b[j][1]);
        }
        // the DMMAs above have consumed the slots' registers: refill with the step STEPS ahead.  (Loading the next step's
        // fragments before this step's DMMAs -- a software pipeline -- was measured SLOWER: 2.64 vs 2.37 ms per 1024
        // functions; the extra live registers and moves cost more than the exposed shared-memory latency.)
        const int nx = ct + STEPS;
        if (nx < ctn1) {
            if (PAIR && nx < ctn0) cp_async16(reinterpret_cast<double*>(s0), tp0 + nx * 64);
            cp_async16(reinterpret_cast<double*>(s1), tp1 + nx * 64);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(UNI_THREADS, 1) k_fwd_uni(const UniArgs a) {
    constexpr int N = 512, B = 256, NC = UNI_NC, CS = 132, PS = NC * CS + 8, PANEL = 2 * PS;
    extern __shared__ __align__(16) double smem[];
    double* panels = smem;                                                  // [2][PANEL]
    double2* rings = reinterpret_cast<double2*>(smem + 2 * PANEL);          // [WARPS][RING][32]
    int* tick = reinterpret_cast<int*>(rings + UNI_WARPS * UNI_RING * 32);  // [4] ticket counters of the task queues
    int* done = tick + 4;  // [0..1] dct_done, [2..3] dmma_done, per panel-buffer parity
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const int sq = warp & 3;  // this warp's SM sub-partition = its task queue
    const int cols_per_fn = a.real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int nitems = a.norders * a.ncoltiles;
    const int nk = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    double2* ring = rings + warp * UNI_RING * 32 + lane;
    const double s_all = 1.0 / sqrt(2.0 * (double)N);  // 1/sqrt(2*size), seminaive.c:174

    if (tid < 8) tick[tid] = 0;
    const UniMeta mt = uni_stage_meta<B>(tick + 16, a.sub_off, a.sub_list, a.nlist, a.order_start, a.table_shift, a.m_lo,
                                         a.m_lo + a.norders);
    __syncthreads();  // the only CTA-wide barrier
    if (warp == 0 && nk > 0) {
        const int m0 = a.m_lo + (int)blockIdx.x / a.ncoltiles;
        prefetch_order_l2(a.table + (uint64_t)mt.ost[m0] * 64, a.order_start[m0 + 1] - a.order_start[m0], lane, 32, 1u << 20);
    }

    // DCT task `q` of item `item`: panel columns 2q (real part) and 2q + 1 (imaginary part) of buffer `buf`
    auto dct_task = [&](int item, int q, double* buf) {
        int m, f0;
        uni_item(a, item, NF, m, f0);
        const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        double* col0 = buf + (2 * q) * CS;
        if (f >= a.nfun || (sgn && m == 0)) {  // dead pair: the contraction still multiplies these columns
            for (int i = lane; i < 2 * CS; i += 32) col0[i] = col0[PS + i] = 0.0;
            return;
        }
        const int mp = sgn ? N - m : m;
        const double* Sr = a.S + ((long)f * 2 * N + mp) * N;
        const double* Si = Sr + (long)N * N;
        const double* w = a.weights + ((m & 1) ? N : 0);
        double xr[16], xi[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int p = lane + 32 * e;
            const int at = a.lat_perm ? p : ((p < B) ? 2 * p : 2 * (N - 1 - p) + 1);
            const double wj = __ldg(w + p);
            xr[e] = __ldg(Sr + at) * wj;
            xi[e] = __ldg(Si + at) * wj;
        }
        d16_dct2_pair_to_panel<CS>(xr, xi, col0, PS, lane, a.tw, a.qtab, s_all);
    };

    // the 32 spectral rows (16 pairs x re / im, 4 KiB each) the DCT tasks of `item` will read: into L2, one row per lane
    auto prefetch_rows = [&](int item) {
        int m, f0;
        uni_item(a, item, NF, m, f0);
        const int q = lane >> 1, part = lane & 1;
        const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        if (f >= a.nfun || (sgn && m == 0)) return;
        const double* row = a.S + (((long)f * 2 + part) * N + (sgn ? N - m : m)) * N;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(N * 8) : "memory");
    };

    // DMMA unit `code` of the item (m, f0) whose panel is `Xs`
    auto dmma_task = [&](int m, int f0, int code, const double* Xs) {
        const int p = code & 1, rt1 = (code >> 1) & 0x7ff, pair = code >> 12;
        const BlockMeta mb0 = block_meta_of(m, 0, B);
        const BlockMeta mb = p ? block_meta_of(m, 1, B) : mb0;
        const double* tblk = a.table + ((uint64_t)mt.ost[m] + (p ? block_tiles_of(mb0) : 0u)) * 64 + lane * 2;
        const double* xp = Xs + p * PS + g * CS + q4;
        const int ctn1 = tiles_in_row(mb, rt1);
        const bool skip1 = mb.len0 + min(8 * rt1 + 7, mb.rows - 1) - 8 * (ctn1 - 1) <= 4;
        const double* tp1 = tblk + (uint64_t)row_tile_start_of(mb, rt1) * 64;
        double acc0[NC / 8][2], acc1[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
        const int rt0 = pair ? rt1 - 1 : -1;
        if (pair) {
            const int ctn0 = tiles_in_row(mb, rt0);
            const bool skip0 = mb.len0 + 8 * rt0 + 7 - 8 * (ctn0 - 1) <= 4;  // rt0 < rt1: a full row tile
            uni_rows<true>(tblk + (uint64_t)row_tile_start_of(mb, rt0) * 64, ctn0, skip0, tp1, ctn1, skip1, xp, acc0, acc1, ring);
        } else {
            uni_rows<false>(tp1, 0, false, tp1, ctn1, skip1, xp, acc0, acc1, ring);
        }
        // ---- epilogue (k_fwd_pipe / K3): the lane's columns 8j + 2 q4 + {0,1} = (function, sign, re / im) in closed form
        const int sgn = a.real_fmt ? 0 : (q4 & 1);
        const int fl0 = a.real_fmt ? q4 : (q4 >> 1), flstep = a.real_fmt ? 4 : 2;
        const long run0 = (long)(f0 + fl0) * a.coef_stride + (sgn ? coef_base(-m, B) : coef_base(m, B));
        const long mrun0 = (long)(f0 + fl0) * a.coef_stride + coef_base(-m, B);
        const unsigned long long flip = (sgn && (m & 1)) ? 0x8000000000000000ull : 0ull;  // FST_semi_memo.c:181-186
        const bool live_sign = !(sgn && m == 0);
        const bool mirror = a.real_fmt && m > 0;  // FST_semi_memo.c:131-145
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 8 * (h ? rt1 : rt0) + g;
            if ((h == 0 && rt0 < 0) || r >= mb.rows || !live_sign) continue;
            const int off = p + 2 * r;  // l - m
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
                if (f0 + fl0 + j * flstep >= a.nfun) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(h ? acc1[j][e] : acc0[j][e]);
                    double* arr = e ? a.ico : a.rco;
                    arr[run0 + (long)j * flstep * a.coef_stride + off] = __longlong_as_double((long long)(bits ^ flip));
                    if (mirror) {
                        const unsigned long long mflip = ((m & 1) ^ e) ? 0x8000000000000000ull : 0ull;
                        arr[mrun0 + (long)j * flstep * a.coef_stride + off] = __longlong_as_double((long long)(bits ^ mflip));
                    }
                }
            }
        }
    };

    // ---- the task streams.  Queue sq hands out tickets; ticket numbers map to (stream, task): stream -1 = the four DCT
    // tasks of item 0, stream k >= 0 = the DCT tasks of item k + 1 interleaved with the DMMA units of item k.  There is no
    // CTA-wide barrier: a task waits only for what it depends on --
    //   DMMA unit of item k      : the 16 DCT tasks of item k          (dct_done[k & 1] >= 16 (k / 2 + 1))
    //   DCT task of item k + 1   : every DMMA unit of item k - 1, the last reader of that panel buffer
    //                                                                   (dmma_done[(k - 1) & 1] >= units of its parity so far)
    // -- so warps flow from item to item and a slow task delays only its dependants.  Counters are per buffer parity:
    // items of one parity are strictly ordered by these two rules, items of different parity may overlap.
    int k = -1, base = 0, nd = 0, nf = nk > 0 ? 4 : 0, ntask = nf, mi = 0, qb = 0, m = 0, f0 = 0;
    int cum[2] = {0, 0};  // DMMA units of the items of each parity up to the current stream
    int tn = 0;
    if (lane == 0) tn = atomicAdd(&tick[sq], 1);
#pragma unroll 1
    for (;;) {
        const int t_abs = __shfl_sync(0xffffffffu, tn, 0);
        while (t_abs >= base + ntask) {  // the ticket belongs to a later stream
            base += ntask;
            if (++k >= nk) break;
            uni_item(a, (int)blockIdx.x + k * (int)gridDim.x, NF, m, f0);
            qb = mt.qoff[4 * m + sq];
            nd = mt.qoff[4 * m + sq + 1] - qb;
            nf = (k + 1 < nk) ? 4 : 0;
            ntask = nd + nf;
            mi = nf < nd ? nf : nd;
            cum[k & 1] += mt.qoff[4 * m + 4] - mt.qoff[4 * m];
        }
        if (k >= nk) break;
        if (lane == 0) tn = atomicAdd(&tick[sq], 1);  // the task after this one: its latency hides behind this task
        const int t = t_abs - base;
        // stream order: `lead` DMMA units first (the heaviest; by the time the DCT tasks come up, the previous readers of
        // their panel buffer are done), then DCT and DMMA tasks alternate while both kinds last, then the rest
        bool is_dct;
        int idx;
        {
            const int ld = nf ? (a.lead < nd ? a.lead : nd) : 0, t2 = t - ld, nd2 = nd - ld, mi2 = nf < nd2 ? nf : nd2;
            if (t2 < 0) {
                is_dct = false;
                idx = t;
            } else if (t2 < 2 * mi2) {
                is_dct = !(t2 & 1);
                idx = is_dct ? (t2 >> 1) : ld + (t2 >> 1);
            } else if (nf > nd2) {
                is_dct = true;
                idx = t2 - nd2;
            } else {
                is_dct = false;
                idx = ld + t2 - nf;
            }
        }
        const int item = (int)blockIdx.x + k * (int)gridDim.x;  // (k = -1: only DCT tasks, of item 0)
        if (sq == 0 && t == 0 && k >= 0) {  // once per item: pull what the coming streams read into L2
            if (k + 1 < nk) {
                const int m2 = a.m_lo + (item + (int)gridDim.x) / a.ncoltiles;
                if (m2 != m)
                    prefetch_order_l2(a.table + (uint64_t)mt.ost[m2] * 64, a.order_start[m2 + 1] - a.order_start[m2], lane, 32,
                                      1u << 20);
            }
            if (k + 2 < nk) prefetch_rows(item + 2 * (int)gridDim.x);
        }
        if (is_dct) {
            uni_wait_ge(&done[2 + ((k + 1) & 1)], cum[(k + 1) & 1], a.sleep_ns);  // dmma_done of the buffer's previous readers
            if (a.debug_skip != 1) dct_task(item + (int)gridDim.x, 4 * sq + idx, panels + ((k + 1) & 1) * PANEL);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) uni_signal(&done[(k + 1) & 1]);
        } else {
            uni_wait_ge(&done[k & 1], 16 * (k / 2 + 1), a.sleep_ns);  // dct_done: the item's panel is complete
            if (a.debug_skip != 2) dmma_task(m, f0, mt.qlist[qb + idx], panels + (k & 1) * PANEL);
            __syncwarp();
            if (lane == 0) uni_signal(&done[2 + (k & 1)]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ launcher
static bool uni_enabled() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_UNI");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on != 0;
}

bool fwd_uni_supported(const s2kit_cuda_plan* p, int nfun, int data_format) {
    if (!uni_enabled() || !p->fast || p->n != 512 || !p->d_sub_list) return false;
    return nfun * (data_format == S2KIT_REAL ? 2 : 4) >= UNI_NC;
}

cudaError_t launch_fwd_uni(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* S, double* rco,
                           double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format, int lat_perm) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    UniArgs a;
    a.table = table;
    a.order_start = p->d_order_start;
    a.table_shift = shift;
    a.sub_off = p->d_sub_off;
    a.sub_list = p->d_sub_list;
    a.nlist = p->n_sub_list;
    a.S = S;
    a.weights = p->d_wv;
    a.tw = p->d_tw_n;
    a.qtab = p->d_q_n;
    a.rco = rco;
    a.ico = ico;
    a.coef_stride = coef_stride;
    a.nfun = nfun;
    a.m_lo = m_lo;
    a.norders = m_hi - m_lo;
    a.real_fmt = data_format == S2KIT_REAL;
    const int NF = UNI_NC / (a.real_fmt ? 2 : 4);
    a.ncoltiles = (nfun + NF - 1) / NF;
    a.lat_perm = lat_perm;
    static const int skip = [] {
        const char* e = getenv("S2KIT_CUDA_UNI_SKIP");
        return e ? atoi(e) : 0;
    }();
    a.debug_skip = skip;
    static const int lead = [] {
        const char* e = getenv("S2KIT_CUDA_UNI_LEAD");
        return e ? atoi(e) : 0;
    }();
    static const int sleep_ns = [] {
        const char* e = getenv("S2KIT_CUDA_UNI_SLEEP");
        return e ? atoi(e) : 32;
    }();
    a.lead = lead;
    a.sleep_ns = sleep_ns;
    constexpr int PANEL = 2 * (UNI_NC * 132 + 8);
    const size_t smem = sizeof(double) * 2 * PANEL + sizeof(double2) * UNI_WARPS * UNI_RING * 32 + 64 +
                        uni_meta_bytes(256, p->n_sub_list);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_fwd_uni), smem);
    if (e != cudaSuccess) return e;
    const int nitems = a.norders * a.ncoltiles;
    int slot = prof_begin(p, S2KIT_K_FUSED_FWD);
    k_fwd_uni<<<nitems < p->sm_count ? nitems : p->sm_count, UNI_THREADS, smem, p->stream>>>(a);
    e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

// =====================================================================================================================
// The batched INVERSE path at bw = 256 as one persistent kernel: K4 + K5 fused.
//
//   transposed triangular contraction . scaling . DCT-III(2bw) . sin(theta) . (-1)^m . 1/sqrt(2 pi)
//   (InvDLTSemi, src/legendre_transform/seminaive.c:56-115, inside the m-loops of InvFSTSemiMemo,
//    src/FST_semi_memo.c:262-342)
//
// The cosine planes V never go through HBM (4.3 of the 9.7 GB the two separate kernels move per 1024 functions).  Per
// work item (order m, 32 columns) the CTA alternates between two phases, all 16 warps in each:
//   A  contraction: V[col, k] = sum_l c[col, l] T_m[l, k] on DMMA -- coefficient panel (A fragments) in shared memory,
//      table tiles (B-fragment order) through the lane-private cp.async rings, sub-items = column tiles of one parity,
//      alone where many row tiles reach them, in adjacent pairs (which share every coefficient fragment) otherwise,
//      taken heaviest first from a shared-memory counter; results land in the V panel in shared memory;
//   B  DCT-III: warp w turns panel columns 2w, 2w + 1 (real / imaginary part of one (function, +-m)) into a spectral-plane
//      row pair with the one-warp FFT (exchange in place in those two columns) and writes it to G; meanwhile the
//      coefficient panel of the NEXT item streams in with cp.async (the panel is idle during this phase).
// Shared memory: coefficient panel + V panel + rings = 199 KB; a second V panel (to overlap the phases as the forward
// kernel does) does not fit beside the coefficient panel.
struct UniInvArgs {
    const double* table;  // B-fragment-ordered tiles
    const uint64_t* order_start;
    uint64_t table_shift;
    const int* sub_off;              // [4 bw + 1], four queues per order
    const unsigned short* sub_list;  // parity | column tile << 1 | pair << 12
    int nlist;
    const double* rco;
    const double* ico;
    long coef_stride;
    double* G;            // spectral planes [f][part][order row][latitude slot]
    const double* sinv;   // sines in the DCT's output order (s2k_host_reordered)
    const double2* tw;
    const double2* qtab;
    int nfun, m_lo, norders, ncoltiles, real_fmt, lat_perm;
};

// One contraction sub-item: column tile ct (and, PAIR, ct + 1) of parity block mb over the row tiles that reach it.
template <bool PAIR>
__device__ __forceinline__ void uni_cols(const double* __restrict__ tblk, const BlockMeta& mb, int ct, const double* cp,
                                         double (&acc0)[UNI_NC / 8][2], double (&acc1)[UNI_NC / 8][2], double2* ring) {
    constexpr int NC = UNI_NC, CS = 132;
    constexpr int STEPS = PAIR ? UNI_RING / 2 : UNI_RING;
    const int rt_min = first_row_tile_reaching(mb, ct);
    const int rt_min1 = PAIR ? first_row_tile_reaching(mb, ct + 1) : 0;  // >= rt_min: rows only grow
    const int cnt = mb.nrt - rt_min;
    if (cnt <= 0) return;
    auto tile = [&](int rt) { return tblk + ((uint64_t)row_tile_start_of(mb, rt) + ct) * 64; };
#pragma unroll
    for (int u = 0; u < STEPS; ++u) {
        if (u < cnt) {
            const int rt = rt_min + u;
            const double* t = tile(rt);
            cp_async16(reinterpret_cast<double*>(ring + ((PAIR ? 2 * u : u) & (UNI_RING - 1)) * 32), t);
            if (PAIR && rt >= rt_min1) cp_async16(reinterpret_cast<double*>(ring + ((2 * u + 1) & (UNI_RING - 1)) * 32), t + 64);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) {
        const int rt = rt_min + i;
        asm volatile("cp.async.wait_group %0;" ::"n"(STEPS - 1) : "memory");
        double2* s0 = ring + ((PAIR ? 2 * i : i) & (UNI_RING - 1)) * 32;
        double2* s1 = ring + ((2 * i + 1) & (UNI_RING - 1)) * 32;
        const bool two = PAIR && rt >= rt_min1;
        const double2 b0 = *s0;
        double2 b1 = make_double2(0.0, 0.0);
        if (two) b1 = *s1;
        double av[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            av[j][0] = cp[j * 8 * CS + 8 * rt];
            av[j][1] = cp[j * 8 * CS + 8 * rt + 4];
        }
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], av[j][0], b0.x);
        if (two) {
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], av[j][0], b1.x);
        }
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], av[j][1], b0.y);
        if (two) {
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], av[j][1], b1.y);
        }
        const int nx = i + STEPS;
        if (nx < cnt) {
            const int rn = rt_min + nx;
            const double* t = tile(rn);
            cp_async16(reinterpret_cast<double*>(s0), t);
            if (PAIR && rn >= rt_min1) cp_async16(reinterpret_cast<double*>(s1), t + 64);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(UNI_THREADS, 1) k_inv_uni(const UniInvArgs a) {
    constexpr int N = 512, B = 256, NC = UNI_NC, CS = 132, PS = NC * CS + 8, PANEL = 2 * PS;
    extern __shared__ __align__(16) double smem[];
    double* Cp = smem;                                                       // coefficient panel [2][NC][CS]
    double* Vp = smem + PANEL;                                               // cosine panel      [2][NC][CS]
    double2* rings = reinterpret_cast<double2*>(smem + 2 * PANEL);           // [WARPS][RING][32]
    int* ctr = reinterpret_cast<int*>(rings + UNI_WARPS * UNI_RING * 32);    // [3 generations][4 queues]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const int sq = warp & 3;
    const int cols_per_fn = a.real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int nitems = a.norders * a.ncoltiles;
    const int nk = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    double2* ring = rings + warp * UNI_RING * 32 + lane;
    const double c_rest = 1.0 / sqrt(2.0 * (double)N);  // 0.5/sqrt(bw), seminaive.c:72
    const double c_zero = 1.0 / sqrt((double)N);        // fcos[0] / sqrt(2 bw), seminaive.c:98
    const double out_scale = 0.39894228040143267794;    // 1/sqrt(2 pi), FST_semi_memo.c:344

    if (tid < 12) ctr[tid] = 0;
    const UniMeta mt = uni_stage_meta<B>(ctr + 16, a.sub_off, a.sub_list, a.nlist, a.order_start, a.table_shift, a.m_lo,
                                         a.m_lo + a.norders);

    // coefficient panel of item `item`: column c = (function, sign, re / im), de-interleaved by the parity of l - m; rows
    // beyond the order's degrees and dead columns are zero (they meet table padding, which must not see NaN garbage)
    auto stage_coeffs = [&](int item) {
        const int oi = item / a.ncoltiles, x = item - oi * a.ncoltiles;
        const int m = a.m_lo + oi, f0 = x * NF;
        const int cnt = B - m, h0 = (cnt + 1) / 2, h1 = cnt / 2;
        const int base_pos = coef_base(m, B), base_neg = coef_base(-m, B);
        for (int col = warp; col < NC; col += UNI_WARPS) {
            const int fl = col / cols_per_fn, sub = col % cols_per_fn;
            const int sgn = a.real_fmt ? 0 : (sub >> 1), part = sub & 1, f = f0 + fl;
            double* d0 = Cp + col * CS;
            double* d1 = d0 + PS;
            if (f >= a.nfun || (sgn && m == 0)) {
                for (int c = lane; c < CS; c += 32) d0[c] = d1[c] = 0.0;
                continue;
            }
            const double* src = (part ? a.ico : a.rco) + (long)f * a.coef_stride + (sgn ? base_neg : base_pos);
            for (int o = lane; o < cnt; o += 32) cp_async8(((o & 1) ? d1 : d0) + (o >> 1), src + o);
            for (int c = h0 + lane; c < CS; c += 32) d0[c] = 0.0;
            for (int c = h1 + lane; c < CS; c += 32) d1[c] = 0.0;
        }
    };

    int tn = 0;
    if (nk > 0) {
        stage_coeffs(blockIdx.x);
        if (warp == 0) {
            const int m0 = a.m_lo + (int)blockIdx.x / a.ncoltiles;
            prefetch_order_l2(a.table + (a.order_start[m0] - a.table_shift) * 64, a.order_start[m0 + 1] - a.order_start[m0], lane,
                              32, 1u << 20);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (nk > 0 && lane == 0) tn = atomicAdd(&ctr[sq], 1);

#pragma unroll 1
    for (int k = 0; k < nk; ++k) {
        const int item = blockIdx.x + k * gridDim.x, gen = k % 3;
        const bool has_next = k + 1 < nk;
        const int oi = item / a.ncoltiles, x = item - oi * a.ncoltiles;
        const int m = a.m_lo + oi, f0 = x * NF;
        const BlockMeta mb0 = block_meta_of(m, 0, B), mb1 = block_meta_of(m, 1, B);
        if (tid < 4) ctr[((k + 2) % 3) * 4 + tid] = 0;
        if (warp == 1 && has_next) {  // the next item's table tiles into L2
            const int m2 = a.m_lo + (item + (int)gridDim.x) / a.ncoltiles;
            if (m2 != m)
                prefetch_order_l2(a.table + (uint64_t)mt.ost[m2] * 64, a.order_start[m2 + 1] - a.order_start[m2], lane, 32,
                                  1u << 20);
        }
        // ---------------------------------------------------------------------------------- phase A: contraction
        const int qb = mt.qoff[4 * m + sq], nd = mt.qoff[4 * m + sq + 1] - qb;
        const double* tord = a.table + (uint64_t)mt.ost[m] * 64 + lane * 2;
        int t = __shfl_sync(0xffffffffu, tn, 0);
        while (t < nd) {
            if (lane == 0) tn = atomicAdd(&ctr[gen * 4 + sq], 1);
            const int code = mt.qlist[qb + t];
            const int p = code & 1, ct = (code >> 1) & 0x7ff, pair = code >> 12;
            const BlockMeta mb = p ? mb1 : mb0;
            const double* tblk = tord + (uint64_t)(p ? block_tiles_of(mb0) : 0u) * 64;
            double acc0[NC / 8][2], acc1[NC / 8][2];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
            const double* cp = Cp + p * PS + g * CS + q4;
            if (pair)
                uni_cols<true>(tblk, mb, ct, cp, acc0, acc1, ring);
            else
                uni_cols<false>(tblk, mb, ct, cp, acc0, acc1, ring);
            // lane holds column 8j + g, cosine slots 8 (ct + h) + 2 q4 + {0, 1} of parity p (adjacent in the panel)
            double* vp = Vp + p * PS + g * CS + 8 * ct + 2 * q4;
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
                *reinterpret_cast<double2*>(vp + j * 8 * CS) = make_double2(acc0[j][0], acc0[j][1]);
                if (pair) *reinterpret_cast<double2*>(vp + j * 8 * CS + 8) = make_double2(acc1[j][0], acc1[j][1]);
            }
            t = __shfl_sync(0xffffffffu, tn, 0);
        }
        if (has_next && lane == 0) tn = atomicAdd(&ctr[((k + 1) % 3) * 4 + sq], 1);
        __syncthreads();  // V panel complete, coefficient panel drained
        // ---------------------------------------------------------------------------------- phase B: DCT-III
        if (has_next) stage_coeffs(item + gridDim.x);  // streams in while the transforms run
        {
            const int q = warp;  // pair index: panel columns 2q (real part), 2q + 1 (imaginary part)
            const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
            const int f = f0 + fl;
            if (q < NC / 2 && f < a.nfun && !(sgn && m == 0)) {
                double* col0 = Vp + (2 * q) * CS;
                const double* Va = col0;
                const double* Vb = col0 + CS;
                const double2 qb2 = __ldg(a.qtab + lane);
                double xr[16], xi[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int kk = f16_in_index(lane, e);
                    // W[k] = e^{i pi k/2n} (Xa[k] - i Xa[n-k]) + i (same for b), X[k >= bw] = 0   (k_dct_inv)
                    double wr = 0.0, wi = 0.0;
                    if (kk != B) {
                        const int src = kk < B ? kk : N - kk;
                        const double sc = (src == 0) ? c_zero : c_rest;
                        const int slot = (src & 1) * PS + (src >> 1);
                        const double va = Va[slot] * sc, vb = Vb[slot] * sc;
                        double qr, qi;
                        f16_quarter_rot(qb2.x, qb2.y, e, qr, qi);
                        const double ur = (kk < B) ? va : vb, ui = (kk < B) ? vb : -va;  // (a + ib) or -i (a + ib)
                        wr = qr * ur - qi * ui;
                        wi = qr * ui + qi * ur;
                    }
                    xr[e] = wi;  // swapped: inverse DFT through the forward transform
                    xi[e] = wr;
                }
                d16_fft512_inplace(xr, xi, col0, PS, lane, a.tw);
                const int mp = sgn ? N - m : m;
                const double sign = (sgn && (m & 1)) ? -out_scale : out_scale;  // (-1)^m for negative orders
                double* Gr = a.G + ((long)f * 2 * N + mp) * N;
                double* Gi = Gr + (long)N * N;
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    const int i = f16_out_index(lane, o);
                    const double s = (m & 1) ? __ldg(a.sinv + i) * sign : sign;
                    const int at = a.lat_perm ? i : ((i < B) ? 2 * i : 2 * (N - 1 - i) + 1);
                    Gr[at] = xi[o] * s;  // Re z -> column a (real part)
                    Gi[at] = xr[o] * s;  // Im z -> column b (imaginary part)
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();  // next coefficient panel landed, V panel free
    }
}

// Opt-in (S2KIT_CUDA_UNI_INV=1): measured at bw = 256, 1024 functions, the fused kernel takes 2.83 ms against 1.17 + 1.19 ms
// for K4 + K5 as separate kernels -- with one V panel its two phases cannot overlap, so it pays both halves in full plus
// two CTA-wide barriers per item, while the separate kernels run at 24 warps per SM (profiles/r2_ncu_uni_summary.md).
static bool uni_inv_enabled() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_UNI_INV");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}

bool inv_uni_supported(const s2kit_cuda_plan* p, int nfun, int data_format) {
    if (!uni_enabled() || !uni_inv_enabled() || !p->fast || p->n != 512 || !p->d_isub_list) return false;
    return nfun * (data_format == S2KIT_REAL ? 2 : 4) >= UNI_NC;
}

cudaError_t launch_inv_uni(s2kit_cuda_plan* p, const double* table_t, uint64_t shift, const double* rco, const double* ico,
                           long coef_stride, double* G, int nfun, int m_lo, int m_hi, int data_format, int lat_perm) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    UniInvArgs a;
    a.table = table_t;
    a.order_start = p->d_order_start;
    a.table_shift = shift;
    a.sub_off = p->d_isub_off;
    a.sub_list = p->d_isub_list;
    a.nlist = p->n_isub_list;
    a.rco = rco;
    a.ico = ico;
    a.coef_stride = coef_stride;
    a.G = G;
    a.sinv = p->d_sv;
    a.tw = p->d_tw_n;
    a.qtab = p->d_q_n;
    a.nfun = nfun;
    a.m_lo = m_lo;
    a.norders = m_hi - m_lo;
    a.real_fmt = data_format == S2KIT_REAL;
    const int NF = UNI_NC / (a.real_fmt ? 2 : 4);
    a.ncoltiles = (nfun + NF - 1) / NF;
    a.lat_perm = lat_perm;
    constexpr int PANEL = 2 * (UNI_NC * 132 + 8);
    const size_t smem = sizeof(double) * 2 * PANEL + sizeof(double2) * UNI_WARPS * UNI_RING * 32 + 64 +
                        uni_meta_bytes(256, p->n_isub_list);
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_inv_uni), smem);
    if (e != cudaSuccess) return e;
    const int nitems = a.norders * a.ncoltiles;
    int slot = prof_begin(p, S2KIT_K_FUSED_INV);
    k_inv_uni<<<nitems < p->sm_count ? nitems : p->sm_count, UNI_THREADS, smem, p->stream>>>(a);
    e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
