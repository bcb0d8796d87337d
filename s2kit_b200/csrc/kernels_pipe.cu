// kernels_pipe.cu -- the batched forward path as ONE persistent, warp-specialised kernel per launch: K2 + K3 fused.
//
//   weights . DCT-II(2bw) . triangular contraction . coefficient placement
//   (DLTSemi, src/legendre_transform/seminaive.c:153-198, inside the m-loops of FSTSemiMemo,
//    src/FST_semi_memo.c:96-108,131-145,175-201)
//
// Why: K2 writes the cosine planes to HBM only for K3 to read them back (4 of 17 MiB per function and direction), and a
// CTA of k_legendre_fwd spends about half of its warp time outside the DMMA loop -- waiting for its panel, for the
// first table tiles of every row tile, in the epilogue (profiles/r1_ncu_pipe_summary.md).  Here one CTA per SM runs for
// the whole launch:
//   * 8 DCT warps (producers) turn the spectral-plane rows of the next work item into its cosine panel directly in shared
//     memory (the panel never goes through HBM), with the loads of the next transform in flight while the current one
//     is computed;
//   * 8 MMA warps (consumers) contract the current panel against the order's table tiles (streamed from L2 into DMMA A
//     fragments) and store the coefficients;
//   * two panel buffers and two mbarriers per buffer (full / empty) decouple them; no CTA-wide barrier in the loop, a
//     consumer warp that runs out of sub-items moves on to the next panel.
// Work items = (order m, column tile of 32 columns = (function, +-m, re/im)), ordered by decreasing cost and dealt
// round-robin to the CTAs, so consecutive CTAs share an order's table through L2.
// Measured (bw = 256, 256 functions per launch): 660 us against 342 + 347 us for K2 + K3 as separate kernels.  The
// kernel is bound by its producers: DFMA and DMMA share the FP64 pipe, the consumers wait for panels a third of their
// time; other role splits (4 + 16, 8 + 16 warps) were slower.  The second half of this file is the unfused alternative,
// K3 alone as a persistent kernel with cp.async-streamed table tiles (S2KIT_CUDA_PIPE=1).
#include <stdlib.h>

#include "s2k_fft.cuh"
#include "s2k_legendre.cuh"

namespace s2k {

constexpr int PIPE_NC = 32;             // panel columns
constexpr int PIPE_MMA_WARPS = 8;       // consumers: warps 0..7
constexpr int PIPE_DCT_THREADS = 256;   // producers: warps 8..15
constexpr int PIPE_THREADS = PIPE_MMA_WARPS * 32 + PIPE_DCT_THREADS;
constexpr int PIPE_STAGES = 2;

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned addr, unsigned parity) {
    unsigned done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (++spins > (1ull << 26)) __trap();  // a lost arrival must not hang the device
    }
}

struct PipeArgs {
    const double* table;
    const uint64_t* order_start;
    uint64_t table_shift;
    const BlockMeta* meta;
    const uint32_t* rt_start;
    const double* S;        // spectral planes [f][part][order row][latitude slot]
    const double* weights;  // 4bw, load order (s2k_host_reordered)
    const double2* tw;
    const double2* qtab;
    double* rco;
    double* ico;
    long coef_stride;
    int nfun, m_lo, norders, ncoltiles, real_fmt, lat_perm;
    const int* order_list;
};

// item index -> (order, first function)
__device__ __forceinline__ void pipe_item(const PipeArgs& a, int t, int NF, int& m, int& f0) {
    const int oi = t / a.ncoltiles, x = t - oi * a.ncoltiles;
    m = a.order_list ? a.order_list[oi] : a.m_lo + oi;
    f0 = x * NF;
}

template <int N>
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_fwd_pipe(const PipeArgs a) {
    constexpr int B = N / 2, T8 = N / 8, FPB = PIPE_DCT_THREADS / T8, NP = fft_padded_len(N), NC = PIPE_NC;
    constexpr int NPAIR = NC / 2;  // one complex FFT serves the re and im column of one (function, sign)
    static_assert(T8 >= 32 && NPAIR % FPB == 0, "FFT groups are whole warps and tile the panel");
    extern __shared__ __align__(16) double smem[];
    const int CS = panel_stride(B), half = B / 2;
    const int panel_doubles = 2 * NC * CS;
    double* panels = smem;                                                                   // [STAGES][2][NC][CS]
    double2* ex = reinterpret_cast<double2*>(smem + PIPE_STAGES * panel_doubles);            // [FPB][NP]
    const unsigned bar0 = static_cast<unsigned>(__cvta_generic_to_shared(ex + FPB * NP));   // full[2], empty[2]
    const int tid = threadIdx.x;
    const int cols_per_fn = a.real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int nitems = a.norders * a.ncoltiles;

    if (tid == 0) {
        for (int s = 0; s < PIPE_STAGES; ++s) {
            mbar_init(bar0 + 8 * s, PIPE_DCT_THREADS);                  // full[s]: every producer thread arrives
            mbar_init(bar0 + 8 * (PIPE_STAGES + s), PIPE_MMA_WARPS);    // empty[s]: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // panel slots beyond the bw/2 cosine indices of a parity are only ever multiplied by zero table padding, but must not
    // hold NaN garbage; the producers never write them
    for (int i = tid; i < PIPE_STAGES * 2 * NC * (CS - half); i += PIPE_THREADS)
        panels[(i / (CS - half)) * CS + half + i % (CS - half)] = 0.0;
    __syncthreads();

    if (tid >= PIPE_MMA_WARPS * 32) {
        // =========================================================================================== producers: DCT
        const int ptid = tid - PIPE_MMA_WARPS * 32, g = ptid / T8, t = ptid % T8;
        double2* sx = ex + g * NP;
        constexpr int ROUNDS = NPAIR / FPB;
        constexpr int R = fft_last_radix(N);
        const double s_all = 1.0 / sqrt(2.0 * (double)N);  // 1/sqrt(2*size), seminaive.c:174
        const int nsteps = ((nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * ROUNDS;

        // raw inputs of the NEXT step, in flight while the current transform is computed
        double nr[8], ni[8], nw[8];
        bool nlive = false;
        auto issue_loads = [&](int step) {
            const int k = step / ROUNDS, round = step % ROUNDS;
            int m, f0;
            pipe_item(a, blockIdx.x + k * gridDim.x, NF, m, f0);
            const int q = round * FPB + g;  // pair index: panel columns 2q (re) and 2q + 1 (im)
            const int fl = a.real_fmt ? q : (q >> 1), sgn = a.real_fmt ? 0 : (q & 1);
            const int f = f0 + fl;
            nlive = (f < a.nfun) && !(sgn && m == 0);
            if (nlive) {
                const int mp = sgn ? N - m : m;
                const double* Sr = a.S + ((long)f * 2 * N + mp) * N;
                const double* Si = Sr + (long)N * N;
                const double* w = a.weights + ((m & 1) ? N : 0);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int p = t + e * T8;
                    const int at = a.lat_perm ? p : ((p < B) ? 2 * p : 2 * (N - 1 - p) + 1);
                    nr[e] = __ldg(Sr + at);
                    ni[e] = __ldg(Si + at);
                    nw[e] = __ldg(w + p);
                }
            }
        };
        if (nsteps > 0) issue_loads(0);
#pragma unroll 1
        for (int step = 0; step < nsteps; ++step) {
            const int k = step / ROUNDS, round = step % ROUNDS, b = k & 1;
            const bool live = nlive;
            double xr[8], xi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                xr[e] = live ? nr[e] * nw[e] : 0.0;
                xi[e] = live ? ni[e] * nw[e] : 0.0;
            }
            if (step + 1 < nsteps) issue_loads(step + 1);
            if (round == 0) {
                // the consumers reach this item's order within microseconds: pull its table tiles into L2 now
                int m, f0;
                pipe_item(a, blockIdx.x + k * gridDim.x, NF, m, f0);
                prefetch_order_l2(a.table + (a.order_start[m] - a.table_shift) * 64,
                                  a.order_start[m + 1] - a.order_start[m], ptid, PIPE_DCT_THREADS, 1u << 20);
                mbar_wait_parity(bar0 + 8 * (PIPE_STAGES + b), ((k >> 1) & 1) ^ 1);  // panel b drained
            }
            if (live) {  // uniform over the transform's threads (whole warps)
                fft_block<N>(xr, xi, sx, t, g, a.tw);
                // Z[k] stays in registers; only the upper half of the spectrum is exchanged (kernels_fft.cu, K2)
                fft_sync<N>(g);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (fft_slot<R>(e) >= 4) sx[fft_pad(fft_out_index<N>(e, t) - B)] = make_double2(xr[e], xi[e]);
                fft_sync<N>(g);
                const int q = round * FPB + g;
                double* x_re = panels + b * panel_doubles + (2 * q) * CS;  // parity 0; parity 1 is NC*CS further
                const double2 q0 = __ldg(a.qtab + t);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (fft_slot<R>(e) >= 4) continue;
                    const int kk = fft_out_index<N>(e, t);
                    const double ar = xr[e], ai = xi[e];
                    double br = ar, bi = ai;  // k = 0: Z[n] = Z[0]
                    if (kk != 0) {
                        const double2 zb = sx[fft_pad(B - kk)];
                        br = zb.x;
                        bi = zb.y;
                    }
                    const double2 qq = quarter_rot(q0, fft_slot<R>(e));  // kk = t + slot n/8
                    double y1 = qq.x * (ar + br) + qq.y * (ai - bi);
                    double y2 = qq.x * (ai + bi) - qq.y * (ar - br);
                    if (kk == 0) {
                        y1 *= 0.70710678118654752440;  // M_SQRT1_2, seminaive.c:173
                        y2 *= 0.70710678118654752440;
                    }
                    double* dst = x_re + (kk & 1) * NC * CS + (kk >> 1);
                    dst[0] = y1 * s_all;
                    dst[CS] = y2 * s_all;
                }
            }
            if (round == ROUNDS - 1) mbar_arrive(bar0 + 8 * b);  // this thread's part of panel b is written
        }
        return;
    }

    // =============================================================================================== consumers: DMMA
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
#pragma unroll 1
    for (int k = 0, item = blockIdx.x; item < nitems; ++k, item += gridDim.x) {
        int m, f0;
        pipe_item(a, item, NF, m, f0);
        const int b = k & 1;
        const BlockMeta mb0 = a.meta[2 * m], mb1 = a.meta[2 * m + 1];
        const double* tbase = a.table + (a.order_start[m] - a.table_shift) * 64 + lane * 2;
        const double* Xs = panels + b * panel_doubles;
        // sub-items: (parity, pair of adjacent row tiles) while that gives every warp work, single row tiles otherwise;
        // heaviest first, snake over the warps, direction alternating from item to item so that a warp with the
        // heavy end of one item gets the light end of the next (warps may drift one panel apart)
        const int unit = mb0.nrt >= 7 ? 2 : 1;
        const int nunit0 = (mb0.nrt + unit - 1) / unit;  // mb0.nrt >= mb1.nrt
        const int slot = (k & 1) ? PIPE_MMA_WARPS - 1 - warp : warp;
        // epilogue addressing (kernels_legendre.cu, K3): the lane's columns 8j + 2 q4 + {0,1}
        const int sgn = a.real_fmt ? 0 : (q4 & 1);
        const int fl0 = a.real_fmt ? q4 : (q4 >> 1), flstep = a.real_fmt ? 4 : 2;
        const long run0 = (long)(f0 + fl0) * a.coef_stride + (sgn ? coef_base(-m, B) : coef_base(m, B));
        const long mrun0 = (long)(f0 + fl0) * a.coef_stride + coef_base(-m, B);
        const unsigned long long flip = (sgn && (m & 1)) ? 0x8000000000000000ull : 0ull;  // FST_semi_memo.c:181-186
        const bool live_sign = !(sgn && m == 0);
        const bool mirror = a.real_fmt && m > 0;  // FST_semi_memo.c:131-145

        mbar_wait_parity(bar0 + 8 * b, (k >> 1) & 1);  // panel b is complete
#pragma unroll 1
        for (int round = 0; round * PIPE_MMA_WARPS < 2 * nunit0; ++round) {
            const int q = snake_item(round, slot, PIPE_MMA_WARPS);
            if (q >= 2 * nunit0) continue;
            const int p = q & 1;
            const BlockMeta mb = p ? mb1 : mb0;
            const int rt1 = mb.nrt - 1 - unit * (q >> 1), rt0 = unit == 2 ? rt1 - 1 : -1;  // rt0 = -1: single tile
            if (rt1 < 0) continue;
            const uint32_t* sr = a.rt_start + mb.rt_base;
            const double* xp = Xs + (p * NC + g) * CS + q4;
            double acc0[NC / 8][2], acc1[NC / 8][2];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
            if (rt0 >= 0)
                fwd_row_tile2<NC, LEG_PF2_FWD>(tbase + (uint64_t)__ldg(sr + rt0) * 64, tiles_in_row(mb, rt0),
                                               tbase + (uint64_t)__ldg(sr + rt1) * 64, tiles_in_row(mb, rt1), xp, CS,
                                               acc0, acc1);
            else
                fwd_row_tile<NC>(tbase + (uint64_t)__ldg(sr + rt1) * 64, xp, CS, tiles_in_row(mb, rt1), acc1);
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
                        const unsigned long long bits =
                            (unsigned long long)__double_as_longlong(h ? acc1[j][e] : acc0[j][e]);
                        double* arr = e ? a.ico : a.rco;
                        arr[run0 + (long)j * flstep * a.coef_stride + off] = __longlong_as_double((long long)(bits ^ flip));
                        if (mirror) {
                            const unsigned long long mflip = ((m & 1) ^ e) ? 0x8000000000000000ull : 0ull;
                            arr[mrun0 + (long)j * flstep * a.coef_stride + off] =
                                __longlong_as_double((long long)(bits ^ mflip));
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (PIPE_STAGES + b));  // this warp is done with panel b
    }
}

// =====================================================================================================================
// K3 as a persistent STREAMED kernel (the panel comes from K2 through HBM).
//
// What starved the DMMA pipe in every earlier version was the table, not the panel: a register ring of N loads does not
// give N tiles of look-ahead (waiting for the oldest load also waits for younger ones that share one of the warp's six
// scoreboards), and every row tile started with an exposed L2 round trip (ncu: the first DMMA of each step on the long
// scoreboard, 36 % of the consumers' time; profiles/r1_ncu_pipe_summary.md).  Here each lane copies the 16 bytes it will
// feed to the DMMAs -- its A fragments of both k-steps of a tile -- into a lane-private slot of a shared-memory ring
// with cp.async, always STR_RING tiles ahead of the tile being multiplied, ACROSS row tiles and work items: the sequence
// of tiles a warp consumes is known in closed form (block_meta_of / row_tile_start_of), so a second cursor walks it
// ahead of the consumer.  Completion is counted with cp.async groups (exact, in order); no other lane ever touches a
// slot, so no barrier is needed.  Panels are staged by four copy warps (cp.async + mbarrier), two buffers.
constexpr int STR_MMA_WARPS = 20;
constexpr int STR_COPY_THREADS = 128;
constexpr int STR_THREADS = STR_MMA_WARPS * 32 + STR_COPY_THREADS;
constexpr int STR_STAGES = 2;
constexpr int STR_RING = 8;  // tiles in flight per warp: 8 x 512 B

struct StreamArgs {
    const double* table;
    const uint64_t* order_start;
    uint64_t table_shift;
    const double* X;  // cosine planes [f][order row][part][cos_slot(k)] written by K2
    double* rco;
    double* ico;
    long coef_stride;
    int bw, nfun, m_lo, norders, ncoltiles, real_fmt;
    const int* order_list;
};

__device__ __forceinline__ void stream_item(const StreamArgs& a, int t, int NF, int& m, int& f0) {
    const int oi = t / a.ncoltiles, x = t - oi * a.ncoltiles;
    m = a.order_list ? a.order_list[oi] : a.m_lo + oi;
    f0 = x * NF;
}

// completion of this thread's earlier cp.async copies counts as one arrival on the mbarrier
__device__ __forceinline__ void cp_async_arrive_on(unsigned mbar_addr) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar_addr) : "memory");
}

// Look-ahead cursor over one warp's tile sequence.  Sub-items of an item: single row tiles (parity p, row tile rt),
// heaviest first, dealt to the warps in snake order whose direction alternates from item to item.
struct StreamCursor {
    const double* tp;  // this lane's 16 bytes of the next tile (tiles of a row tile are 512 bytes apart); valid if rem > 0
    int rem;           // tiles left in the current row tile; 0 = stream exhausted
    int item, kk, round, m;
    const double* tord;  // first tile of order m
};
__device__ __forceinline__ int stream_slot(int kk, int warp) { return (kk & 1) ? STR_MMA_WARPS - 1 - warp : warp; }
// sub-item `round` of warp `warp` in an order with parity blocks mb0 / mb1: false = none
__device__ __forceinline__ bool stream_sub(const BlockMeta& mb0, const BlockMeta& mb1, int round, int slot, int& p,
                                           int& rt) {
    const int q = snake_item(round, slot, STR_MMA_WARPS);
    if (q >= 2 * mb0.nrt) return false;  // mb0.nrt >= mb1.nrt
    p = q & 1;
    rt = (p ? mb1.nrt : mb0.nrt) - 1 - (q >> 1);
    return rt >= 0;
}
__device__ __forceinline__ void stream_seek(StreamCursor& c, const StreamArgs& a, int nitems, int warp, int lane) {
    c.rem = 0;
    for (;;) {
        ++c.round;
        BlockMeta mb0 = block_meta_of(c.m, 0, a.bw);
        if (c.item < 0 || c.round * STR_MMA_WARPS >= 2 * mb0.nrt) {
            c.item = c.item < 0 ? (int)blockIdx.x : c.item + (int)gridDim.x;
            ++c.kk;
            if (c.item >= nitems) return;
            int f0;
            stream_item(a, c.item, 1, c.m, f0);
            c.tord = a.table + (a.order_start[c.m] - a.table_shift) * 64;
            c.round = 0;
            mb0 = block_meta_of(c.m, 0, a.bw);
        }
        const BlockMeta mb1 = block_meta_of(c.m, 1, a.bw);
        int p, rt;
        if (stream_sub(mb0, mb1, c.round, stream_slot(c.kk, warp), p, rt)) {
            const BlockMeta& mb = p ? mb1 : mb0;
            c.tp = c.tord + (uint64_t)((p ? block_tiles_of(mb0) : 0u) + row_tile_start_of(mb, rt)) * 64 + lane * 2;
            c.rem = tiles_in_row(mb, rt);
            return;
        }
    }
}

__global__ void __launch_bounds__(STR_THREADS, 1) k_leg_fwd_stream(const StreamArgs a) {
    constexpr int NC = PIPE_NC;
    extern __shared__ __align__(16) double smem[];
    const int B = a.bw, n = 2 * B, CS = panel_stride(B), half = (B + 1) / 2;
    const int panel_doubles = 2 * NC * CS;
    double* panels = smem;                                                        // [STAGES][2][NC][CS]
    double* after = smem + STR_STAGES * panel_doubles;
    const unsigned bar0 = static_cast<unsigned>(__cvta_generic_to_shared(after));  // full[STAGES], empty[STAGES]
    double2* rings = reinterpret_cast<double2*>(after + 2 * 2 * STR_STAGES);      // [MMA_WARPS][RING][32 lanes]
    const int tid = threadIdx.x;
    const int cols_per_fn = a.real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int nitems = a.norders * a.ncoltiles;

    if (tid == 0) {
        for (int s = 0; s < STR_STAGES; ++s) {
            mbar_init(bar0 + 8 * s, STR_COPY_THREADS);               // full[s]: every copy thread's cp.async arrival
            mbar_init(bar0 + 8 * (STR_STAGES + s), STR_MMA_WARPS);   // empty[s]: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the pad slots [half, CS) of every panel column are only ever multiplied by zero table padding, but must not hold
    // NaN garbage; the copies never write them
    for (int i = tid; i < STR_STAGES * 2 * NC * (CS - half); i += STR_THREADS)
        panels[(i / (CS - half)) * CS + half + i % (CS - half)] = 0.0;
    __syncthreads();

    if (tid >= STR_MMA_WARPS * 32) {
        // ============================================================================================ producers: copy
        // STR_COPY_THREADS / NC threads per panel column; a column is two contiguous runs (even cosine indices, odd
        // ones).  Nothing blocks here except the wait for a drained buffer, so the copies of item k + 1 are in flight
        // while the consumers work on item k.
        constexpr int TPC = STR_COPY_THREADS / NC;
        const int ptid = tid - STR_MMA_WARPS * 32;
        const int col = ptid / TPC, part16 = ptid % TPC;
        const int fl = col / cols_per_fn, sub = col % cols_per_fn;
        const int sgn = a.real_fmt ? 0 : (sub >> 1), part = sub & 1;
#pragma unroll 1
        for (int k = 0, item = blockIdx.x; item < nitems; ++k, item += gridDim.x) {
            int m, f0;
            stream_item(a, item, NF, m, f0);
            const int b = k % STR_STAGES, f = f0 + fl;
            if (ptid < 32)  // the order's tiles into L2 ahead of the consumers
                prefetch_order_l2(a.table + (a.order_start[m] - a.table_shift) * 64,
                                  a.order_start[m + 1] - a.order_start[m], ptid, 32, 1u << 20);
            mbar_wait_parity(bar0 + 8 * (STR_STAGES + b), ((k / STR_STAGES) & 1) ^ 1);  // panel b drained
            if (f < a.nfun && !(sgn && m == 0)) {
                const int mp = sgn ? n - m : m;
                const double* src = a.X + (((long)f * n + mp) * 2 + part) * B;
                double* d0 = panels + b * panel_doubles + col * CS;
                double* d1 = d0 + NC * CS;
                for (int c = 2 * part16; c < half; c += 2 * TPC) {
                    cp_async16(d0 + c, src + c);
                    cp_async16(d1 + c, src + half + c);
                }
            }
            cp_async_arrive_on(bar0 + 8 * b);
        }
        return;
    }

    // ================================================================================================ consumers: DMMA
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    double2* ring = rings + warp * STR_RING * 32 + lane;
    StreamCursor cur;
    cur.item = -1;
    cur.kk = -1;
    cur.round = 0;
    cur.m = 0;
    cur.tord = a.table;
    stream_seek(cur, a, nitems, warp, lane);
    auto refill = [&](double2* slot_ptr) {
        if (cur.rem) {
            cp_async16(reinterpret_cast<double*>(slot_ptr), cur.tp);
            cur.tp += 64;
            if (--cur.rem == 0) stream_seek(cur, a, nitems, warp, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per slot, empty once the stream has ended
    };
#pragma unroll 1
    for (int u = 0; u < STR_RING; ++u) refill(ring + u * 32);
    unsigned seq = 0;

#pragma unroll 1
    for (int k = 0, item = blockIdx.x; item < nitems; ++k, item += gridDim.x) {
        int m, f0;
        stream_item(a, item, NF, m, f0);
        const int b = k % STR_STAGES;
        const BlockMeta mb0 = block_meta_of(m, 0, B), mb1 = block_meta_of(m, 1, B);
        const double* Xs = panels + b * panel_doubles;
        const int slot = stream_slot(k, warp);
        // epilogue addressing (kernels_legendre.cu, K3): the lane's columns 8j + 2 q4 + {0,1}
        const int sgn = a.real_fmt ? 0 : (q4 & 1);
        const int fl0 = a.real_fmt ? q4 : (q4 >> 1), flstep = a.real_fmt ? 4 : 2;
        const long run0 = (long)(f0 + fl0) * a.coef_stride + (sgn ? coef_base(-m, B) : coef_base(m, B));
        const long mrun0 = (long)(f0 + fl0) * a.coef_stride + coef_base(-m, B);
        const unsigned long long flip = (sgn && (m & 1)) ? 0x8000000000000000ull : 0ull;  // FST_semi_memo.c:181-186
        const bool live_sign = !(sgn && m == 0);
        const bool mirror = a.real_fmt && m > 0;  // FST_semi_memo.c:131-145

        mbar_wait_parity(bar0 + 8 * b, (k / STR_STAGES) & 1);  // panel b is complete
#pragma unroll 1
        for (int round = 0; round * STR_MMA_WARPS < 2 * mb0.nrt; ++round) {
            int p, rt;
            if (!stream_sub(mb0, mb1, round, slot, p, rt)) continue;
            const BlockMeta mb = p ? mb1 : mb0;
            const double* xp = Xs + (p * NC + g) * CS + q4;
            const int ctn = tiles_in_row(mb, rt);
            double acc[NC / 8][2];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 1
            for (int ct = 0; ct < ctn; ++ct, ++seq) {
                double2* slot_ptr = ring + (seq & (STR_RING - 1)) * 32;
                asm volatile("cp.async.wait_group %0;" ::"n"(STR_RING - 1) : "memory");  // the oldest slot has landed
                const double2 av = *slot_ptr;
                double bf[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    bf[j][0] = xp[j * 8 * CS + 8 * ct];
                    bf[j][1] = xp[j * 8 * CS + 8 * ct + 4];
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.x, bf[j][0]);
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.y, bf[j][1]);
                refill(slot_ptr);  // the DMMAs above have consumed the slot's registers
            }
            const int r = 8 * rt + g;
            if (r >= mb.rows || !live_sign) continue;
            const int off = p + 2 * r;  // l - m
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
                if (f0 + fl0 + j * flstep >= a.nfun) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(acc[j][e]);
                    double* arr = e ? a.ico : a.rco;
                    arr[run0 + (long)j * flstep * a.coef_stride + off] = __longlong_as_double((long long)(bits ^ flip));
                    if (mirror) {
                        const unsigned long long mflip = ((m & 1) ^ e) ? 0x8000000000000000ull : 0ull;
                        arr[mrun0 + (long)j * flstep * a.coef_stride + off] =
                            __longlong_as_double((long long)(bits ^ mflip));
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (STR_STAGES + b));  // this warp is done with panel b
    }
}

// ------------------------------------------------------------------------------------------------ launcher
static bool pipe_enabled() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_PIPE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on != 0;
}

// batched launches at bw = 128 / 256 whose columns fill at least one 32-column panel
bool fwd_pipe_supported(const s2kit_cuda_plan* p, int nfun, int data_format) {
    if (!pipe_enabled() || !p->fast) return false;
    if (p->n != 256 && p->n != 512) return false;
    return nfun * (data_format == S2KIT_REAL ? 2 : 4) >= PIPE_NC;
}

template <int N>
static cudaError_t fwd_pipe_n(s2kit_cuda_plan* p, const PipeArgs& a) {
    constexpr int FPB = PIPE_DCT_THREADS / (N / 8);
    const int CS = panel_stride(N / 2);
    const size_t smem = sizeof(double) * PIPE_STAGES * 2 * PIPE_NC * CS + sizeof(double2) * FPB * fft_padded_len(N) +
                        8 * 2 * PIPE_STAGES;
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_fwd_pipe<N>), smem);
    if (e != cudaSuccess) return e;
    const int sms = p->sm_count;
    const int nitems = a.norders * a.ncoltiles;
    k_fwd_pipe<N><<<nitems < sms ? nitems : sms, PIPE_THREADS, smem, p->stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fwd_pipe(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* S, double* rco,
                            double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format, int lat_perm) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    PipeArgs a;
    a.table = table;
    a.order_start = p->d_order_start;
    a.table_shift = shift;
    a.meta = p->d_meta;
    a.rt_start = p->d_rt_start;
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
    const int NF = PIPE_NC / (a.real_fmt ? 2 : 4);
    a.ncoltiles = (nfun + NF - 1) / NF;
    a.lat_perm = lat_perm;
    a.order_list = nullptr;
    int slot = prof_begin(p, S2KIT_K_FUSED_FWD);
    cudaError_t e = p->n == 512 ? fwd_pipe_n<512>(p, a) : fwd_pipe_n<256>(p, a);
    prof_end(p, slot);
    return e;
}

// S2KIT_CUDA_PIPE=2: fused K2+K3 (the DCT runs inside the persistent kernel); 1: K2 as its own kernel + streamed K3
bool fwd_pipe_fused() {
    static int fused = [] {
        const char* e = getenv("S2KIT_CUDA_PIPE");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    return fused != 0;
}

cudaError_t launch_leg_fwd_stream(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* X, double* rco,
                                  double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    StreamArgs a;
    a.table = table;
    a.order_start = p->d_order_start;
    a.table_shift = shift;
    a.X = X;
    a.rco = rco;
    a.ico = ico;
    a.coef_stride = coef_stride;
    a.bw = p->bw;
    a.nfun = nfun;
    a.m_lo = m_lo;
    a.norders = m_hi - m_lo;
    a.real_fmt = data_format == S2KIT_REAL;
    const int NF = PIPE_NC / (a.real_fmt ? 2 : 4);
    a.ncoltiles = (nfun + NF - 1) / NF;
    a.order_list = nullptr;
    const size_t smem = sizeof(double) * STR_STAGES * 2 * PIPE_NC * panel_stride(p->bw) + sizeof(double) * 4 * STR_STAGES +
                        sizeof(double2) * STR_MMA_WARPS * STR_RING * 32;
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_leg_fwd_stream), smem);
    if (e != cudaSuccess) return e;
    const int sms = p->sm_count;
    const int nitems = a.norders * a.ncoltiles;
    int slot = prof_begin(p, S2KIT_K_LEGENDRE_FWD);
    k_leg_fwd_stream<<<nitems < sms ? nitems : sms, STR_THREADS, smem, p->stream>>>(a);
    e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
