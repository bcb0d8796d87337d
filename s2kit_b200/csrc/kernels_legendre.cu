// kernels_legendre.cu -- K3 / K4: the seminaive Legendre contraction on FP64 tensor cores (DMMA).
//
// K3 replaces the triangular dot of DLTSemi     src/legendre_transform/seminaive.c:183-197
//    plus the coefficient placement / (-1)^m / REAL-format symmetry of FSTSemiMemo
//                                                src/FST_semi_memo.c:96-108,131-145,175-201
// K4 replaces the transposed dot of InvDLTSemi  src/legendre_transform/seminaive.c:74-95
//
// The batched contraction is a real dense GEMM per order m: coefficients[l, col] = sum_k T_m[l,k] X[k,col]
// with columns = (function, +m / -m, re / im).  T_m is checkerboard-sparse, so it is handled as two
// lower-trapezoidal parity blocks (s2k_internal.cuh), tiled 8x8 in DMMA fragment order: the table streams
// from L2/HBM straight into mma.sync.m8n8k4.f64 A fragments with one coalesced 128-bit load per lane and
// tile, never touching shared memory; the X (or coefficient) panel of the CTA's columns sits in shared
// memory for the whole order.  The inverse reads the SAME tiles as B fragments (4 rows x 8 columns), so
// Memo plans at bw >= 512 store ONE table copy (the inverse gathers its B fragments from the A-order tiles with two 8-byte
// copies per lane); smaller bandwidths also keep a tile-transposed copy for the wide batched kernels, 26 MB at bw = 256
// (the reference always keeps a transposed table: cospml.c:301-362).
#include <stdlib.h>

#include "s2k_legendre.cuh"

namespace s2k {

// destination of one panel column in the epilogues
struct ColOut {
    double* dst;     // element (l = m) of the column's output run, nullptr = dead column
    double* mirror;  // K3, REAL format: the (-m) copy
    double scale, mscale;
};

// Stores the degrees l - m = 2r (v0) and 2r + 1 (v1) of one column.  `run` points at the column's element l = m.
// If that element is 16-byte aligned the pair (2r, 2r+1) is one 128-bit store; otherwise the aligned pairs are
// (2r+1, 2r+2) = (v1 of this row, v0 of the next row, which sits 4 lanes up), with the tile's first v0 stored alone.
__device__ __forceinline__ void store_pair(double* run, double scale, int r, int g, double v0, double v1, double n0,
                                           bool v0ok, bool v1ok, bool n0ok) {
    if (!run) return;
    if ((reinterpret_cast<uintptr_t>(run) & 8) == 0) {
        if (v1ok)
            *reinterpret_cast<double2*>(run + 2 * r) = make_double2(v0 * scale, v1 * scale);
        else if (v0ok)
            run[2 * r] = v0 * scale;
    } else {
        if (g == 0 && v0ok) run[2 * r] = v0 * scale;
        if (v1ok) {
            if (n0ok)
                *reinterpret_cast<double2*>(run + 2 * r + 1) = make_double2(v1 * scale, n0 * scale);
            else
                run[2 * r + 1] = v1 * scale;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3
// grid: x = column tile, y = order (heavy orders first), z = row split.  NC = MMA columns per CTA (multiple of 8),
// PC = columns actually stored in the panel: PC = 4 < NC = 8 is the single-field case (re/im x +-m), where the
// other half of the 8-wide MMA tile is fed zeros from registers so the panel -- and with it the shared memory per
// CTA -- halves and three CTAs fit an SM even at bw = 2048.
template <int NC, int PC>
__global__ void __launch_bounds__(LEG_WARPS * 32, NC >= 32 ? 2 : 3) k_legendre_fwd(
    const double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t table_shift,
    const BlockMeta* __restrict__ meta, const uint32_t* __restrict__ rt_start, const double* __restrict__ X,
    double* __restrict__ rco, double* __restrict__ ico, long coef_stride, int bw, int nfun, int m_lo, int real_fmt,
    const int* __restrict__ order_list, unsigned l2pf_cap, int tma_tables) {
    extern __shared__ double smem[];
    const int n = 2 * bw, CS = panel_stride(bw);
    const int m = order_list ? order_list[blockIdx.y] : m_lo + blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cols_per_fn = real_fmt ? 2 : 4;
    const int NF = PC / cols_per_fn;
    const int f0 = blockIdx.x * NF;
    double* Xs = smem;  // [2][PC][CS]

    prefetch_order_l2(table + (order_start[m] - table_shift) * 64, order_start[m + 1] - order_start[m], tid, blockDim.x,
                      l2pf_cap);
    // block metadata first: the loads fly while the panel is staged
    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    // ---- stage the X panel, de-interleaved by cosine-index parity; dead columns and the pad slots are zero
    const int half = (bw + 1) / 2;
    for (int col = warp; col < PC; col += LEG_WARPS) {
        int fl = col / cols_per_fn, sub = col % cols_per_fn;
        int sgn = real_fmt ? 0 : (sub >> 1), part = sub & 1;
        int f = f0 + fl;
        double* d0 = Xs + col * CS;
        double* d1 = Xs + (PC + col) * CS;
        if (f >= nfun || (sgn && m == 0)) {
            for (int c = lane; c < half; c += 32) d0[c] = d1[c] = 0.0;
            continue;
        }
        int mp = sgn ? n - m : m;
        const double* src = X + (((long)f * n + mp) * 2 + part) * bw;
        // the plane stores even cosine indices first, then odd ones (cos_slot): two contiguous runs per column
        if ((bw & 3) == 0) {
            for (int c = 2 * lane; c < half; c += 64) {
                cp_async16(d0 + c, src + c);
                cp_async16(d1 + c, src + half + c);
            }
        } else {
            for (int c = lane; c < half; c += 32) cp_async8(d0 + c, src + c);
            for (int c = lane; c < bw / 2; c += 32) cp_async8(d1 + c, src + half + c);
        }
        if ((bw & 1) && lane == 0) d1[half - 1] = 0.0;  // odd bw: parity 1 has one entry less
    }
    {
        // pad slots [half, CS) of every panel column
        const int padw = CS - half;
        for (int i = tid; i < 2 * PC * padw; i += blockDim.x) Xs[(i / padw) * CS + half + i % padw] = 0.0;
    }
    const int total = mb0.nrt + mb1.nrt;
    uint32_t* srt = reinterpret_cast<uint32_t*>(Xs + 2 * PC * CS);  // row-tile starts of both parity blocks
    ColOut* cinfo = reinterpret_cast<ColOut*>(srt + ((bw / 8 + 8 + 3) & ~3));
    // narrow panels: lane-private table-tile ring [LEG_WARPS][LEG_RING][32] behind the column table (s2k_legendre.cuh)
    double2* ring = reinterpret_cast<double2*>(cinfo + NC) + (warp * LEG_RING * 32 + lane);
    // TMA-staged tiles (narrow panels): BULK_NS mbarriers per warp behind the rings
    const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<double2*>(cinfo + NC) +
                                                                        LEG_WARPS * LEG_RING * 32)) + warp * BULK_NS_MAX * 8;
    unsigned phases = 0;
    if (NC < 32 && tma_tables && lane == 0) {
        for (int s = 0; s < BULK_NS_MAX; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NC < 32 && tid < NC) {
        // where column `tid` of the panel lands: f^(+-m, l) of function f, re or im array, with its sign
        ColOut co = {nullptr, nullptr, 1.0, 1.0};
        const int fl = tid / cols_per_fn, sub = tid % cols_per_fn, f = f0 + fl;
        if (tid < PC && f < nfun) {
            const int part = sub & 1, sgn = real_fmt ? 0 : (sub >> 1);
            double* arr = (part ? ico : rco) + (long)f * coef_stride;
            const double sneg = (m & 1) ? -1.0 : 1.0;
            if (!sgn) {
                co.dst = arr + coef_base(m, bw);
                if (real_fmt && m > 0) {  // f^(-m,l) = (-1)^m conj f^(m,l)   (FST_semi_memo.c:131-145)
                    co.mirror = arr + coef_base(-m, bw);
                    co.mscale = part ? -sneg : sneg;
                }
            } else if (m > 0) {  // (-1)^m' on the negative orders  (FST_semi_memo.c:181-186)
                co.dst = arr + coef_base(-m, bw);
                co.scale = sneg;
            }
        }
        cinfo[tid] = co;
    }
    for (int i = tid; i < total; i += blockDim.x)
        srt[i] = rt_start[(i < mb0.nrt ? mb0.rt_base : mb1.rt_base - mb0.nrt) + i];
    cp_async_wait_all();
    __syncthreads();
    const double* tbase = table + (order_start[m] - table_shift) * 64 + lane * 2;
    const int g = lane >> 2, q4 = lane & 3;
    if constexpr (NC >= 32) {
        // Batched panels: one item = two adjacent row tiles (rt, rt - 1) of one parity block, which share every panel
        // fragment (fwd_row_tile2); an odd block's lightest tile rides alone.  Items sorted by decreasing cost --
        // (parity 0, pair i), (parity 1, pair i) from the longest rows down -- and dealt in snake order.
        const int nslots = LEG_WARPS * gridDim.z, slot = blockIdx.z * LEG_WARPS + warp;
        const int npair0 = (mb0.nrt + 1) / 2;  // mb0.nrt >= mb1.nrt
        for (int round = 0; round * nslots < 2 * npair0; ++round) {
            const int q = snake_item(round, slot, nslots);
            if (q >= 2 * npair0) continue;
            const int p = q & 1;
            const BlockMeta mb = p ? mb1 : mb0;
            const int rt1 = mb.nrt - 1 - 2 * (q >> 1), rt0 = rt1 - 1;  // rt0 = -1: single tile
            if (rt1 < 0) continue;
            const uint32_t* sr = srt + (p ? mb0.nrt : 0);
            const double* xp = Xs + (p * PC + g) * CS + q4;
            double acc0[NC / 8][2], acc1[NC / 8][2];
    #pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
            if (rt0 >= 0)
                fwd_row_tile2<NC, LEG_PF2_FWD>(tbase + (uint64_t)sr[rt0] * 64, tiles_in_row(mb, rt0), tbase + (uint64_t)sr[rt1] * 64,
                                  tiles_in_row(mb, rt1), xp, CS, acc0, acc1);
            else
                fwd_row_tile<NC>(tbase + (uint64_t)sr[rt1] * 64, xp, CS, tiles_in_row(mb, rt1), acc1);

            // ---- epilogue: lane holds rows 8 rt + g of both tiles, columns 8j + 2 q4 + {0,1}.  Column c of the panel is
            // (function f0 + c / cpf, sign, part = c & 1), so the lane's eight destinations differ only by the function
            // and the re / im array: closed form, no per-column table (its shared-memory reads cost more wavefronts
            // than the main loop's fragments, profiles/r1_ncu_pipe_summary.md).  The signs are +-1: flip the sign bit on the
            // integer pipe instead of a DMUL that competes with the DMMAs.
            const int sgn = real_fmt ? 0 : (q4 & 1);
            const int fl0 = real_fmt ? q4 : (q4 >> 1), flstep = real_fmt ? 4 : 2;
            const long run0 = (long)(f0 + fl0) * coef_stride + (sgn ? coef_base(-m, bw) : coef_base(m, bw));
            const long mrun0 = (long)(f0 + fl0) * coef_stride + coef_base(-m, bw);
            const unsigned long long flip = (sgn && (m & 1)) ? 0x8000000000000000ull : 0ull;  // (-1)^m, FST_semi_memo.c:181-186
            const bool live_sign = !(sgn && m == 0);
            const bool mirror = real_fmt && m > 0;  // f^(-m,l) = (-1)^m conj f^(m,l), FST_semi_memo.c:131-145
    #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 8 * (h ? rt1 : rt0) + g;
                if ((h == 0 && rt0 < 0) || r >= mb.rows || !live_sign) continue;
                const int off = p + 2 * r;  // l - m
    #pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    if (f0 + fl0 + j * flstep >= nfun) continue;
    #pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned long long bits =
                            (unsigned long long)__double_as_longlong(h ? acc1[j][e] : acc0[j][e]);
                        double* arr = e ? ico : rco;
                        arr[run0 + (long)j * flstep * coef_stride + off] = __longlong_as_double((long long)(bits ^ flip));
                        if (mirror) {
                            const unsigned long long mflip = ((m & 1) ^ e) ? 0x8000000000000000ull : 0ull;
                            arr[mrun0 + (long)j * flstep * coef_stride + off] =
                                __longlong_as_double((long long)(bits ^ mflip));
                        }
                    }
                }
            }
        }
    } else {
        // One work item = row tile rt of BOTH parity blocks (16 consecutive degrees l), heaviest first, snake over the
        // slots.  Keeping the two parities in one warp lets the epilogue store 16-byte pairs (l, l+1): the scattered
        // 8-byte stores of one parity at a time touched every 32-byte sector twice and cost 20 % of the kernel.
        const int nslots = LEG_WARPS * gridDim.z, slot = blockIdx.z * LEG_WARPS + warp;
        for (int round = 0; round * nslots < mb0.nrt; ++round) {
            const int q = snake_item(round, slot, nslots);
            if (q >= mb0.nrt) continue;
            const int rt = mb0.nrt - 1 - q;  // mb0.nrt >= mb1.nrt
            const int gsel = PC < NC ? (g & (PC - 1)) : g;
            double acc0[NC / 8][2], acc1[NC / 8][2];
    #pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
            if (tma_tables) {
                double2* wring = ring - lane;
                const double* t0 = tbase - lane * 2 + (uint64_t)srt[rt] * 64;
                const double* t1 = tbase - lane * 2 + (uint64_t)srt[mb0.nrt + (rt < mb1.nrt ? rt : 0)] * 64;
                const double* x0 = Xs + gsel * CS + q4;
                const double* x1 = Xs + (PC + gsel) * CS + q4;
                const bool dead = PC < NC && g >= PC;
                const int c0 = tiles_in_row(mb0, rt), c1 = rt < mb1.nrt ? tiles_in_row(mb1, rt) : 0;
                // copies of 1, 2 or 4 tiles (S2KIT_CUDA_TMA_TABLES = 3, 1, 2), always 8 tiles in flight per warp
                if (tma_tables == 2) {
                    fwd_row_tile_bulk<NC, 4, 2>(t0, x0, CS, c0, acc0, dead, wring, bar, phases, lane);
                    if (c1) fwd_row_tile_bulk<NC, 4, 2>(t1, x1, CS, c1, acc1, dead, wring, bar, phases, lane);
                } else if (tma_tables == 3) {
                    fwd_row_tile_bulk<NC, 1, 8>(t0, x0, CS, c0, acc0, dead, wring, bar, phases, lane);
                    if (c1) fwd_row_tile_bulk<NC, 1, 8>(t1, x1, CS, c1, acc1, dead, wring, bar, phases, lane);
                } else {
                    fwd_row_tile_bulk<NC, 2, 4>(t0, x0, CS, c0, acc0, dead, wring, bar, phases, lane);
                    if (c1) fwd_row_tile_bulk<NC, 2, 4>(t1, x1, CS, c1, acc1, dead, wring, bar, phases, lane);
                }
            } else {
                fwd_row_tile_async<NC>(tbase + (uint64_t)srt[rt] * 64, Xs + gsel * CS + q4, CS, tiles_in_row(mb0, rt), acc0,
                                       PC < NC && g >= PC, ring);
                if (rt < mb1.nrt)
                    fwd_row_tile_async<NC>(tbase + (uint64_t)srt[mb0.nrt + rt] * 64, Xs + (PC + gsel) * CS + q4, CS,
                                           tiles_in_row(mb1, rt), acc1, PC < NC && g >= PC, ring);
            }

            // ---- epilogue: lane holds row r = 8rt + g of both parities = degrees l - m = 2r, 2r+1, columns 8j + 2 q4 + {0,1}
            const int r = 8 * rt + g;
            const bool v0ok = r < mb0.rows, v1ok = r < mb1.rows;
    #pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
    #pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const ColOut co = cinfo[8 * j + 2 * q4 + e];
                    const double v0 = acc0[j][e], v1 = acc1[j][e];
                    // parity-0 value of the next row (lane + 4): partner of v1 when the run starts at an odd element
                    const double n0 = __shfl_down_sync(0xffffffffu, v0, 4);
                    const bool n0ok = g < 7 && r + 1 < mb0.rows;
                    store_pair(co.dst, co.scale, r, g, v0, v1, n0, v0ok, v1ok, n0ok);
                    store_pair(co.mirror, co.mscale, r, g, v0, v1, n0, v0ok, v1ok, n0ok);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4
// V[col, k] = sum_l T_m[l, k] c[l, col]: D(8 cols x 8 k) += A(8 cols x 4 l) * B(4 l x 8 k)
template <int NC, int PC>
__global__ void __launch_bounds__(LEG_WARPS * 32, 3) k_legendre_inv(
    const double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t table_shift,
    const BlockMeta* __restrict__ meta, const uint32_t* __restrict__ rt_start, const double* __restrict__ rco,
    const double* __restrict__ ico, long coef_stride, double* __restrict__ V, int bw, int nfun, int m_lo,
    int real_fmt, const int* __restrict__ order_list, unsigned l2pf_cap, int a_order) {
    extern __shared__ double smem[];
    const int n = 2 * bw, CS = panel_stride(bw);
    const int m = order_list ? order_list[blockIdx.y] : m_lo + blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cols_per_fn = real_fmt ? 2 : 4;
    const int NF = PC / cols_per_fn;
    const int f0 = blockIdx.x * NF;
    double* Cs = smem;  // [2][PC][CS], row index r = (l-m)>>1

    prefetch_order_l2(table + (order_start[m] - table_shift) * 64, order_start[m + 1] - order_start[m], tid, blockDim.x,
                      l2pf_cap);
    const int base_pos = coef_base(m, bw), base_neg = coef_base(-m, bw);
    for (int col = warp; col < PC; col += LEG_WARPS) {
        int fl = col / cols_per_fn, sub = col % cols_per_fn;
        int sgn = real_fmt ? 0 : (sub >> 1), part = sub & 1;
        int f = f0 + fl;
        double* d0 = Cs + col * CS;
        double* d1 = Cs + (PC + col) * CS;
        // The contraction reads panel rows < 8 ceil(h / 8) only (whole row tiles): just the (at most 7) rows between a
        // column's last degree and the end of its last row tile have to be cleared -- they meet zero table padding, which
        // must not see NaN garbage.  (Clearing up to CS cost 10 % of the kernel's samples at high orders: short columns,
        // long tails -- profiles/r2_dram_traffic.json capture, source view.)
        const int cnt = bw - m, h0 = (cnt + 1) / 2, h1 = cnt / 2;  // entries of parity 0 / 1
        const int e0 = (h0 + 7) & ~7, e1 = (h1 + 7) & ~7;
        if (f >= nfun || (sgn && m == 0)) {
            for (int c = lane; c < e0; c += 32) d0[c] = d1[c] = 0.0;
            continue;
        }
        const double* src = (part ? ico : rco) + (long)f * coef_stride + (sgn ? base_neg : base_pos);
        for (int o = lane; o < cnt; o += 32) cp_async8(((o & 1) ? d1 : d0) + (o >> 1), src + o);
        if (h0 + lane < e0) d0[h0 + lane] = 0.0;
        if (h1 + lane < e1) d1[h1 + lane] = 0.0;
    }
    cp_async_wait_all();
    __syncthreads();

    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    const int nct = (((bw + 1) / 2) + 7) >> 3;  // column tiles needed to cover every k < bw of one parity
    uint32_t* srt = reinterpret_cast<uint32_t*>(Cs + 2 * PC * CS);
    ColOut* cinfo = reinterpret_cast<ColOut*>(srt + ((bw / 8 + 8 + 3) & ~3));
    double2* ring = reinterpret_cast<double2*>(cinfo + NC) + (warp * LEG_RING * 32 + lane);  // narrow panels only
    if (tid < NC) {
        ColOut co = {nullptr, nullptr, 1.0, 1.0};
        const int fl = tid / cols_per_fn, sub = tid % cols_per_fn, f = f0 + fl;
        const int sgn = real_fmt ? 0 : (sub >> 1), part = sub & 1;
        if (tid < PC && f < nfun && !(sgn && m == 0))
            co.dst = V + (((long)f * n + (sgn ? n - m : m)) * 2 + part) * bw;  // row of the cosine plane
        cinfo[tid] = co;
    }
    for (int i = tid; i < mb0.nrt + mb1.nrt; i += blockDim.x)
        srt[i] = rt_start[(i < mb0.nrt ? mb0.rt_base : mb1.rt_base - mb0.nrt) + i];
    __syncthreads();
    const double* tbase = table + (order_start[m] - table_shift) * 64 + lane * 2;  // B-fragment-ordered tiles
    const int g = lane >> 2, q4 = lane & 3;

    const int nslots = LEG_WARPS * gridDim.z, slot = blockIdx.z * LEG_WARPS + warp;
    if constexpr (NC >= 32) {
        // Batched panels: one item = column tiles (2i, 2i + 1) of one parity block, which share every coefficient
        // fragment (inv_col_tile2); low column tiles (most rows) first.
        const int npair = (nct + 1) / 2;
        for (int round = 0; round * nslots < 2 * npair; ++round) {
            const int q = snake_item(round, slot, nslots);
            if (q >= 2 * npair) continue;
            const int p = q & 1, ct = 2 * (q >> 1);
            const BlockMeta mb = p ? mb1 : mb0;
            double acc0[NC / 8][2], acc1[NC / 8][2];
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
            inv_col_tile2<NC, LEG_PF2_INV>(tbase, srt + (p ? mb0.nrt : 0), mb, ct, Cs + (p * PC + g) * CS + q4, CS, acc0, acc1);
            // ---- epilogue: lane holds column 8j + g, cosine slots c = 8 (ct + h) + 2 q4 + {0,1} of parity p, adjacent in
            // the parity-split plane
            const int hp = p ? bw / 2 : (bw + 1) / 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = 8 * (ct + h) + 2 * q4;
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    double* dst = cinfo[8 * j + g].dst;
                    if (!dst) continue;
                    double* d = dst + p * ((bw + 1) / 2) + c0;
                    const double v0 = h ? acc1[j][0] : acc0[j][0], v1 = h ? acc1[j][1] : acc0[j][1];
                    if (c0 + 1 < hp && ((bw & 3) == 0)) {
                        *reinterpret_cast<double2*>(d) = make_double2(v0, v1);
                    } else {
                        if (c0 < hp) d[0] = v0;
                        if (c0 + 1 < hp) d[1] = v1;
                    }
                }
            }
        }
        return;
    }
    for (int round = 0; round * nslots < 2 * nct; ++round) {
        const int q = snake_item(round, slot, nslots);
        if (q >= 2 * nct) continue;
        const int p = q & 1, ct = q >> 1;  // low column tiles (most rows) first
        const BlockMeta mb = p ? mb1 : mb0;
        double acc[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) acc[j][0] = acc[j][1] = 0.0;
        inv_col_tile_async<NC>(tbase - lane * 2, srt + (p ? mb0.nrt : 0), mb, ct,
                               Cs + (p * PC + (PC < NC ? (g & (PC - 1)) : g)) * CS + q4, CS, acc, PC < NC && g >= PC, ring,
                               a_order, lane);
        // ---- epilogue: lane holds column 8j + g, cosine slots c = 8ct + 2 q4 + {0,1}
        // slots c0, c0+1 of parity p are adjacent in the parity-split plane
        const int c0 = 8 * ct + 2 * q4, hp = p ? bw / 2 : (bw + 1) / 2;
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            double* dst = cinfo[8 * j + g].dst;
            if (!dst) continue;
            double* d = dst + p * ((bw + 1) / 2) + c0;
            if (c0 + 1 < hp && ((bw & 3) == 0)) {
                *reinterpret_cast<double2*>(d) = make_double2(acc[j][0], acc[j][1]);
            } else {
                if (c0 < hp) d[0] = acc[j][0];
                if (c0 + 1 < hp) d[1] = acc[j][1];
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------ K4, quad units
// The batched contraction with the inner loop of k_inv_flow (kernels_flow.cu) inside the ordinary one-CTA-per-(order, 32
// columns) grid: a warp owns FOUR adjacent column tiles of one parity block (16 accumulator fragments, 32 DMMAs per
// row-tile step, every coefficient fragment feeds four DMMAs instead of two) and its four table tiles of a step -- 2 KB,
// contiguous -- arrive by ONE cp.async.bulk copy (TMA) into a two-stage ring.  Eight units per CTA = eight warps; the
// units are dealt so that the two warps of a sub-partition get a long and a short one.
constexpr int LQ = 4, LQ_STAGES = 2;

__global__ void __launch_bounds__(LEG_WARPS * 32, 2) k_legendre_inv_q(
    const double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t table_shift,
    const double* __restrict__ rco, const double* __restrict__ ico, long coef_stride, double* __restrict__ V, int bw, int nfun,
    int m_lo, int real_fmt, unsigned l2pf_cap) {
    constexpr int NC = 32;
    extern __shared__ __align__(128) double smem[];
    const int n = 2 * bw, CS = panel_stride(bw);
    const int m = m_lo + blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const int cols_per_fn = real_fmt ? 2 : 4, NF = NC / cols_per_fn;
    const int f0 = blockIdx.x * NF;
    double* Cs = smem;  // [2][NC][CS], row index r = (l-m)>>1
    double2* ring = reinterpret_cast<double2*>(Cs + 2 * NC * CS) + warp * LQ_STAGES * LQ * 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double2*>(Cs + 2 * NC * CS) + LEG_WARPS * LQ_STAGES * LQ * 32);
    const unsigned bar0 = static_cast<unsigned>(__cvta_generic_to_shared(bars + warp * LQ_STAGES));
    if (lane == 0) {
        for (int s = 0; s < LQ_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    prefetch_order_l2(table + (order_start[m] - table_shift) * 64, order_start[m + 1] - order_start[m], tid, blockDim.x, l2pf_cap);
    const int base_pos = coef_base(m, bw), base_neg = coef_base(-m, bw);
    for (int col = warp; col < NC; col += LEG_WARPS) {
        const int fl = col / cols_per_fn, sub = col % cols_per_fn;
        const int sgn = real_fmt ? 0 : (sub >> 1), part = sub & 1, f = f0 + fl;
        double* d0 = Cs + col * CS;
        double* d1 = Cs + (NC + col) * CS;
        const int cnt = bw - m, h0 = (cnt + 1) / 2, h1 = cnt / 2;  // entries of parity 0 / 1
        const int e0 = (h0 + 7) & ~7, e1 = (h1 + 7) & ~7;
        if (f >= nfun || (sgn && m == 0)) {
            for (int c = lane; c < e0; c += 32) d0[c] = d1[c] = 0.0;
            continue;
        }
        const double* src = (part ? ico : rco) + (long)f * coef_stride + (sgn ? base_neg : base_pos);
        for (int o = lane; o < cnt; o += 32) cp_async8(((o & 1) ? d1 : d0) + (o >> 1), src + o);
        if (h0 + lane < e0) d0[h0 + lane] = 0.0;
        if (h1 + lane < e1) d1[h1 + lane] = 0.0;
    }
    cp_async_wait_all();
    __syncthreads();

    // unit of this warp: parity p, column tiles ct0 .. ct0 + 3.  Quads 0 (most row tiles) .. 3 (fewest) dealt as 0 0 1 1 3 3 2 2:
    // warps w and w + 4 share a sub-partition and get quads (0, 3) or (1, 2)
    const int p = warp & 1;
    const int quad = (0x22331100 >> (4 * warp)) & 0xf;
    const int nct = (((bw + 1) / 2) + 7) >> 3;
    const int ct0 = LQ * quad;
    if (ct0 >= nct) return;
    const BlockMeta mb0 = block_meta_of(m, 0, bw);
    const BlockMeta mb = p ? block_meta_of(m, 1, bw) : mb0;
    const double* tblk = table + ((order_start[m] - table_shift) + (p ? block_tiles_of(mb0) : 0u)) * 64;
    const int rt_min = first_row_tile_reaching(mb, ct0);
    const int cnt = mb.nrt - rt_min;
    double acc[LQ][NC / 8][2];
#pragma unroll
    for (int c = 0; c < LQ; ++c)
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) acc[c][j][0] = acc[c][j][1] = 0.0;
    const double* cp = Cs + (p * NC + g) * CS + q4;
    auto tiles_at = [&](int rt) {
        const int t = tiles_in_row(mb, rt) - ct0;
        return t < LQ ? t : LQ;
    };
    auto issue = [&](int i) {
        const int rt = rt_min + i, s = i % LQ_STAGES;
        bulk_tile_copy(ring + s * LQ * 32, tblk + ((uint64_t)row_tile_start_of(mb, rt) + ct0) * 64, 512u * (unsigned)tiles_at(rt),
                       bar0 + 8 * s);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < LQ_STAGES; ++s)
            if (s < cnt) issue(s);
    }
    unsigned phases = 0;
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) {
        const int rt = rt_min + i, s = i % LQ_STAGES;
        const int nv = tiles_at(rt);
        double av[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            av[j][0] = cp[j * 8 * CS + 8 * rt];
            av[j][1] = cp[j * 8 * CS + 8 * rt + 4];
        }
        bulk_wait(bar0 + 8 * s, (phases >> s) & 1u);
        phases ^= 1u << s;
        double2 bv[LQ];
#pragma unroll
        for (int c = 0; c < LQ; ++c) bv[c] = ring[(s * LQ + c) * 32 + lane];
        if (nv == LQ) {
#pragma unroll
            for (int c = 0; c < LQ; ++c)
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][0], bv[c].x);
#pragma unroll
            for (int c = 0; c < LQ; ++c)
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][1], bv[c].y);
        } else {
#pragma unroll
            for (int c = 0; c < LQ - 1; ++c)
                if (c < nv) {
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][0], bv[c].x);
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc[c][j], av[j][1], bv[c].y);
                }
        }
        if (i + LQ_STAGES < cnt) {
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(i + LQ_STAGES);
            }
        }
    }
    // epilogue: lane holds column 8j + g, cosine slots 8 (ct0 + c) + 2 q4 + {0, 1} of parity p
    const int sgn = real_fmt ? 0 : ((g >> 1) & 1), part = g & 1;
    const int fl0 = real_fmt ? (g >> 1) : (g >> 2), flstep = real_fmt ? 4 : 2;
    if (sgn && m == 0) return;
    const int mp = sgn ? n - m : m, hp = p ? bw / 2 : (bw + 1) / 2;
#pragma unroll
    for (int j = 0; j < NC / 8; ++j) {
        const int f = f0 + fl0 + j * flstep;
        if (f >= nfun) continue;
        double* d = V + (((long)f * n + mp) * 2 + part) * bw + p * ((bw + 1) / 2) + 8 * ct0 + 2 * q4;
#pragma unroll
        for (int c = 0; c < LQ; ++c) {
            const int c0 = 8 * (ct0 + c) + 2 * q4;
            if (c0 + 1 < hp) *reinterpret_cast<double2*>(d + 8 * c) = make_double2(acc[c][j][0], acc[c][j][1]);
        }
    }
}

// Orders up to this many bytes are pulled into L2 by one bulk prefetch at CTA start.  Larger orders (single large-bw
// fields) are streamed by the register ring alone: with hundreds of CTAs in flight a whole-order prefetch of
// megabytes each overruns L2 and the data is fetched twice.
static unsigned l2_prefetch_cap(int panel_cols) {
    static long cap = [] {
        const char* e = getenv("S2KIT_CUDA_L2PF_MAX");
        return e ? (long)strtoul(e, nullptr, 10) : -1L;
    }();
    if (cap >= 0) return (unsigned)cap;
    // measured (profiles/r1_ncu_summary.md): the prefetch pays when 32-column CTAs of a batch share an order; a
    // single field streams every tile exactly once and is faster without it (bw 512: 68 vs 84 us, bw 2048: 3.2 vs 4.0 ms)
    return panel_cols >= 16 ? (1u << 20) : 0u;
}

// ------------------------------------------------------------------------------------------------ launchers
template <int NC, int PC = NC>
static cudaError_t leg_fwd_nc(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* X, double* rco,
                              double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int real_fmt, int rowsplit,
                              const int* order_list) {
    int cols_per_fn = real_fmt ? 2 : 4;
    int NF = PC / cols_per_fn;
    size_t smem = sizeof(double) * 2 * PC * panel_stride(p->bw) + sizeof(uint32_t) * ((p->bw / 8 + 8 + 3) & ~3) +
                  sizeof(ColOut) * NC + (NC < 32 ? sizeof(double2) * LEG_WARPS * LEG_RING * 32 + 8 * LEG_WARPS * BULK_NS_MAX : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_legendre_fwd<NC, PC>), smem);
        if (e != cudaSuccess) return e;
    }
    // Narrow panels (single fields stream every tile once).  Default: the lane-private cp.async ring.  S2KIT_CUDA_TMA_TABLES
    // = 1 / 2 / 3 stages the tiles with TMA bulk copies of 2 / 4 / 1 tiles instead (fwd_row_tile_bulk): measured slower --
    // bw = 2048 forward 2.27 ms (cp.async) vs 2.34 / 2.30 / 2.94 ms, bw = 1024 0.36 vs 0.47 ms (profiles/r2_single_field_variants.json):
    // the ring already runs the table stream at the HBM roofline (6.6 TB/s), and one elected lane issuing the copies and
    // arming the mbarriers per chunk costs more than 32 lanes each issuing their own 16-byte copy.
    static const int tma_tables = [] {
        const char* e = getenv("S2KIT_CUDA_TMA_TABLES");
        return e ? atoi(e) : 0;
    }();
    dim3 grid((nfun + NF - 1) / NF, m_hi - m_lo, rowsplit);
    k_legendre_fwd<NC, PC><<<grid, LEG_WARPS * 32, smem, p->stream>>>(table, p->d_order_start, shift, p->d_meta,
                                                                  p->d_rt_start, X, rco, ico, coef_stride, p->bw, nfun,
                                                                  m_lo, real_fmt, order_list, l2_prefetch_cap(PC),
                                                                  NC < 32 ? tma_tables : 0);
    return cudaGetLastError();
}

template <int NC, int PC = NC>
static cudaError_t leg_inv_nc(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* rco,
                              const double* ico, long coef_stride, double* V, int nfun, int m_lo, int m_hi,
                              int real_fmt, int rowsplit, const int* order_list) {
    int cols_per_fn = real_fmt ? 2 : 4;
    int NF = PC / cols_per_fn;
    size_t smem = sizeof(double) * 2 * PC * panel_stride(p->bw) + sizeof(uint32_t) * ((p->bw / 8 + 8 + 3) & ~3) +
                  sizeof(ColOut) * NC + (NC < 32 ? sizeof(double2) * LEG_WARPS * LEG_RING * 32 : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_legendre_inv<NC, PC>), smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((nfun + NF - 1) / NF, m_hi - m_lo, rowsplit);
    k_legendre_inv<NC, PC><<<grid, LEG_WARPS * 32, smem, p->stream>>>(table, p->d_order_start, shift, p->d_meta,
                                                                  p->d_rt_start, rco, ico, coef_stride, V, p->bw, nfun,
                                                                  m_lo, real_fmt, order_list, l2_prefetch_cap(PC),
                                                                  (NC < 32 && p->table_single && table == p->d_table) ? 1 : 0);
    return cudaGetLastError();
}

// Column-tile width: as wide as the batch and 200 KB of shared memory allow (table reuse per CTA), at
// least 8 (one MMA n-tile; a single field's 4 columns are padded with zeros).
static int pick_nc(int bw, int nfun, int real_fmt) {
    int cols = nfun * (real_fmt ? 2 : 4);
    size_t per_col = sizeof(double) * 2 * panel_stride(bw);
    int nc = 32;
    if (const char* e = getenv("S2KIT_CUDA_NC")) {  // tuning override: 8, 16 or 32 panel columns
        int v = atoi(e);
        if (v == 8 || v == 16 || v == 32) nc = v;
    }
    // prefer panels <= 110 KB (two CTAs per SM); a single 8-column panel may take up to the whole SM
    while (nc > 8 && (nc / 2 >= cols || per_col * nc > 110 * 1024)) nc /= 2;
    return nc;
}

// Row split: single-field / small-batch launches have too few CTAs per order; split each order's row
// tiles over several CTAs so that the grid covers the 148 SMs a few times over.
static int pick_rowsplit(int bw, int ncoltiles, int norders) {
    long ctas = (long)ncoltiles * norders;
    int rs = 1;
    int max_rs = (bw / 16 + LEG_WARPS - 1) / LEG_WARPS;  // at most one row tile per warp and parity
    if (max_rs < 1) max_rs = 1;
    while (ctas * rs < 148 * 4 && rs < max_rs) rs *= 2;
    return rs;
}

cudaError_t launch_legendre_fwd(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* X, double* rco,
                                double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format,
                                const int* order_list) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    int real_fmt = data_format == S2KIT_REAL;
    int nc = pick_nc(p->bw, nfun, real_fmt);
    if (nc == 8 && nfun * (real_fmt ? 2 : 4) <= 4) nc = 4;  // single field: half-width panel
    int NF = nc / (real_fmt ? 2 : 4);
    int rs = pick_rowsplit(p->bw, (nfun + NF - 1) / NF, m_hi - m_lo);
    int slot = prof_begin(p, S2KIT_K_LEGENDRE_FWD);
    cudaError_t e;
    switch (nc) {
        case 4: e = leg_fwd_nc<8, 4>(p, table, shift, X, rco, ico, coef_stride, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        case 8: e = leg_fwd_nc<8>(p, table, shift, X, rco, ico, coef_stride, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        case 16: e = leg_fwd_nc<16>(p, table, shift, X, rco, ico, coef_stride, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        default: e = leg_fwd_nc<32>(p, table, shift, X, rco, ico, coef_stride, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
    }
    prof_end(p, slot);
    return e;
}

cudaError_t launch_legendre_inv(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* rco,
                                const double* ico, long coef_stride, double* V, int nfun, int m_lo, int m_hi,
                                int data_format, const int* order_list) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    int real_fmt = data_format == S2KIT_REAL;
    int nc = pick_nc(p->bw, nfun, real_fmt);
    if (nc == 8 && nfun * (real_fmt ? 2 : 4) <= 4) nc = 4;  // single field: half-width panel
    if (nc >= 32 && p->table_single && table == p->d_table) nc = 16;  // one table copy: the A-order reader is the narrow path
    int NF = nc / (real_fmt ? 2 : 4);
    int rs = pick_rowsplit(p->bw, (nfun + NF - 1) / NF, m_hi - m_lo);
    int slot = prof_begin(p, S2KIT_K_LEGENDRE_INV);
    cudaError_t e;
    // batched bw = 256: four-column-tile units with TMA-staged table tiles (k_legendre_inv_q).  Default: 1.17-1.21 ms per
    // 1024 functions against 1.18-1.22 ms for k_legendre_inv<32,32> (three alternating runs on one box), with the LSU data pipe
    // at 47 % instead of 70 %.  S2KIT_CUDA_K4_QUAD=0 selects the pair kernel.
    static const int quad = [] {
        const char* ev = getenv("S2KIT_CUDA_K4_QUAD");
        return (ev && ev[0] == '0') ? 0 : 1;
    }();
    if (quad && nc == 32 && p->bw == 256 && !order_list && !(p->table_single && table == p->d_table)) {
        const size_t smem = sizeof(double) * 2 * 32 * panel_stride(p->bw) + sizeof(double2) * LEG_WARPS * LQ_STAGES * LQ * 32 +
                            8 * LEG_WARPS * LQ_STAGES;
        e = ensure_smem(reinterpret_cast<const void*>(k_legendre_inv_q), smem);
        if (e == cudaSuccess) {
            k_legendre_inv_q<<<dim3((nfun + NF - 1) / NF, m_hi - m_lo), LEG_WARPS * 32, smem, p->stream>>>(
                table, p->d_order_start, shift, rco, ico, coef_stride, V, p->bw, nfun, m_lo, real_fmt, l2_prefetch_cap(32));
            e = cudaGetLastError();
        }
        prof_end(p, slot);
        return e;
    }
    switch (nc) {
        case 4: e = leg_inv_nc<8, 4>(p, table, shift, rco, ico, coef_stride, V, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        case 8: e = leg_inv_nc<8>(p, table, shift, rco, ico, coef_stride, V, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        case 16: e = leg_inv_nc<16>(p, table, shift, rco, ico, coef_stride, V, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
        default: e = leg_inv_nc<32>(p, table, shift, rco, ico, coef_stride, V, nfun, m_lo, m_hi, real_fmt, rs, order_list); break;
    }
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
