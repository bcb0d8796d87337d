// s2k_legendre.cuh -- helpers shared by the Legendre contraction kernels (kernels_legendre.cu, kernels_fused.cu).
#pragma once
#include "s2k_internal.cuh"

namespace s2k {

constexpr int LEG_WARPS = 8;
// table tiles kept in flight per warp: few when the panel is wide (many DMMAs per tile hide the latency), many
// when a single field streams the table from HBM with 1-2 MMA column tiles per table tile
__host__ __device__ constexpr int leg_prefetch(int nc) { return nc >= 32 ? 4 : (nc >= 16 ? 8 : 12); }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// index of f^(m,l=|m|) in the coefficient arrays  (IndexOfHarmonicCoeff, util.c:42-49)
__device__ __forceinline__ int coef_base(int m, int bw) {
    if (m >= 0) return m * bw - (m * (m - 1)) / 2;
    int big = bw - 1;
    return (big * (big + 3)) / 2 + 1 + ((big + m) * (big + m + 1)) / 2;
}

// stride (doubles) of one column of the shared-memory panel: holds ceil(bw/2) entries of one parity,
// == 4 (mod 16) so that the 64-bit MMA fragment loads are bank-conflict free
__host__ __device__ inline int panel_stride(int bw) {
    int hb = ((bw + 1) / 2 + 7) / 8 * 8;
    return hb + ((4 - hb % 16) + 16) % 16;
}

// Pull one order's table (contiguous tiles) into L2 at CTA start.  The table is touched once per launch, in a short
// window per order, and the streaming traffic of the batch evicts it between launches: without this every tile of
// the main loop is a first-touch DRAM miss (~1 us) that the few tiles of register prefetch cannot cover
// (profiles/r1_ncu_summary.md).
// Bulk L2 prefetch of the whole order at CTA start (cp.async.bulk.prefetch.L2, 256 KiB per instruction, issued by
// one warp): no per-line LSU traffic, and for single large-bw fields it is what keeps HBM busy -- the memory system
// works through the queued prefetch while the main loop consumes tiles behind it (measured at bw = 2048: 2.5 TB/s
// with the whole-order prefetch vs 1.3-1.5 TB/s with register prefetch alone or a per-warp look-ahead).
__device__ __forceinline__ void prefetch_order_l2(const double* base, uint64_t tiles, int tid, int nthreads,
                                                  unsigned cap_bytes = 0xffffffffu) {
    (void)nthreads;
    const uint64_t bytes = tiles * 512;
    if (bytes > cap_bytes) return;
    const uint64_t piece = 256u << 10;
    if (tid < 32)
        for (uint64_t off = (uint64_t)tid * piece; off < bytes; off += 32 * piece) {
            unsigned len = (unsigned)((bytes - off) < piece ? (bytes - off) : piece);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(base) + off),
                         "r"(len)
                         : "memory");
        }
}

// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, all copies of a thread in flight at once
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Work items of one order are dealt to the W = (warps per CTA) x (row split) parallel slots in "snake" order:
// round i gives slot w the item i*W + w (i even) or i*W + W-1-w (i odd).  With the items sorted by decreasing
// cost this keeps the per-slot sums within a few percent of each other; plain round-robin left the first warp
// of a CTA with 2.4x the work of the last one at m = 0.
__device__ __forceinline__ int snake_item(int round, int slot, int nslots) {
    return round * nslots + ((round & 1) ? nslots - 1 - slot : slot);
}

// ---- the tile layout in closed form (what build_layout, plan.cu, tabulates in meta[] / rt_start[]) ----------------
// Kernels that want the addresses of tiles they will need LATER (prefetch) compute them instead of chasing the tables
// through dependent global loads.
__host__ __device__ inline BlockMeta block_meta_of(int m, int par, int bw) {
    BlockMeta mb;
    mb.rt_base = 0;  // not used with the closed form
    int rows = (bw - m - par + 1) / 2;
    mb.rows = rows < 0 ? 0 : rows;
    const int l = m + par;
    mb.len0 = (m & 1) ? (l - 1) / 2 + 1 : l / 2 + 1;  // RowSize(m, m + par), cospml.c:250-258
    mb.nrt = (mb.rows + 7) / 8;
    return mb;
}
// first tile of row tile rt (rt < nrt) relative to the block's first tile: rows grow by 8 entries = one tile per row tile
__host__ __device__ inline uint32_t row_tile_start_of(const BlockMeta& mb, int rt) {
    return (uint32_t)(rt * (rt - 1) / 2 + rt * ((mb.len0 + 14) >> 3));
}
// tiles of a whole parity block (= where the other parity's block starts inside the order)
__host__ __device__ inline uint32_t block_tiles_of(const BlockMeta& mb) {
    if (mb.nrt == 0) return 0;
    return row_tile_start_of(mb, mb.nrt - 1) + (uint32_t)((mb.len0 + mb.rows - 1 + 7) >> 3);
}

// number of 8-wide column tiles of row tile rt in a parity block
__device__ __forceinline__ int tiles_in_row(const BlockMeta& mb, int rt) {
    return (mb.len0 + min(8 * rt + 7, mb.rows - 1) + 7) >> 3;
}

// Forward main loop for one (parity, row tile): acc[j] += T_tile * X_panel for NC/8 column tiles.
// tp: first tile of the row tile (+ 2*lane); xp: panel base of this lane (parity, column g, slot q4).
template <int NC, int PFD = leg_prefetch(NC)>
__device__ __forceinline__ void fwd_row_tile(const double* __restrict__ tp, const double* xp, int CS, int ctn,
                                             double (&acc)[NC / 8][2], bool dead_lane = false) {
    // Ring of LEG_PREFETCH tiles in registers.  The refill is UNCONDITIONAL (index clamped to the last tile): a
    // predicated refill made ptxas load into a temporary and copy it into the ring slot right away, which waits
    // for the load and serialises the whole prefetch.
    constexpr int LEG_PREFETCH = PFD;
    double2 abuf[LEG_PREFETCH];
#pragma unroll
    for (int u = 0; u < LEG_PREFETCH; ++u)
        abuf[u] = __ldg(reinterpret_cast<const double2*>(tp + min(u, ctn - 1) * 64));
    for (int ct0 = 0; ct0 < ctn; ct0 += LEG_PREFETCH) {
#pragma unroll
        for (int u = 0; u < LEG_PREFETCH; ++u) {
            const int ct = ct0 + u;
            if (ct < ctn) {
                double b[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    b[j][0] = xp[j * 8 * CS + 8 * ct];
                    b[j][1] = xp[j * 8 * CS + 8 * ct + 4];
                    if (dead_lane) b[j][0] = b[j][1] = 0.0;  // MMA columns beyond a half-width panel
                }
                // k-step outer: consecutive DMMAs go to different accumulators
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], abuf[u].x, b[j][0]);
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], abuf[u].y, b[j][1]);
                // refill this slot only after its last use so the load can target the slot registers directly
                abuf[u] = __ldg(reinterpret_cast<const double2*>(tp + min(ct + LEG_PREFETCH, ctn - 1) * 64));
            }
        }
    }
}

// Inverse main loop for one (parity, column tile ct): acc[j] += C_panel * T_tile over the row tiles that reach ct.
// tbase: the table in B-fragment (tile-transposed) order + 2*lane, so one 128-bit load yields both k-steps' fragments.
// srt: this parity block's row-tile starts (shared memory); cp: panel base of this lane.
template <int NC>
__device__ __forceinline__ void inv_col_tile(const double* __restrict__ tbase, const uint32_t* srt, const BlockMeta& mb,
                                             int ct, const double* cp, int CS, double (&acc)[NC / 8][2],
                                             bool dead_lane = false) {
    int rt_min = 0;
    if (8 * ct >= mb.len0 + 7) rt_min = (8 * ct - mb.len0 - 7) / 8 + 1;
    // rows below rt_min never reach ct; from the first row tile that does, all later ones do (lengths grow)
    while (rt_min < mb.nrt && ct >= tiles_in_row(mb, rt_min)) ++rt_min;
    const int cnt = mb.nrt - rt_min;
    if (cnt <= 0) return;
    constexpr int LEG_PREFETCH = leg_prefetch(NC);
    double2 bbuf[LEG_PREFETCH];
#pragma unroll
    for (int u = 0; u < LEG_PREFETCH; ++u)
        bbuf[u] = __ldg(reinterpret_cast<const double2*>(tbase + ((uint64_t)srt[rt_min + min(u, cnt - 1)] + ct) * 64));
    for (int i0 = 0; i0 < cnt; i0 += LEG_PREFETCH) {
#pragma unroll
        for (int u = 0; u < LEG_PREFETCH; ++u) {
            const int i = i0 + u;
            if (i < cnt) {
                const int rt = rt_min + i;
                double a[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    a[j][0] = cp[j * 8 * CS + 8 * rt];
                    a[j][1] = cp[j * 8 * CS + 8 * rt + 4];
                    if (dead_lane) a[j][0] = a[j][1] = 0.0;
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], a[j][0], bbuf[u].x);
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], a[j][1], bbuf[u].y);
                // unconditional refill after the last use, clamped (see fwd_row_tile)
                bbuf[u] = __ldg(reinterpret_cast<const double2*>(
                    tbase + ((uint64_t)srt[rt_min + min(i + LEG_PREFETCH, cnt - 1)] + ct) * 64));
            }
        }
    }
}

// ---- register-blocked variants for wide (batched) panels ---------------------------------------------------------
// The panel fragments come from shared memory: one 64-bit load per lane and DMMA, which at NC = 32 keeps the LSU /
// shared-memory pipe busier than the FP64 tensor pipe (profiles/r1_ncu_summary.md: 71-88 % vs 53-55 %).  Handling TWO
// table tiles that need the same panel fragments per step -- two adjacent row tiles of one parity in the forward
// direction, two adjacent column tiles in the inverse -- halves that traffic.  Four tiles of register prefetch per
// stream (two left the first DMMA of every step waiting ~30 % of the step on the tile load, ncu source view); the
// kernels give up the third CTA per SM for the registers (128 instead of 80, which also stops ptxas from
// rematerialising the panel address from S2R / S2UR in every step).
constexpr int LEG_PF2_FWD = 4;  // k_legendre_fwd: 2 CTAs per SM, 128 registers
constexpr int LEG_PF2_INV = 2;  // k_legendre_inv: measured faster with 3 CTAs per SM (80 registers) and two tiles

// Forward: row tiles rt (tp0, ctn0 column tiles) and rt + 1 (tp1, ctn1 >= ctn0) of one parity block.
template <int NC, int PF>
__device__ __forceinline__ void fwd_row_tile2(const double* __restrict__ tp0, int ctn0, const double* __restrict__ tp1,
                                              int ctn1, const double* xp, int CS, double (&acc0)[NC / 8][2],
                                              double (&acc1)[NC / 8][2]) {
    double2 a0[PF], a1[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        a0[u] = __ldg(reinterpret_cast<const double2*>(tp0 + min(u, ctn0 - 1) * 64));
        a1[u] = __ldg(reinterpret_cast<const double2*>(tp1 + min(u, ctn1 - 1) * 64));
    }
    for (int ct0 = 0; ct0 < ctn1; ct0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int ct = ct0 + u;
            if (ct < ctn1) {
                double b[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    b[j][0] = xp[j * 8 * CS + 8 * ct];
                    b[j][1] = xp[j * 8 * CS + 8 * ct + 4];
                }
                if (ct < ctn0) {
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a0[u].x, b[j][0]);
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a1[u].x, b[j][0]);
                if (ct < ctn0) {
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a0[u].y, b[j][1]);
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a1[u].y, b[j][1]);
                a0[u] = __ldg(reinterpret_cast<const double2*>(tp0 + min(ct + PF, ctn0 - 1) * 64));
                a1[u] = __ldg(reinterpret_cast<const double2*>(tp1 + min(ct + PF, ctn1 - 1) * 64));
            }
        }
    }
}

// first row tile of a parity block whose row reaches column tile ct (== nrt when none does)
__device__ __forceinline__ int first_row_tile_reaching(const BlockMeta& mb, int ct) {
    int rt_min = 0;
    if (8 * ct >= mb.len0 + 7) rt_min = (8 * ct - mb.len0 - 7) / 8 + 1;
    while (rt_min < mb.nrt && ct >= tiles_in_row(mb, rt_min)) ++rt_min;
    return rt_min;
}

// Inverse: column tiles ct and ct + 1 of one parity block; the coefficient fragments of a row tile serve both.
template <int NC, int PF>
__device__ __forceinline__ void inv_col_tile2(const double* __restrict__ tbase, const uint32_t* srt, const BlockMeta& mb,
                                              int ct, const double* cp, int CS, double (&acc0)[NC / 8][2],
                                              double (&acc1)[NC / 8][2]) {
    const int rt_min = first_row_tile_reaching(mb, ct);
    const int rt_min1 = first_row_tile_reaching(mb, ct + 1);  // >= rt_min: rows only grow
    const int cnt = mb.nrt - rt_min;
    if (cnt <= 0) return;
    double2 b0[PF], b1[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        const int rt = rt_min + min(u, cnt - 1);
        const double* t = tbase + ((uint64_t)srt[rt] + ct) * 64;
        b0[u] = __ldg(reinterpret_cast<const double2*>(t));
        b1[u] = __ldg(reinterpret_cast<const double2*>(t + (rt >= rt_min1 ? 64 : 0)));
    }
    for (int i0 = 0; i0 < cnt; i0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 + u;
            if (i < cnt) {
                const int rt = rt_min + i;
                double a[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    a[j][0] = cp[j * 8 * CS + 8 * rt];
                    a[j][1] = cp[j * 8 * CS + 8 * rt + 4];
                }
                const bool two = rt >= rt_min1;
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a[j][0], b0[u].x);
                if (two) {
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a[j][0], b1[u].x);
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc0[j], a[j][1], b0[u].y);
                if (two) {
#pragma unroll
                    for (int j = 0; j < NC / 8; ++j) dmma(acc1[j], a[j][1], b1[u].y);
                }
                const int rn = rt_min + min(i + PF, cnt - 1);
                const double* t = tbase + ((uint64_t)srt[rn] + ct) * 64;
                b0[u] = __ldg(reinterpret_cast<const double2*>(t));
                b1[u] = __ldg(reinterpret_cast<const double2*>(t + (rn >= rt_min1 ? 64 : 0)));
            }
        }
    }
}

// ---- table tiles through a lane-private shared-memory ring (narrow panels: single fields stream the table from HBM) ---
// A register ring of N loads does not give N tiles of look-ahead: a warp has six scoreboards, the loads of a ring share
// them, and waiting for the oldest load also waits for younger ones (profiles/r1_ncu_pipe_summary.md).  Here every lane
// copies the 16 bytes it will feed to the DMMAs into its own slot of a ring in shared memory with cp.async;
// cp.async.wait_group counts completions exactly and in order, so LEG_RING tiles (x 512 bytes per warp) really are in
// flight -- what a single field needs to keep HBM busy, since every tile is used exactly once.  No other lane ever
// touches a slot: no barrier.  ring: this lane's slot 0 (slots are 32 double2 apart).
constexpr int LEG_RING = 8;

template <int NC>
__device__ __forceinline__ void fwd_row_tile_async(const double* __restrict__ tp, const double* xp, int CS, int ctn,
                                                   double (&acc)[NC / 8][2], bool dead_lane, double2* ring) {
#pragma unroll
    for (int u = 0; u < LEG_RING; ++u) {
        if (u < ctn) cp_async16(reinterpret_cast<double*>(ring + u * 32), tp + u * 64);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int ct = 0; ct < ctn; ++ct) {
        double2* slot = ring + (ct & (LEG_RING - 1)) * 32;
        asm volatile("cp.async.wait_group %0;" ::"n"(LEG_RING - 1) : "memory");
        const double2 av = *slot;
        double b[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            b[j][0] = xp[j * 8 * CS + 8 * ct];
            b[j][1] = xp[j * 8 * CS + 8 * ct + 4];
            if (dead_lane) b[j][0] = b[j][1] = 0.0;  // MMA columns beyond a half-width panel
        }
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.x, b[j][0]);
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.y, b[j][1]);
        // the DMMAs above have consumed the slot's registers: refill it with the tile LEG_RING ahead
        if (ct + LEG_RING < ctn) cp_async16(reinterpret_cast<double*>(slot), tp + (ct + LEG_RING) * 64);
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, possibly empty
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- table tiles staged by the TMA (bulk async copies) -----------------------------------------------------------
// The column tiles of one row tile are contiguous in memory, so instead of one 16-byte cp.async per lane and tile a single
// elected lane issues cp.async.bulk copies of BULK_CH tiles (1 KiB) into the warp's ring; completion is counted in
// bytes on one mbarrier per ring stage (mbarrier::complete_tx), the lanes then read their 16 bytes per tile from shared
// memory as before.  The copies are written by the async proxy: no LSU instruction or wavefront per tile on the way in.
// ring: the WARP's ring (BULK_NS * BULK_CH tiles of 32 double2); bar: shared-memory address of the warp's BULK_NS
// mbarriers (initialised to one arrival each); phases: the warp's phase bits, kept across calls.
constexpr int BULK_NS_MAX = 8;  // mbarriers reserved per warp

__device__ __forceinline__ void bulk_tile_copy(double2* dst, const double* src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     static_cast<unsigned>(__cvta_generic_to_shared(dst))),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned bar, unsigned parity) {
    unsigned done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (++spins > (1u << 26)) __trap();  // a lost completion must not hang the device
    }
}

// tile0: first tile of the row tile, no lane offset.  BULK_CH tiles per bulk copy, BULK_NS stages.
template <int NC, int BULK_CH, int BULK_NS>
__device__ __forceinline__ void fwd_row_tile_bulk(const double* __restrict__ tile0, const double* xp, int CS, int ctn,
                                                  double (&acc)[NC / 8][2], bool dead_lane, double2* ring, unsigned bar,
                                                  unsigned& phases, int lane) {
    static_assert(BULK_CH * BULK_NS == LEG_RING && BULK_NS <= BULK_NS_MAX, "the bulk stages fill the lane-private ring's memory");
    const int nch = (ctn + BULK_CH - 1) / BULK_CH;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < BULK_NS; ++s)
            if (s < nch)
                bulk_tile_copy(ring + s * BULK_CH * 32, tile0 + s * BULK_CH * 64, 512u * min(BULK_CH, ctn - s * BULK_CH), bar + 8 * s);
    }
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int s = c & (BULK_NS - 1);
        bulk_wait(bar + 8 * s, (phases >> s) & 1u);
        phases ^= 1u << s;
#pragma unroll
        for (int u = 0; u < BULK_CH; ++u) {
            const int ct = c * BULK_CH + u;
            if (ct < ctn) {
                const double2 av = ring[(s * BULK_CH + u) * 32 + lane];
                double b[NC / 8][2];
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) {
                    b[j][0] = xp[j * 8 * CS + 8 * ct];
                    b[j][1] = xp[j * 8 * CS + 8 * ct + 4];
                    if (dead_lane) b[j][0] = b[j][1] = 0.0;  // MMA columns beyond a half-width panel
                }
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.x, b[j][0]);
#pragma unroll
                for (int j = 0; j < NC / 8; ++j) dmma(acc[j], av.y, b[j][1]);
            }
        }
        __syncwarp();  // every lane has its values of this stage in registers: the stage may be overwritten
        if (lane == 0 && c + BULK_NS < nch) {
            const int cn = c + BULK_NS;
            bulk_tile_copy(ring + s * BULK_CH * 32, tile0 + cn * BULK_CH * 64, 512u * min(BULK_CH, ctn - cn * BULK_CH), bar + 8 * s);
        }
    }
}

// a_order: the tiles are stored in A-fragment order only (large-bandwidth Memo plans keep ONE table copy, plan.cu).  The
// copy into the ring is the same coalesced 16 bytes per lane as in the forward direction; the lane's two B-fragment
// values -- tile elements (q4, g) and (q4 + 4, g) -- are then READ from the other lanes' slots (two conflict-free 64-bit
// loads at the transposed positions), which makes the ring slot warp-shared: a __syncwarp after the lane's own copy has
// landed (then every lane's has) and one before the slot is refilled.  (Gathering with two 8-byte cp.async per lane
// instead was 34 % slower at bw = 2048: 2.44 vs 1.83 ms.)
// tbase: first tile of the order WITHOUT any lane offset.
template <int NC>
__device__ __forceinline__ void inv_col_tile_async(const double* __restrict__ tbase, const uint32_t* srt,
                                                   const BlockMeta& mb, int ct, const double* cp, int CS,
                                                   double (&acc)[NC / 8][2], bool dead_lane, double2* ring, int a_order,
                                                   int lane) {
    const int rt_min = first_row_tile_reaching(mb, ct);
    const int cnt = mb.nrt - rt_min;
    if (cnt <= 0) return;
    const int g = lane >> 2, q4 = lane & 3;
    const int o1 = tile_elem_offset(q4, g);  // element (q4, g) inside an A-order tile; (q4 + 4, g) is 32 doubles further
    auto copy_tile = [&](double2* slot, int rt) {
        cp_async16(reinterpret_cast<double*>(slot), tbase + ((uint64_t)srt[rt] + ct) * 64 + 2 * lane);
    };
#pragma unroll
    for (int u = 0; u < LEG_RING; ++u) {
        if (u < cnt) copy_tile(ring + u * 32, rt_min + u);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) {
        const int rt = rt_min + i;
        double2* slot = ring + (i & (LEG_RING - 1)) * 32;
        asm volatile("cp.async.wait_group %0;" ::"n"(LEG_RING - 1) : "memory");
        double2 bv;
        if (a_order) {
            __syncwarp();  // every lane's 16 bytes of this tile have landed
            const double* tile = reinterpret_cast<const double*>(slot - lane);
            bv = make_double2(tile[o1], tile[o1 + 32]);
        } else {
            bv = *slot;
        }
        double a[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            a[j][0] = cp[j * 8 * CS + 8 * rt];
            a[j][1] = cp[j * 8 * CS + 8 * rt + 4];
            if (dead_lane) a[j][0] = a[j][1] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc[j], a[j][0], bv.x);
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) dmma(acc[j], a[j][1], bv.y);
        if (a_order) __syncwarp();  // every lane has read the slot: it may be overwritten
        if (i + LEG_RING < cnt) copy_tile(slot, rt + LEG_RING);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace s2k
