// kernels_table.cu -- K7: device generation of the cosine-series tables of the associated Legendre functions.
//
// Replaces GenerateCosPmlTable                         src/legendre_polynomials/cospml.c:161-242
// with recurrence coefficients from L2_an / L2_cn     src/legendre_polynomials/util/l2_norms.c:16-38
// The bit-sensitive inputs -- Chebyshev nodes x_i = cos((2i+1) pi/2bw) (chebyshev_nodes.c:29-34) and the seeds
// P~_m^m(theta_i) [/ sin theta_i for odd m] (pmm.c:21-33, cospml.c:181-192) -- are computed on the HOST with the
// same libm and expression order as the reference and uploaded (trap 1).  The three-term
// recurrence runs here with the reference's operation order and explicitly un-fused multiplies/adds
// (cospml.c:218-221), so the sampled P~_l^m are bit-identical to the reference's; only the DCT differs in
// rounding (FFTW there, the radix kernels of s2k_fft.cuh here).
//
// Work unit = (order m, LCH consecutive degrees starting at l0): bw/8 threads hold the bw samples of two
// consecutive degrees in registers, roll the recurrence from l = m up to l0 (cheap), then per pair of
// degrees run ONE complex FFT of length bw (two real DCT-IIs via even/odd reordering + conjugate symmetry)
// and scatter the kept entries straight into the DMMA-tiled layout (s2k_internal.cuh).
#include <stdlib.h>

#include "s2k_fft.cuh"
#include "s2k_internal.cuh"
#include "s2k_legendre.cuh"

namespace s2k {

__device__ __forceinline__ void rec_step(const double (&x)[8], double (&prev)[8], double (&cur)[8], double2 ac) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        double t1 = __dmul_rn(ac.y, prev[e]);
        double t2 = __dmul_rn(cur[e], x[e]);
        double t3 = __dmul_rn(ac.x, t2);
        prev[e] = cur[e];
        cur[e] = __dadd_rn(t3, t1);
    }
}

// transposed = 0: A-fragment order (forward contraction); 1: the same tile with rows and columns swapped inside the
// tile = B-fragment order, so the inverse contraction also reads one coalesced 128-bit value pair per lane
__device__ __forceinline__ void tile_store(double* __restrict__ table, uint64_t order_tile0, const BlockMeta& mb,
                                           const uint32_t* __restrict__ rt_start, int r, int c, double v,
                                           int transposed) {
    uint64_t tile = order_tile0 + rt_start[mb.rt_base + (r >> 3)] + (uint64_t)(c >> 3);
    table[tile * 64 + (transposed ? tile_elem_offset(c & 7, r & 7) : tile_elem_offset(r & 7, c & 7))] = v;
}

template <int NB, int G>
__global__ void __launch_bounds__(NB / 8 * G) k_table_gen(double* __restrict__ table,
                                                          const uint64_t* __restrict__ order_start, uint64_t shift,
                                                          const BlockMeta* __restrict__ meta,
                                                          const uint32_t* __restrict__ rt_start,
                                                          const int* __restrict__ units, int unit_lo, int unit_hi,
                                                          int lch, int transposed, const double* __restrict__ nodes,
                                                          const double* __restrict__ seeds,
                                                          const double2* __restrict__ rec,
                                                          const double2* __restrict__ tw,
                                                          const double2* __restrict__ qtab) {
    constexpr int T8 = NB / 8, NP = fft_padded_len(NB);
    extern __shared__ double2 smem2[];
    const int tid = threadIdx.x, g = tid / T8, t = tid % T8;
    double2* sx = smem2 + g * NP;
    int u = unit_lo + blockIdx.x * G + g;
    const bool live = u < unit_hi;
    if (!live) u = unit_lo;
    const int m = units[2 * u], l0 = units[2 * u + 1];
    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    const uint64_t tile0 = order_start[m] - shift;

    double x[8], prev[8], cur[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int p = t + e * T8;
        int i = (p < NB / 2) ? 2 * p : 2 * (NB - 1 - p) + 1;  // even/odd reordering of the DCT input
        x[e] = __ldg(nodes + i);
        cur[e] = __ldg(seeds + (long)m * NB + i);
        prev[e] = 0.0;
    }
    const double2* rc = rec + (long)m * NB;
    {
        // roll the recurrence up to the unit's first degree; the next step's (a, c) pair is loaded one step ahead (the
        // load used to sit in front of every step: 19 % of the kernel's samples, profiles/r1_ncu_summary.md section 6)
        double2 ac = __ldg(rc + m);
        for (int l = m; l < l0; ++l) {
            const double2 nx = __ldg(rc + min(l + 1, NB - 1));
            rec_step(x, prev, cur, ac);
            ac = nx;
        }
    }

    const double fudge = 1.0 / sqrt((double)NB);  // cospml.c:206
    for (int pair = 0; pair < lch / 2; ++pair) {
        const int l = l0 + 2 * pair;
        double xr[8], xi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) xr[e] = cur[e];
        if (l + 1 < NB) rec_step(x, prev, cur, __ldg(rc + l));
#pragma unroll
        for (int e = 0; e < 8; ++e) xi[e] = (l + 1 < NB) ? cur[e] : 0.0;
        if (l + 2 < NB) rec_step(x, prev, cur, __ldg(rc + l + 1));

        fft_block<NB>(xr, xi, sx, t, g, tw);
        fft_sync<NB>(g);
#pragma unroll
        for (int e = 0; e < 8; ++e) sx[fft_pad(fft_out_index<NB>(e, t))] = make_double2(xr[e], xi[e]);
        fft_sync<NB>(g);
        if (live && l < NB) {
            // degree l keeps cosine indices of parity 0 (l0 - m is even), degree l+1 those of parity 1
            const int ra = (l - m) >> 1;  // row inside either parity block
            for (int k = t; k < NB; k += T8) {
                const int pk = k & 1, c = k >> 1;
                if (pk && l + 1 >= NB) continue;
                const BlockMeta& mb = pk ? mb1 : mb0;
                if (c >= mb.len0 + ra) continue;
                int nk = (NB - k) & (NB - 1);
                const double2 za = sx[fft_pad(k)], zb = sx[fft_pad(nk)];
                const double ar = za.x, ai = za.y, br = zb.x, bi = zb.y;
                double2 q = __ldg(qtab + k);
                double y = pk ? (q.x * (ai + bi) - q.y * (ar - br)) : (q.x * (ar + br) + q.y * (ai - bi));
                if (k == 0) y *= 0.70710678118654752440;  // cospml.c:205
                tile_store(table, tile0, mb, rt_start, ra, c, y * fudge, transposed);
            }
        }
    }
}


// ---- K7 on the HALF grid (bw >= 1024) ------------------------------------------------------------------------------
// P~_l^m(cos theta) is symmetric (l - m even) or antisymmetric (l - m odd) about the equator, which is why a table row only
// keeps every other cosine index.  The reference (cospml.c:203-224) -- and k_table_gen above -- still take a full
// length-bw DCT-II of every row and throw half of the outputs away.  With a[i] the samples on the first bw/2 nodes:
//   symmetric row      X[2j]     = 2 * sum_{i < bw/2} a[i] cos(pi j (2i+1) / bw)             = 2 DCT-II_{bw/2}(a)[j]
//   antisymmetric row  X[2j + 1] = 2 * sum_{i < bw/2} a[i] cos(pi (2j+1)(2i+1) / (2 bw))     = 2 DCT-IV_{bw/2}(a)[j]
// so the recurrence runs on HALF the nodes and a step of FOUR degrees costs one complex FFT of length bw/2 (the two
// symmetric rows as real and imaginary part, separated by conjugate symmetry as before) plus two complex FFTs of length
// bw/4 (one DCT-IV each: u[n] = a[2n] + i a[M-1-2n], pre-twiddle e^{-i pi (4n+1)/(4M)}, FFT, post-twiddle e^{-i pi k/M},
// D[2k] = Re, D[M-1-2k] = -Im; M = bw/2) instead of two complex FFTs of length bw: 2.4x fewer FFT flops, half the
// recurrence, a third of the shared-memory traffic (k_table_gen<1024> ran the LSU data pipe at 75 %, the FP64 pipe at
// 53 %: profiles/r2_fly_bw1024.json).  The samples are the reference's bit for bit (same seeds, same recurrence); what
// changes is the rounding inside the cosine transform, as it already did against FFTW.
// Group = bw/16 threads (8 half-grid nodes each, in the even/odd-reordered order the length-bw/2 DCT-II FFT loads); the
// lower / upper half of the group runs the DCT-IV transform of the first / second antisymmetric row.
// first tile (relative to the order's first tile) of the row tile holding row r of parity block `par`, in closed form
// (s2k_legendre.cuh: row_tile_start_of / block_tiles_of -- the same numbers build_layout tabulates in rt_start[])
__device__ __forceinline__ uint32_t half_row_tile0(const BlockMeta& mb0, const BlockMeta& mb, int par, int r) {
    return (par ? block_tiles_of(mb0) : 0u) + row_tile_start_of(mb, r >> 3);
}
// entries of row r including the zero padding up to the end of its last tile
__device__ __forceinline__ int half_padded_len(const BlockMeta& mb, int r) { return 8 * tiles_in_row(mb, r >> 3); }

// Recurrence checkpoints for the half-grid generator: (P~_{l0-1}^m, P~_{l0}^m) on the half grid at the first degree l0 of
// every work unit, in the generator's own thread order (position p = t + e T).  A unit used to roll the recurrence up
// from l = m to its l0 on its own -- at m = 0 the eight units of an order repeated 3.5x the order's whole recurrence.  The
// values are what the roll-up produced, bit for bit (same operations in the same order); one CTA per order writes them
// once per plan (bw = 1024: 67 MB, 1/22 of the table the Fly variant avoids storing).
template <int NB>
__global__ void __launch_bounds__(NB / 8) k_rec_checkpoints(double* __restrict__ ckpt, const int* __restrict__ unit_first,
                                                           int lch, const double* __restrict__ nodes,
                                                           const double* __restrict__ seeds,
                                                           const double2* __restrict__ rec) {
    constexpr int M = NB / 2, T = NB / 8;
    const int t = threadIdx.x, m = blockIdx.x;
    const int u0 = unit_first[m], nu = unit_first[m + 1] - u0;  // units of order m, LAST degrees first
    double x[4], prev[4], cur[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int p = t + e * T;
        const int i = (p < M / 2) ? 2 * p : 2 * (M - 1 - p) + 1;
        x[e] = __ldg(nodes + i);
        cur[e] = __ldg(seeds + (long)m * NB + i);
        prev[e] = 0.0;
    }
    const double2* rc = rec + (long)m * NB;
    for (int j = 0; j < nu; ++j) {
        double* dst = ckpt + (size_t)(u0 + nu - 1 - j) * 2 * M;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            dst[t + e * T] = prev[e];
            dst[M + t + e * T] = cur[e];
        }
        if (j + 1 == nu) break;
        const int l0 = m + j * lch;
        double2 ac[4];
        for (int l = l0; l < l0 + lch; l += 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ac[i] = __ldg(rc + l + i);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double t1 = __dmul_rn(ac[i].y, prev[e]);
                    const double t2 = __dmul_rn(cur[e], x[e]);
                    const double t3 = __dmul_rn(ac[i].x, t2);
                    prev[e] = cur[e];
                    cur[e] = __dadd_rn(t3, t1);
                }
            }
        }
    }
}

// One CTA = one work unit = NB/8 threads.  Every thread carries FOUR half-grid nodes through the recurrence (12 doubles of
// state); the rows go to shared memory in the order their transforms load them, then the CTA splits: the first half of
// the threads runs the length-M transform of the two symmetric rows, the third and fourth quarter one DCT-IV each -- the
// three transforms of a step run side by side instead of one after the other, and nobody holds recurrence state for
// eight nodes next to sixteen transform registers (the first version of this kernel: 128 registers with spills, 16
// warps per SM, the transforms in sequence; 1.53 ms per direction at bw = 1024 against 1.2 here).
template <int NB>
__global__ void __launch_bounds__(NB / 8, 640 / (NB / 8)) k_table_gen_half(
    double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t shift, const BlockMeta* __restrict__ meta,
    const int* __restrict__ units, int unit_lo, int unit_hi, int lch, int transposed, const double* __restrict__ nodes,
    const double* __restrict__ ckpt, const double2* __restrict__ rec, const double2* __restrict__ tw,
    const double2* __restrict__ qtab) {
    constexpr int M = NB / 2, K = NB / 4, T = NB / 8, TA = T / 2, TB = T / 4;
    constexpr int LA = fft_padded_len(M), LB = fft_padded_len(K);
    static_assert(TB >= 32, "every transform is owned by whole warps");
    extern __shared__ double2 smem2[];
    double2* sa = smem2;       // inputs / exchange row of the length-M transform
    double2* sb = smem2 + LA;  // DCT-IV inputs / exchange rows: first antisymmetric row, then the second
    const int t = threadIdx.x;
    const int u = unit_lo + blockIdx.x;
    if (u >= unit_hi) return;
    const int m = units[2 * u], l0 = units[2 * u + 1];
    const BlockMeta mb0 = block_meta_of(m, 0, NB), mb1 = block_meta_of(m, 1, NB);
    double* const otab = table + (order_start[m] - shift) * 64;
    // element (r, c) of a parity block whose row tile starts at tile rt0: every tile element of the launch's orders is
    // written exactly once (values, or zeros in the padding), so the scratch table needs no memset
    auto put = [&](uint32_t rt0, int r, int c, double v) {
        otab[((uint64_t)rt0 + (uint32_t)(c >> 3)) * 64 + (transposed ? tile_elem_offset(c & 7, r & 7) : tile_elem_offset(r & 7, c & 7))] = v;
    };

    // offset of element (r, c) inside its tile
    auto eoff = [&](int r, int c7) { return transposed ? tile_elem_offset(c7, r & 7) : tile_elem_offset(r & 7, c7); };

    // positions p = t + e T of the even/odd-reordered length-M DCT-II input; registers 0, 1 hold nodes 2n (n = t, t + T),
    // registers 2, 3 their DCT-IV partners M-1-2n
    double x[4], prev[4], cur[4];
    {
        const double* ck = ckpt + (size_t)u * 2 * M;  // the recurrence state at the unit's first degree (k_rec_checkpoints)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = t + e * T;
            const int i = (p < M / 2) ? 2 * p : 2 * (M - 1 - p) + 1;
            x[e] = __ldg(nodes + i);
            prev[e] = __ldg(ck + p);
            cur[e] = __ldg(ck + M + p);
        }
    }
    auto step = [&](double2 ac) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double t1 = __dmul_rn(ac.y, prev[e]);
            const double t2 = __dmul_rn(cur[e], x[e]);
            const double t3 = __dmul_rn(ac.x, t2);
            prev[e] = cur[e];
            cur[e] = __dadd_rn(t3, t1);
        }
    };
    const double2* rc = rec + (long)m * NB;
    // DCT-IV pre-twiddles (cos, sin)(pi (4n+1) / (2 bw)), n = t, t + T
    const double2 pre0 = __ldg(qtab + 4 * t + 1), pre1 = __ldg(qtab + 4 * (t + T) + 1);
    auto stage_antisym = [&](double2* dst, bool have) {
        const double a0 = have ? cur[0] : 0.0, b0 = have ? cur[2] : 0.0, a1 = have ? cur[1] : 0.0, b1 = have ? cur[3] : 0.0;
        dst[fft_pad(t)] = make_double2(a0 * pre0.x + b0 * pre0.y, b0 * pre0.x - a0 * pre0.y);
        dst[fft_pad(t + T)] = make_double2(a1 * pre1.x + b1 * pre1.y, b1 * pre1.x - a1 * pre1.y);
    };
    // roles after the rows are staged
    const bool role_a = t < TA;
    const int hb = (t - TA) / TB, tb = (t - TA) % TB;  // DCT-IV transform and thread inside it (threads >= TA)
    // post-twiddle bases: the entries for j = t + i TA (separation) and k = tb + s K/8 (DCT-IV) are these rotated by multiples of pi/16
    const double2 q0 = __ldg(qtab + (role_a ? 2 * t : 4 * tb));
    const double fudge = 1.0 / sqrt((double)NB);  // cospml.c:206

    double2 ac[4];  // recurrence coefficients of the coming four steps, loaded one quad ahead
#pragma unroll
    for (int i = 0; i < 4; ++i) ac[i] = __ldg(rc + min(l0 + i, NB - 1));
    for (int quad = 0; quad < lch / 4; ++quad) {
        const int l = l0 + 4 * quad;
        if (l >= NB) break;
        const int ra = (l - m) >> 1;  // row of degree l (parity 0) and of degree l + 1 (parity 1) in their blocks
        double r0[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) r0[e] = cur[e];
        if (l + 1 < NB) step(ac[0]);
        stage_antisym(sb, l + 1 < NB);
        if (l + 2 < NB) step(ac[1]);
#pragma unroll
        for (int e = 0; e < 4; ++e) sa[fft_pad(t + e * T)] = make_double2(r0[e], (l + 2 < NB) ? cur[e] : 0.0);
        if (l + 3 < NB) step(ac[2]);
        stage_antisym(sb + LB, l + 3 < NB);
        if (l + 4 < NB) step(ac[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) ac[i] = __ldg(rc + min(l + 4 + i, NB - 1));
        __syncthreads();

        if (role_a) {
            // ---- the two symmetric rows: one complex FFT of length M, separated by conjugate symmetry
            double xr[8], xi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double2 v = sa[fft_pad(t + e * TA)];
                xr[e] = v.x;
                xi[e] = v.y;
            }
            fft_block<M, 2>(xr, xi, sa, t, 0, tw);
            // Separation needs Z[k] and Z[M-k].  A thread keeps the lower half of its outputs (k < M/2) in registers and
            // fetches their partners from the upper half, which alone goes through shared memory (at k - M/2); it then
            // finishes BOTH columns k and M - k of the two rows: (cos, sin)(pi (M-k)/2M) = (sin, cos)(pi k/2M).
            constexpr int RA = fft_last_radix(M);
            fft_sync<M>(0);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (fft_slot<RA>(e) >= 4) sa[fft_pad(fft_out_index<M>(e, t) - M / 2)] = make_double2(xr[e], xi[e]);
            fft_sync<M>(0);
            const bool two = l + 2 < NB;
            const int len_a = mb0.len0 + ra, len_b = two ? len_a + 1 : 0;
            const int pad_a = half_padded_len(mb0, ra), pad_b = two ? half_padded_len(mb0, ra + 1) : 0;
            const uint32_t rta = half_row_tile0(mb0, mb0, 0, ra), rtb = half_row_tile0(mb0, mb0, 0, ra + 1);
            // column k = t + s TA: the tile advances by TA/8 per s, the position inside the tile (t & 7) never changes;
            // column M - k: tile (M - t)/8 - s TA/8, position (M - t) & 7
            const int cm = M - t;
            double* const pa = otab + ((uint64_t)rta + (uint32_t)(t >> 3)) * 64 + eoff(ra, t & 7);
            double* const pb = otab + ((uint64_t)rtb + (uint32_t)(t >> 3)) * 64 + eoff(ra + 1, t & 7);
            double* const pam = otab + ((uint64_t)rta + (uint32_t)(cm >> 3)) * 64 + eoff(ra, cm & 7);
            double* const pbm = otab + ((uint64_t)rtb + (uint32_t)(cm >> 3)) * 64 + eoff(ra + 1, cm & 7);
            const double scale = 2.0 * fudge;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int sl = fft_slot<RA>(e);
                if (sl >= 4) continue;
                const int k = t + sl * TA, km = M - k;  // k < M/2 < km (km = M for k = 0: no such column)
                if (k >= pad_a && k >= pad_b && km >= pad_a && km >= pad_b) continue;
                const double ar = xr[e], ai = xi[e];
                double br = ar, bi = ai;  // k = 0: Z[M] = Z[0]
                if (k != 0) {
                    const double2 zb = sa[fft_pad(M / 2 - k)];
                    br = zb.x;
                    bi = zb.y;
                }
                const double2 q = quarter_rot(q0, sl);  // (cos, sin)(pi k / 2M)
                const double sr = ar + br, dr = ar - br, si = ai + bi, di = ai - bi;
                const double sc0 = (k == 0) ? scale * 0.70710678118654752440 : scale;  // cospml.c:205
                if (k < pad_a) pa[sl * (TA / 8) * 64] = k < len_a ? (q.x * sr + q.y * di) * sc0 : 0.0;
                if (k < pad_b) pb[sl * (TA / 8) * 64] = k < len_b ? (q.x * si - q.y * dr) * sc0 : 0.0;
                if (k != 0) {
                    if (km < pad_a) pam[-sl * (TA / 8) * 64] = km < len_a ? (q.y * sr - q.x * di) * scale : 0.0;
                    if (km < pad_b) pbm[-sl * (TA / 8) * 64] = km < len_b ? (q.y * si + q.x * dr) * scale : 0.0;
                }
            }
            // column M/2 pairs with itself; it lives in the upper half (slot 4 of thread 0)
            if (t == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (fft_slot<RA>(e) == 4) {
                        const int k = M / 2;
                        const double h = 0.70710678118654752440 * 2.0 * scale;  // cos = sin = sqrt(1/2), ar + br = 2 ar
                        const uint32_t ta = rta + (uint32_t)(k >> 3), tb2 = rtb + (uint32_t)(k >> 3);
                        if (k < pad_a) otab[(uint64_t)ta * 64 + eoff(ra, 0)] = k < len_a ? xr[e] * h : 0.0;
                        if (k < pad_b) otab[(uint64_t)tb2 * 64 + eoff(ra + 1, 0)] = k < len_b ? xi[e] * h : 0.0;
                    }
            }
        } else {
            // ---- one antisymmetric row per quarter of the CTA: DCT-IV through a complex FFT of length K
            double2* sx = sb + hb * LB;
            double yr[8], yi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double2 v = sx[fft_pad(tb + e * TB)];
                yr[e] = v.x;
                yi[e] = v.y;
            }
            fft_block<K, 4>(yr, yi, sx, tb, 1 + hb, tw);
            const int lrow = l + 1 + 2 * hb, rrow = ra + hb;
            if (lrow < NB) {
                const int len = mb1.len0 + rrow, pad = half_padded_len(mb1, rrow);
                const uint32_t rt0 = half_row_tile0(mb0, mb1, 1, rrow);
                const double scale = 4.0 * fudge;
                constexpr int R = fft_last_radix(K);
                // columns 2k and M-1-2k with k = tb + s TB: the tiles move by +- s TB/4, the positions inside the tile stay
                const int c0b = 2 * tb, c1b = M - 1 - 2 * tb;
                double* const p0 = otab + ((uint64_t)rt0 + (uint32_t)(c0b >> 3)) * 64 + eoff(rrow, c0b & 7);
                double* const p1 = otab + ((uint64_t)rt0 + (uint32_t)(c1b >> 3)) * 64 + eoff(rrow, c1b & 7);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int sl = fft_slot<R>(e);
                    const double2 q = quarter_rot(q0, sl);  // (cos, sin)(pi k / M)
                    const int c0 = c0b + 2 * sl * TB, c1 = c1b - 2 * sl * TB;
                    if (c0 < pad) p0[sl * (TB / 4) * 64] = c0 < len ? (yr[e] * q.x + yi[e] * q.y) * scale : 0.0;
                    if (c1 < pad) p1[-sl * (TB / 4) * 64] = c1 < len ? (yr[e] * q.y - yi[e] * q.x) * scale : 0.0;
                }
            }
        }
        __syncthreads();  // sa / sb are rewritten by the next step
    }
    // the unit that ends the order clears the padding rows of both blocks' last row tiles
    if (l0 + lch >= NB) {
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const BlockMeta& mb = par ? mb1 : mb0;
            if (mb.nrt == 0) continue;
            const int pad = half_padded_len(mb, mb.rows - 1);
            const uint32_t rt0 = half_row_tile0(mb0, mb, par, mb.rows - 1);
            for (int r = mb.rows; r < 8 * mb.nrt; ++r)
                for (int c = t; c < pad; c += T) put(rt0, r, c, 0.0);
        }
    }
}

// Any bandwidth: one CTA per order, thread i owns the nodes i, i + blockDim, ... (two per thread above bw = 1024), DCT by
// the O(bw^2) definition.  Used for bandwidths that are not powers of two (the reference accepts any bw).
constexpr int DIRECT_NODES = 2;
__global__ void k_table_gen_direct(double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t shift,
                                   const BlockMeta* __restrict__ meta, const uint32_t* __restrict__ rt_start, int m_lo,
                                   int bw, int transposed, const double* __restrict__ nodes, const double* __restrict__ seeds,
                                   const double2* __restrict__ rec, const double2* __restrict__ qtab) {
    extern __shared__ double sm[];
    const int m = m_lo + blockIdx.x;
    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    const uint64_t tile0 = order_start[m] - shift;
    double x[DIRECT_NODES], prev[DIRECT_NODES], cur[DIRECT_NODES];
#pragma unroll
    for (int u = 0; u < DIRECT_NODES; ++u) {
        const int i = threadIdx.x + u * blockDim.x;
        x[u] = prev[u] = cur[u] = 0.0;
        if (i < bw) {
            x[u] = nodes[i];
            cur[u] = seeds[(long)m * bw + i];
        }
    }
    const double fudge = 1.0 / sqrt((double)bw);
    for (int l = m; l < bw; ++l) {
#pragma unroll
        for (int u = 0; u < DIRECT_NODES; ++u) {
            const int i = threadIdx.x + u * blockDim.x;
            if (i < bw) sm[i] = cur[u];
        }
        __syncthreads();
        const int p = (l - m) & 1, r = (l - m) >> 1;
        const BlockMeta& mb = p ? mb1 : mb0;
        for (int k = threadIdx.x; k < bw; k += blockDim.x) {
            if ((k & 1) != p || (k >> 1) >= mb.len0 + r) continue;
            double acc = 0.0;
            for (int s = 0; s < bw; ++s) acc += sm[s] * qtab[(int)(((long)(2 * s + 1) * k) % (4 * bw))].x;
            acc *= 2.0;
            if (k == 0) acc *= 0.70710678118654752440;
            tile_store(table, tile0, mb, rt_start, r, k >> 1, acc * fudge, transposed);
        }
        __syncthreads();
        if (l + 1 < bw) {
            const double2 ac = rec[(long)m * bw + l];
#pragma unroll
            for (int u = 0; u < DIRECT_NODES; ++u) {
                double t1 = __dmul_rn(ac.y, prev[u]);
                double t2 = __dmul_rn(cur[u], x[u]);
                double t3 = __dmul_rn(ac.x, t2);
                prev[u] = cur[u];
                cur[u] = __dadd_rn(t3, t1);
            }
        }
    }
}

// (a_l^m, c_l^m), same expression order as l2_norms.c:16-38, every operation individually rounded
__global__ void k_rec_coeffs(double2* __restrict__ rec, int bw) {
    int l = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (l >= bw) return;
    double2 out = make_double2(0.0, 0.0);
    if (l >= m) {
        double dl = (double)l;
        double two_l = __dmul_rn(2.0, dl);
        double lm1 = (double)(l - m) + 1.0;       // l - m + 1.
        double lpm1 = (double)(l + m) + 1.0;      // l + m + 1.
        double r1 = __ddiv_rn(two_l + 3.0, two_l + 1.0);
        double r2 = __ddiv_rn(lm1, lpm1);
        double a = __dmul_rn(__dsqrt_rn(__dmul_rn(r1, r2)), __ddiv_rn(two_l + 1.0, lm1));
        double c = 0.0;
        if (l != 0) {
            double s1 = __ddiv_rn(two_l + 3.0, two_l - 1.0);
            double s3 = __ddiv_rn((double)l - (double)m, (double)l + (double)m);
            double prod = __dmul_rn(__dmul_rn(s1, r2), s3);
            c = __dmul_rn(__dmul_rn(-1.0, __dsqrt_rn(prod)), __ddiv_rn((double)(l + m), lm1));
        }
        out = make_double2(a, c);
    }
    rec[(long)m * bw + l] = out;
}

// tile layout -> the reference's packed layout (rows l = m..bw-1, RowSize(m,l) entries each)
__device__ __forceinline__ int packed_h(int l) {  // sum_{d<l} (d/2 + 1)
    int h = (l / 2) * (l / 2 + 1);
    return (l & 1) ? h + l / 2 + 1 : h;
}

__global__ void k_table_unpack(const double* __restrict__ table, const uint64_t* __restrict__ order_start,
                               uint64_t shift, const BlockMeta* __restrict__ meta,
                               const uint32_t* __restrict__ rt_start, int m, int bw, double* __restrict__ out) {
    const int l = m + blockIdx.x;
    const int p = (l - m) & 1, r = (l - m) >> 1;
    const BlockMeta mb = meta[2 * m + p];
    const int len = mb.len0 + r;
    // TableOffset(m,l), cospml.c:123-134
    const int row0 = (m & 1) ? packed_h(l - 1) - packed_h(m - 1) : packed_h(l) - packed_h(m);
    const uint64_t tile0 = order_start[m] - shift;
    for (int c = threadIdx.x; c < len; c += blockDim.x) {
        uint64_t tile = tile0 + rt_start[mb.rt_base + (r >> 3)] + (uint64_t)(c >> 3);
        out[row0 + c] = table[tile * 64 + tile_elem_offset(r & 7, c & 7)];
    }
}

// ------------------------------------------------------------------------------------------------ launchers
// half-grid generator at bw >= 1024 (S2KIT_CUDA_TABLE_FULL=1 keeps the full-length transforms for comparison)
static bool table_gen_half_used(const s2kit_cuda_plan* p) {
    static const int full = [] {
        const char* e = getenv("S2KIT_CUDA_TABLE_FULL");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return !full && p->fast && p->bw >= 1024 && table_unit_rows(p->bw) % 4 == 0;
}

template <int NB>
static cudaError_t table_gen_nb(s2kit_cuda_plan* p, double* table, uint64_t shift, int unit_lo, int unit_hi, int lch,
                                int transposed) {
    constexpr int T8 = NB / 8;
    constexpr int G = (256 / T8) < 1 ? 1 : ((256 / T8) > 8 ? 8 : (256 / T8));
    size_t smem = sizeof(double2) * G * fft_padded_len(NB);
    if (smem > 48 * 1024) {
        cudaError_t e =
            ensure_smem(reinterpret_cast<const void*>(k_table_gen<NB, G>), smem);
        if (e != cudaSuccess) return e;
    }
    int nunits = unit_hi - unit_lo;
    if constexpr (NB >= 1024) {
        if (table_gen_half_used(p)) {
            const size_t smem_h = sizeof(double2) * (fft_padded_len(NB / 2) + 2 * fft_padded_len(NB / 4));
            cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_table_gen_half<NB>), smem_h);
            if (e != cudaSuccess) return e;
            if (!p->d_ckpt) {  // first use: the recurrence state at every unit's first degree
                const size_t nu = p->h_units.size() / 2;
                e = cudaMalloc((void**)&p->d_ckpt, sizeof(double) * nu * NB);
                if (e != cudaSuccess) return e;
                p->own_ckpt = true;
                if (!p->d_unit_first) {
                    e = cudaMalloc((void**)&p->d_unit_first, sizeof(int) * p->h_unit_first.size());
                    if (e == cudaSuccess)
                        e = cudaMemcpyAsync(p->d_unit_first, p->h_unit_first.data(), sizeof(int) * p->h_unit_first.size(),
                                            cudaMemcpyHostToDevice, p->stream);
                    if (e != cudaSuccess) return e;
                }
                k_rec_checkpoints<NB><<<NB, NB / 8, 0, p->stream>>>(p->d_ckpt, p->d_unit_first, lch, p->d_nodes, p->d_seeds, p->d_rec);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
            }
            k_table_gen_half<NB><<<nunits, NB / 8, smem_h, p->stream>>>(table, p->d_order_start, shift, p->d_meta, p->d_units,
                                                                        unit_lo, unit_hi, lch, transposed, p->d_nodes,
                                                                        p->d_ckpt, p->d_rec, p->d_tw_b, p->d_q_b);
            return cudaGetLastError();
        }
    }
    k_table_gen<NB, G><<<(nunits + G - 1) / G, T8 * G, smem, p->stream>>>(
        table, p->d_order_start, shift, p->d_meta, p->d_rt_start, p->d_units, unit_lo, unit_hi, lch, transposed, p->d_nodes,
        p->d_seeds, p->d_rec, p->d_tw_b, p->d_q_b);
    return cudaGetLastError();
}

// degrees per generator work unit: every unit rolls the recurrence up from l = m, so longer units repeat less of it, but
// a unit is one CTA's serial chain (S2KIT_CUDA_TABLE_LCH overrides for tuning; must be even)
int table_unit_rows(int bw) {
    static int forced = [] {
        const char* e = getenv("S2KIT_CUDA_TABLE_LCH");
        int v = e ? atoi(e) : 0;
        return (v >= 2 && v % 2 == 0) ? v : 0;
    }();
    if (forced) return forced;
    return bw <= 512 ? 32 : 128;  // bw = 1024 Fly forward: 32 -> 3.57 ms, 64 -> 3.06, 128 -> 2.91, 256 -> 3.09
}

cudaError_t launch_table_gen(s2kit_cuda_plan* p, double* table, uint64_t shift, int m_lo, int m_hi, int transposed) {
    if (m_hi <= m_lo) return cudaSuccess;
    uint64_t t0 = p->h_order_start[m_lo], t1 = p->h_order_start[m_hi];
    cudaError_t e = cudaSuccess;
    // the half-grid generator writes every tile element (padding included); the others only the entries they compute
    if (!table_gen_half_used(p)) e = cudaMemsetAsync(table + (t0 - shift) * 64, 0, (t1 - t0) * 64 * sizeof(double), p->stream);
    if (e != cudaSuccess) return e;
    int slot = prof_begin(p, S2KIT_K_TABLE_GEN);
    if (p->fast) {
        int ulo = p->h_unit_first[m_lo], uhi = p->h_unit_first[m_hi], lch = table_unit_rows(p->bw);
        switch (p->bw) {
            case 16: e = table_gen_nb<16>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 32: e = table_gen_nb<32>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 64: e = table_gen_nb<64>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 128: e = table_gen_nb<128>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 256: e = table_gen_nb<256>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 512: e = table_gen_nb<512>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 1024: e = table_gen_nb<1024>(p, table, shift, ulo, uhi, lch, transposed); break;
            case 2048: e = table_gen_nb<2048>(p, table, shift, ulo, uhi, lch, transposed); break;
            default: e = cudaErrorInvalidValue;
        }
    } else {
        int nt = ((p->bw + 31) / 32) * 32;
        if (nt > 1024) nt = ((p->bw + 63) / 64) * 32;  // two nodes per thread
        k_table_gen_direct<<<m_hi - m_lo, nt, sizeof(double) * p->bw, p->stream>>>(
            table, p->d_order_start, shift, p->d_meta, p->d_rt_start, m_lo, p->bw, transposed, p->d_nodes, p->d_seeds, p->d_rec,
            p->d_q_b);
        e = cudaGetLastError();
    }
    prof_end(p, slot);
    return e;
}

cudaError_t launch_table_unpack(s2kit_cuda_plan* p, const double* table, uint64_t shift, int m, double* out) {
    k_table_unpack<<<p->bw - m, 128, 0, p->stream>>>(table, p->d_order_start, shift, p->d_meta, p->d_rt_start, m, p->bw,
                                                     out);
    return cudaGetLastError();
}

cudaError_t launch_rec_coeffs(s2kit_cuda_plan* p) {
    k_rec_coeffs<<<dim3((p->bw + 127) / 128, p->bw), 128, 0, p->stream>>>(p->d_rec, p->bw);
    return cudaGetLastError();
}

}  // namespace s2k
