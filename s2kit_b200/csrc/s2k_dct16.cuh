// s2k_dct16.cuh -- the two DCTs of the seminaive Legendre transform at bw = 256 on the one-warp 512-point FFT
// (s2k_fft16.cuh), written so that the lane-level pieces also run on the host for 32 emulated lanes
// (tests/host_checks/dct16_check.cu):
//
//   forward (DLTSemi, src/legendre_transform/seminaive.c:162-176): the weighted real and imaginary column of one order
//   ride on ONE complex FFT (even/odd reordering), the two real spectra are separated with Z[k] and Z[N-k] -- which
//   always live in the same warp, so the separation is a shuffle -- and the first bw cosine coefficients of both
//   columns are written straight into the contraction's shared-memory panel.
//
// The FFT's single register <-> shared-memory exchange runs IN PLACE inside the warp's own two panel columns (2 x 2
// parities x CS doubles = exactly the 16 x 33 doubles of the exchange, done once for the real and once for the
// imaginary parts), so the persistent kernels need no exchange buffers beside their panels.
#pragma once
#include "s2k_fft16.cuh"

namespace s2k {

// ---- in-place exchange: element (k1, t) of the 16 x 32 exchange lives in the parity-(k1 >> 3) copy of the warp's column
// pair, at (k1 & 7) * 33 + t doubles from that copy's start; the two copies are `ps` doubles apart with ps == 8 (mod 16),
// which keeps the 64-bit reads of a half-warp on distinct banks.
__host__ __device__ constexpr int d16_ex_write(int t, int k1, int ps) { return (k1 >> 3) * ps + (k1 & 7) * 33 + t; }
__host__ __device__ constexpr int d16_ex_read(int lane, int j, int ps) {
    return ((lane & 15) >> 3) * ps + ((lane & 15) & 7) * 33 + (lane >> 4) + 2 * j;
}

// ---- Z[N - k] for the outputs k < 256 a lane finishes.  Lane (k1, h) holds Z[k], k = k1 + 16 (qi + 8 h), in register
// o = 2 qi.  In shuffle round qi it READS lane d16_src_lane(lane, qi) and every lane OFFERS register d16_offer_reg(lane, qi):
//   k1 != 0 : Z[N-k] sits in lane (16 - k1, 1 - h), register 2 (7 - qi) + 1
//   k1 == 0 : k = 16 (qi + 8h); Z[N-k] sits in lane (0, h'), register 2 qi' + 1 with 8 h' + qi' = 16 - qi - 8h
//             (k = 0 pairs with itself: the caller uses Z[0])
// Lanes 0 and 16 are only ever read by each other (or themselves), so they may offer a different register than the rest.
__host__ __device__ constexpr int d16_src_lane(int lane, int qi) {
    const int k1 = lane & 15, h = lane >> 4;
    if (k1) return (16 - k1) + 16 * (1 - h);
    const int q2 = 16 - qi - 8 * h;  // in [1, 16]; 16 only for k = 0 (unused)
    return q2 >= 16 ? lane : 16 * (q2 >> 3);
}
__host__ __device__ constexpr int d16_offer_reg(int lane, int qi) {
    if (lane & 15) return 2 * (7 - qi) + 1;
    return qi == 0 ? 1 : 2 * (8 - qi) + 1;
}

// cos / sin (pi qi / 64), qi < 8: (cos, sin)(pi k / 2n) for k = k1 + 128 h + 16 qi (n = 512) is the lane's table entry
// (cos, sin)(pi (k1 + 128 h) / 2n) rotated by these
__host__ __device__ inline void d16_quarter_rot(double q0r, double q0i, int qi, double& qr, double& qim) {
    constexpr double C[8] = {1.0, 0.99879545620517239271, 0.99518472667219688624, 0.98917650996478097345,
                             0.98078528040323044913, 0.97003125319454399260, 0.95694033573220886494, 0.94154406518302077841};
    constexpr double S[8] = {0.0, 0.04906767432741801426, 0.09801714032956060199, 0.14673047445536175166,
                             0.19509032201612826785, 0.24298017990326388995, 0.29028467725446236764, 0.33688985339222005069};
    qr = q0r * C[qi] - q0i * S[qi];
    qim = q0r * S[qi] + q0i * C[qi];
}

// DCT-II separation for one output index kk < 256 (seminaive.c:170-176 after the FFT): (ar, ai) = Z[kk], (br, bi) = Z[N-kk]
// (Z[0] for kk = 0), (qr, qi) = (cos, sin)(pi kk / 2n), s_all = 1 / sqrt(2 * 2bw).  y1 / y2: cosine coefficient kk of the
// real / imaginary column.
__host__ __device__ inline void d16_separate(double ar, double ai, double br, double bi, double qr, double qi, int kk,
                                             double s_all, double& y1, double& y2) {
    y1 = qr * (ar + br) + qi * (ai - bi);
    y2 = qr * (ai + bi) - qi * (ar - br);
    if (kk == 0) {
        y1 *= 0.70710678118654752440;  // M_SQRT1_2, seminaive.c:173
        y2 *= 0.70710678118654752440;
    }
    y1 *= s_all;
    y2 *= s_all;
}

// panel slot of cosine index kk inside a column: parity-split, parity copies `ps` doubles apart
__host__ __device__ constexpr int d16_panel_slot(int kk, int ps) { return (kk & 1) * ps + (kk >> 1); }

#ifdef __CUDACC__
// 512-point FFT of the calling warp with the register <-> shared-memory exchange done IN PLACE in the warp's own column
// pair (col0: parity-0 copy, the parity-1 copy ps doubles further), real parts first, then imaginary parts.  Same
// contract as f16_fft512_warp: register e holds x[lane + 32 e] on entry, register o holds X[f16_out_index(lane, o)] on
// exit.  The column pair's previous contents must already be in registers (or dead); its contents afterwards are garbage.
__device__ __forceinline__ void d16_fft512_inplace(double (&xr)[16], double (&xi)[16], double* col0, int ps, int lane,
                                                   const double2* __restrict__ tw) {
    const double2 w1 = __ldg(tw + lane);
    f16_phase1(xr, xi, w1.x, w1.y);
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) col0[d16_ex_write(lane, k1, ps)] = xr[k1];
    __syncwarp();
    const double* rd = col0 + d16_ex_read(lane, 0, ps);
#pragma unroll
    for (int j = 0; j < 16; ++j) xr[j] = rd[2 * j];
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) col0[d16_ex_write(lane, k1, ps)] = xi[k1];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) xi[j] = rd[2 * j];
    __syncwarp();  // the column pair is free again
    f16_dft16(xr, xi);
    const int h = lane >> 4;
    double pr[8], pi_[8];
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
        const double sr = h ? xr[qi] : xr[8 + qi], si = h ? xi[qi] : xi[8 + qi];
        pr[qi] = __shfl_xor_sync(0xffffffffu, sr, 16);
        pi_[qi] = __shfl_xor_sync(0xffffffffu, si, 16);
    }
    f16_phase3(xr, xi, pr, pi_, h);
}

// The whole forward DCT pair for the calling warp.  On entry xr/xi[e] = weighted samples at reordered position
// p = lane + 32 e (real / imaginary column).  col0: the warp's first panel column, parity-0 copy (the second column is CS
// doubles further, the parity-1 copies ps doubles further).  On exit the first 256 cosine coefficients of both columns
// are in the panel and the pad slots [128, CS) of the four column copies are zero.  Only __syncwarp inside.
template <int CS>
__device__ __forceinline__ void d16_dct2_pair_to_panel(double (&xr)[16], double (&xi)[16], double* col0, int ps, int lane,
                                                       const double2* __restrict__ tw, const double2* __restrict__ qtab,
                                                       double s_all) {
    static_assert(2 * CS == 8 * 33, "the exchange fills the column pair exactly");
    d16_fft512_inplace(xr, xi, col0, ps, lane, tw);
    const int h = lane >> 4, k1 = lane & 15;
    // separation + write: output kk = k1 + 16 (qi + 8h) -> parity kk & 1 = k1 & 1, slot (k1 >> 1) + 8 qi + 64 h
    const double2 q0 = __ldg(qtab + k1 + 128 * h);
    double* out = col0 + (k1 & 1) * ps + (k1 >> 1) + 64 * h;
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
        // every lane offers the register its reader wants (static indices, one select)
        const double offr = k1 ? xr[2 * (7 - qi) + 1] : xr[qi == 0 ? 1 : 2 * (8 - qi) + 1];
        const double offi = k1 ? xi[2 * (7 - qi) + 1] : xi[qi == 0 ? 1 : 2 * (8 - qi) + 1];
        const int src = d16_src_lane(lane, qi);
        double br = __shfl_sync(0xffffffffu, offr, src), bi = __shfl_sync(0xffffffffu, offi, src);
        const double ar = xr[2 * qi], ai = xi[2 * qi];
        const int kk = k1 + 16 * (qi + 8 * h);
        if (kk == 0) {
            br = ar;
            bi = ai;
        }
        double qr, qim, y1, y2;
        d16_quarter_rot(q0.x, q0.y, qi, qr, qim);
        d16_separate(ar, ai, br, bi, qr, qim, kk, s_all, y1, y2);
        out[8 * qi] = y1;
        out[8 * qi + CS] = y2;
    }
    // the exchange left garbage in the pad slots of the four column copies
    if (lane < 16) col0[(lane >> 3) * ps + ((lane >> 2) & 1) * CS + (CS - 4) + (lane & 3)] = 0.0;
}
#endif

}  // namespace s2k
