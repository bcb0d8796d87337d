// s2k_fft16.cuh -- 512-point FP64 complex FFT owned by ONE warp, 16 points per thread.
//
// Successor of the 8-points-per-thread block FFT of s2k_fft.cuh for the kernels that are bound by shared-memory / LSU
// wavefronts (profiles/r1_ncu_summary.md section 5): one shared-memory exchange and one shuffle stage instead of two
// shared-memory exchanges and four named barriers, and the whole transform lives in one warp (__syncwarp only).
// The index algebra is modelled in tools/fft16_model.py; tests/host_checks/fft16_check.cu runs the very functions of this
// header for 32 emulated lanes on the host against a long-double DFT.
//
//   phase 1   lane t holds x[t + 32 e], e < 16         radix-16 DFT over e, then the twiddle W512^(t k1)
//   exchange  (t, k1) -> lane L = k1 + 16 h, slot j, with t = h + 2 j            (the one shared-memory round trip)
//   phase 2   radix-16 DFT over j                       B[k1][h][q]
//   shuffle   lanes (k1, 0) <-> (k1, 1) swap the halves q >= 8 / q < 8 they do not finish themselves
//   phase 3   X[k1 + 16 q + 256 r] = B0[q] + (-1)^r W32^q B1[q]     lane h finishes q = 8 h + qi, qi < 8, r = 0, 1
// Output register o = 2 qi + r of lane L holds X[f16_out_index(L, o)].
#pragma once
#include <cuda_runtime.h>

namespace s2k {

constexpr int F16_N = 512;
constexpr int F16_EX_STRIDE = 33;                       // double2 elements per k1 row: 32 lanes + 1 pad (odd: conflict-free reads)
constexpr int F16_EX_ELEMS = 16 * F16_EX_STRIDE;        // exchange buffer of one transform (double2)

__host__ __device__ constexpr int f16_in_index(int t, int e) { return t + 32 * e; }
__host__ __device__ constexpr int f16_ex_write(int t, int k1) { return k1 * F16_EX_STRIDE + t; }
__host__ __device__ constexpr int f16_ex_read(int lane, int j) { return (lane & 15) * F16_EX_STRIDE + (lane >> 4) + 2 * j; }
__host__ __device__ constexpr int f16_out_index(int lane, int o) {
    return (lane & 15) + 16 * ((o >> 1) + 8 * (lane >> 4)) + 256 * (o & 1);
}

// ---- 4-point and 16-point DFTs in registers (forward sign), natural order in and out
__host__ __device__ inline void f16_dft4(double& r0, double& i0, double& r1, double& i1, double& r2, double& i2, double& r3,
                                         double& i3) {
    const double ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    const double cr = r1 + r3, ci = i1 + i3;
    const double dr = i1 - i3, di = r3 - r1;  // (x1 - x3) * (-i)
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + dr; i1 = bi + di;
    r3 = br - dr; i3 = bi - di;
}

// X[c + 4 d] = sum_b W4^(b d) W16^(b c) sum_a W4^(a c) x[4 a + b]
__host__ __device__ inline void f16_dft16(double (&xr)[16], double (&xi)[16]) {
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173;  // cos, sin(pi/8)
    constexpr double H = 0.70710678118654752440;
    // step 1: for every b, DFT-4 over a (inputs 4a + b); result Y[b][c] stored at 4c + b
#pragma unroll
    for (int b = 0; b < 4; ++b) f16_dft4(xr[b], xi[b], xr[4 + b], xi[4 + b], xr[8 + b], xi[8 + b], xr[12 + b], xi[12 + b]);
    // step 2: Y[b][c] *= W16^(b c), W16 = e^{-2 pi i / 16}
    auto rot = [](double& r, double& i, double c, double s) {  // times (c - i s)
        const double t = r * c + i * s;
        i = i * c - r * s;
        r = t;
    };
    // c = 1: b = 1, 2, 3 -> W16^1, W16^2, W16^3
    rot(xr[4 + 1], xi[4 + 1], C1, S1);
    rot(xr[4 + 2], xi[4 + 2], H, H);
    rot(xr[4 + 3], xi[4 + 3], S1, C1);
    // c = 2: W16^2, W16^4 = -i, W16^6
    rot(xr[8 + 1], xi[8 + 1], H, H);
    {
        const double t = xr[8 + 2];
        xr[8 + 2] = xi[8 + 2];
        xi[8 + 2] = -t;
    }
    rot(xr[8 + 3], xi[8 + 3], -H, H);
    // c = 3: W16^3, W16^6, W16^9
    rot(xr[12 + 1], xi[12 + 1], S1, C1);
    rot(xr[12 + 2], xi[12 + 2], -H, H);
    rot(xr[12 + 3], xi[12 + 3], -C1, -S1);
    // step 3: for every c, DFT-4 over b (inputs at 4c + b); result X[c + 4 d] lands at 4c + d
#pragma unroll
    for (int c = 0; c < 4; ++c)
        f16_dft4(xr[4 * c], xi[4 * c], xr[4 * c + 1], xi[4 * c + 1], xr[4 * c + 2], xi[4 * c + 2], xr[4 * c + 3],
                 xi[4 * c + 3]);
    // register 4c + d holds X[c + 4d]: transpose the 4 x 4 index
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            double t = xr[4 * c + d];
            xr[4 * c + d] = xr[4 * d + c];
            xr[4 * d + c] = t;
            t = xi[4 * c + d];
            xi[4 * c + d] = xi[4 * d + c];
            xi[4 * d + c] = t;
        }
}

// phase 1 of lane t: registers e -> registers k1, including the inter-phase twiddle W512^(t k1).
// w1 = (cos, -sin)(2 pi t / 512), the lane's entry of the twiddle table.
__host__ __device__ inline void f16_phase1(double (&xr)[16], double (&xi)[16], double w1r, double w1i) {
    f16_dft16(xr, xi);
    double wr = w1r, wi = w1i;  // W^(t k1), k1 = 1, 2, ...
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        const double a = xr[k1], b = xi[k1];
        xr[k1] = a * wr - b * wi;
        xi[k1] = a * wi + b * wr;
        if (k1 < 15) {
            const double nr = wr * w1r - wi * w1i;
            wi = wr * w1i + wi * w1r;
            wr = nr;
        }
    }
}

// phase 3 of lane (k1, h): own = B[k1][h][.] after phase 2 (16 values), other = the partner lane's 8 values for q = 8h + qi.
// On exit register o = 2 qi + r holds X[k1 + 16 (8h + qi) + 256 r].
__host__ __device__ inline void f16_phase3(double (&xr)[16], double (&xi)[16], const double (&pr)[8], const double (&pi_)[8],
                                           int h) {
    // W32^qi, qi < 8 (cos, -sin)(2 pi qi / 32); for h = 1 the twiddle is W32^(qi + 8) = -i W32^qi
    constexpr double TC[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                              0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    constexpr double TS[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                              0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
    double outr[16], outi[16];
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
        // b0 = even-t half (lanes h = 0), b1 = odd-t half (lanes h = 1), both at q = 8h + qi
        const double ownr = h ? xr[8 + qi] : xr[qi], owni = h ? xi[8 + qi] : xi[qi];  // static register indices
        const double b0r = h ? pr[qi] : ownr, b0i = h ? pi_[qi] : owni;
        const double b1r = h ? ownr : pr[qi], b1i = h ? owni : pi_[qi];
        double tr = TC[qi], ti = -TS[qi];  // W32^qi
        if (h) {                            // times -i
            const double t = tr;
            tr = ti;
            ti = -t;
        }
        const double mr = b1r * tr - b1i * ti, mi = b1r * ti + b1i * tr;
        outr[2 * qi] = b0r + mr;
        outi[2 * qi] = b0i + mi;
        outr[2 * qi + 1] = b0r - mr;
        outi[2 * qi + 1] = b0i - mi;
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        xr[o] = outr[o];
        xi[o] = outi[o];
    }
}

// (cos, sin)(pi k / 2n) for k = t + 32 e (n = 512) is (cos, sin)(pi t / 2n) rotated by pi e / 32: the DCT kernels load the
// lane's entry of the quarter-wave table once and rotate it by these constants
__host__ __device__ inline void f16_quarter_rot(double q0r, double q0i, int e, double& qr, double& qi) {
    constexpr double C[16] = {1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
                              0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708, 0.77301045336273696081,
                              0.70710678118654752440, 0.63439328416364549822, 0.55557023301960222474, 0.47139673682599764856,
                              0.38268343236508977173, 0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199};
    constexpr double S[16] = {0.0, 0.09801714032956060199, 0.19509032201612826785, 0.29028467725446236764,
                              0.38268343236508977173, 0.47139673682599764856, 0.55557023301960222474, 0.63439328416364549822,
                              0.70710678118654752440, 0.77301045336273696081, 0.83146961230254523708, 0.88192126434835502971,
                              0.92387953251128675613, 0.95694033573220886494, 0.98078528040323044913, 0.99518472667219688624};
    qr = q0r * C[e] - q0i * S[e];
    qi = q0r * S[e] + q0i * C[e];
}

#ifdef __CUDACC__
// The whole transform for the calling warp.  On entry register e holds x[f16_in_index(lane, e)]; on exit register o
// holds X[f16_out_index(lane, o)].  ex: this warp's exchange buffer (F16_EX_ELEMS double2); tw: (cos, -sin)(2 pi q / 512).
// The caller must not touch ex concurrently; a __syncwarp() separates the buffer's reuse from the reads here.
__device__ __forceinline__ void f16_fft512_warp(double (&xr)[16], double (&xi)[16], double2* ex, int lane,
                                                const double2* __restrict__ tw) {
    const double2 w1 = __ldg(tw + lane);
    f16_phase1(xr, xi, w1.x, w1.y);
    __syncwarp();  // earlier readers of ex are done
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) ex[f16_ex_write(lane, k1)] = make_double2(xr[k1], xi[k1]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const double2 v = ex[f16_ex_read(lane, j)];
        xr[j] = v.x;
        xi[j] = v.y;
    }
    f16_dft16(xr, xi);
    // shuffle stage: lane (k1, h) sends the half it does not finish (q in [8 (1-h), 8 (1-h) + 8)) to lane (k1, 1-h)
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
#endif

}  // namespace s2k
