// s2k_fft.cuh -- block-level FP64 complex FFT building block for sm_100a.
//
// Replaces FFTW's role in the reference (fftw_execute_split_dft at src/FST_semi_memo.c:81,350 and the
// REDFT10/REDFT01 plans at src/legendre_transform/seminaive.c:107,170, src/legendre_polynomials/cospml.c:203-224).
//
// Design: a length-N transform (N = 2^k, 8 <= N <= 4096) is owned by N/8 threads; every thread keeps
// 8 complex points in registers for the whole transform.  Passes are Stockham autosort radix-8 (plus one
// final radix-2/4 pass when log2 N is not a multiple of 3), so the result comes out in natural order and
// the first pass reads / the last pass leaves the SAME eight slots {t + s*N/8}: the first pass can be fed
// straight from global memory and the last pass can be consumed straight from registers.  Between
// passes the points are exchanged through shared memory as interleaved double2 (128-bit accesses: half the LSU
// instructions of split re/im arrays), index padded by i>>3, which keeps every pass within 1.08x of conflict-free.
// The threads of one transform synchronise on their own named barrier when they are whole warps.  The base twiddle of each butterfly
// comes from a host-computed exact table W[q] = (cos 2 pi q/N, -sin 2 pi q/N).
#pragma once
#include <cuda_runtime.h>

namespace s2k {

// shared-memory exchange rows hold double2 elements; index padding for 128-bit accesses (8 lanes per wavefront)
__host__ __device__ constexpr int fft_pad(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int fft_padded_len(int n) { return n + (n >> 3) + 1; }

// barrier among the N/8 threads of transform `group` of the CTA (named barrier 1 + group when they are whole
// warps, the CTA-wide barrier otherwise -- then every thread of the CTA must take part)
template <int N>
__device__ __forceinline__ void fft_sync(int group) {
    if constexpr (N / 8 == 32) {
        (void)group;
        __syncwarp();  // the transform is exactly one warp (groups are cut at multiples of N/8 threads)
    } else if constexpr (N / 8 > 32) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(N / 8) : "memory");
    } else {
        (void)group;
        __syncthreads();
    }
}

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
// last-pass radix: 8 if log2 N % 3 == 0, else 2 or 4
__host__ __device__ constexpr int fft_last_radix(int n) {
    return (ilog2(n) % 3 == 0) ? 8 : (1 << (ilog2(n) % 3));
}
__host__ __device__ constexpr int fft_num_r8(int n) { return ilog2(n) / 3; }

// ---- small in-register DFTs (forward sign), natural order out
__device__ __forceinline__ void dft2(double* xr, double* xi) {
    double ar = xr[0], ai = xi[0];
    xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
    xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
}

__device__ __forceinline__ void dft4(double* xr, double* xi) {
    double t0r = xr[0] + xr[2], t0i = xi[0] + xi[2];
    double t1r = xr[0] - xr[2], t1i = xi[0] - xi[2];
    double t2r = xr[1] + xr[3], t2i = xi[1] + xi[3];
    // (x1 - x3) * (-i) = (im, -re)
    double t3r = xi[1] - xi[3], t3i = xr[3] - xr[1];
    xr[0] = t0r + t2r; xi[0] = t0i + t2i;
    xr[2] = t0r - t2r; xi[2] = t0i - t2i;
    xr[1] = t1r + t3r; xi[1] = t1i + t3i;
    xr[3] = t1r - t3r; xi[3] = t1i - t3i;
}

__device__ __forceinline__ void dft8(double* xr, double* xi) {
    constexpr double H = 0.70710678118654752440;
    double er[4] = {xr[0], xr[2], xr[4], xr[6]}, ei[4] = {xi[0], xi[2], xi[4], xi[6]};
    double orr[4] = {xr[1], xr[3], xr[5], xr[7]}, oi[4] = {xi[1], xi[3], xi[5], xi[7]};
    dft4(er, ei);
    dft4(orr, oi);
    // odd part times w8^k: w8 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
    double r1 = (orr[1] + oi[1]) * H, i1 = (oi[1] - orr[1]) * H;
    double r2 = oi[2], i2 = -orr[2];
    double r3 = (oi[3] - orr[3]) * H, i3 = -(orr[3] + oi[3]) * H;
    xr[0] = er[0] + orr[0]; xi[0] = ei[0] + oi[0];
    xr[4] = er[0] - orr[0]; xi[4] = ei[0] - oi[0];
    xr[1] = er[1] + r1; xi[1] = ei[1] + i1;
    xr[5] = er[1] - r1; xi[5] = ei[1] - i1;
    xr[2] = er[2] + r2; xi[2] = ei[2] + i2;
    xr[6] = er[2] - r2; xi[6] = ei[2] - i2;
    xr[3] = er[3] + r3; xi[3] = ei[3] + i3;
    xr[7] = er[3] - r3; xi[7] = ei[3] - i3;
}

template <int R>
__device__ __forceinline__ void dftR(double* xr, double* xi) {
    if constexpr (R == 8) dft8(xr, xi);
    else if constexpr (R == 4) dft4(xr, xi);
    else dft2(xr, xi);
}

// Register e of a pass with radix R belongs to butterfly q = e / R, leg r = e % R and sits at slot
// s = q + (8/R) r, i.e. at index t + s * N/8.
template <int R>
__host__ __device__ constexpr int fft_slot(int e) { return (e / R) + (8 / R) * (e % R); }

// natural index of register e after the whole transform / before the first pass (first radix is 8)
template <int N>
__device__ __forceinline__ int fft_out_index(int e, int t) {
    constexpr int R = fft_last_radix(N);
    return t + fft_slot<R>(e) * (N / 8);
}
template <int N>
__device__ __forceinline__ int fft_in_index(int e, int t) { return t + e * (N / 8); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Twiddles: ONE table load per butterfly (w = W^(k*STEP)); the higher powers w^2..w^7 are formed by at most
// three chained complex multiplications.  (Loading all seven from the table cost 7 scattered 16-byte L1
// requests per thread and pass and saturated the LSU data pipe -- profiles/r1_ncu_summary.md.)
// TWMUL: the table passed in belongs to a transform TWMUL times longer (W_N^q = table[q * TWMUL])
template <int N, int NS, int R, int TWMUL = 1>
__device__ __forceinline__ void fft_pass_compute(double (&xr)[8], double (&xi)[8], int t,
                                                 const double2* __restrict__ tw) {
    constexpr int T8 = N / 8, NB = 8 / R, STEP = TWMUL * (N / (NS * R));
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        if constexpr (NS > 1) {
            int k = (t + q * T8) & (NS - 1);
            double2 w[R];
            w[1] = __ldg(&tw[k * STEP]);
            if constexpr (R >= 4) {
                w[2] = cmul(w[1], w[1]);
                w[3] = cmul(w[2], w[1]);
            }
            if constexpr (R == 8) {
                w[4] = cmul(w[2], w[2]);
                w[5] = cmul(w[4], w[1]);
                w[6] = cmul(w[3], w[3]);
                w[7] = cmul(w[4], w[3]);
            }
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double a = xr[q * R + r], b = xi[q * R + r];
                xr[q * R + r] = a * w[r].x - b * w[r].y;
                xi[q * R + r] = a * w[r].y + b * w[r].x;
            }
        }
        dftR<R>(&xr[q * R], &xi[q * R]);
    }
}

template <int N, int NS, int R>
__device__ __forceinline__ void fft_pass_write(const double (&xr)[8], const double (&xi)[8], double2* sx, int t) {
    constexpr int T8 = N / 8, NB = 8 / R;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        int j = t + q * T8;
        int k = j & (NS - 1);
        int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            sx[fft_pad(base + r * NS)] = make_double2(xr[q * R + r], xi[q * R + r]);
        }
    }
}

template <int N, int R>
__device__ __forceinline__ void fft_pass_read(double (&xr)[8], double (&xi)[8], const double2* sx, int t) {
    constexpr int T8 = N / 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        double2 v = sx[fft_pad(t + fft_slot<R>(e) * T8)];
        xr[e] = v.x;
        xi[e] = v.y;
    }
}

template <int N, int NS, int LEFT, int TWMUL = 1>
struct FftRest {
    // LEFT = number of radix-8 passes still to run after the first one
    __device__ static __forceinline__ void run(double (&xr)[8], double (&xi)[8], double2* sx, int t, int group,
                                               const double2* __restrict__ tw) {
        if constexpr (LEFT > 0) {
            fft_sync<N>(group);
            fft_pass_write<N, NS / 8, 8>(xr, xi, sx, t);
            fft_sync<N>(group);
            fft_pass_read<N, 8>(xr, xi, sx, t);
            fft_pass_compute<N, NS, 8, TWMUL>(xr, xi, t, tw);
            FftRest<N, NS * 8, LEFT - 1, TWMUL>::run(xr, xi, sx, t, group, tw);
        } else if constexpr (NS < N) {
            constexpr int R = N / NS;  // 2 or 4
            fft_sync<N>(group);
            fft_pass_write<N, NS / 8, 8>(xr, xi, sx, t);
            fft_sync<N>(group);
            fft_pass_read<N, R>(xr, xi, sx, t);
            fft_pass_compute<N, NS, R, TWMUL>(xr, xi, t, tw);
        }
    }
};


// ---- all-radix-8 transforms (N = 512, 4096) with the pass twiddles loaded UP FRONT ---------------------------------
// In fft_block every pass starts with a table load whose latency (an L2 round trip under load) sits between two
// barriers, on the critical path of a CTA that owns one long transform.  Here the (at most three) base twiddles of a
// thread are requested before the first pass, next to the data loads, and are long there when the passes need them.
template <int N, int NS>
__device__ __forceinline__ void fft_r8_pass_w(double (&xr)[8], double (&xi)[8], double2 w1) {
    double2 w[8];
    w[1] = w1;
    w[2] = cmul(w[1], w[1]);
    w[3] = cmul(w[2], w[1]);
    w[4] = cmul(w[2], w[2]);
    w[5] = cmul(w[4], w[1]);
    w[6] = cmul(w[3], w[3]);
    w[7] = cmul(w[4], w[3]);
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const double a = xr[r], b = xi[r];
        xr[r] = a * w[r].x - b * w[r].y;
        xi[r] = a * w[r].y + b * w[r].x;
    }
    dft8(xr, xi);
}
template <int N>
struct FftTw8 {
    double2 w[3];
};
template <int N>
__device__ __forceinline__ FftTw8<N> fft_r8_twiddles(int t, const double2* __restrict__ tw) {
    static_assert(ilog2(N) % 3 == 0 && N >= 64 && N <= 4096, "N = 64, 512 or 4096");
    FftTw8<N> f;
    f.w[0] = __ldg(&tw[(t & 7) * (N / 64)]);
    f.w[1] = (N >= 512) ? __ldg(&tw[(t & 63) * (N / 512)]) : make_double2(1.0, 0.0);
    f.w[2] = (N >= 4096) ? __ldg(&tw[(t & 511) * (N / 4096)]) : make_double2(1.0, 0.0);
    return f;
}
template <int N>
__device__ __forceinline__ void fft_block_r8(double (&xr)[8], double (&xi)[8], double2* sx, int t, int group,
                                             const FftTw8<N>& f) {
    dft8(xr, xi);
    fft_sync<N>(group);
    fft_pass_write<N, 1, 8>(xr, xi, sx, t);
    fft_sync<N>(group);
    fft_pass_read<N, 8>(xr, xi, sx, t);
    fft_r8_pass_w<N, 8>(xr, xi, f.w[0]);
    if constexpr (N >= 512) {
        fft_sync<N>(group);
        fft_pass_write<N, 8, 8>(xr, xi, sx, t);
        fft_sync<N>(group);
        fft_pass_read<N, 8>(xr, xi, sx, t);
        fft_r8_pass_w<N, 64>(xr, xi, f.w[1]);
    }
    if constexpr (N >= 4096) {
        fft_sync<N>(group);
        fft_pass_write<N, 64, 8>(xr, xi, sx, t);
        fft_sync<N>(group);
        fft_pass_read<N, 8>(xr, xi, sx, t);
        fft_r8_pass_w<N, 512>(xr, xi, f.w[2]);
    }
}

// (cos, sin)(pi k / 2n) for k = t + s n/8 is (cos, sin)(pi t / 2n) rotated by pi s / 16: the DCT kernels load one entry
// of the quarter-wave table per thread and rotate it by these constants instead of loading eight entries (K5 sat at
// 95 % of the LSU pipe with the FP64 pipe half idle, profiles/r1_ncu_full_metrics_final2.csv).
__device__ __forceinline__ double2 quarter_rot(double2 q0, int s) {
    constexpr double QC[8] = {1.0, 0.9807852804032304491, 0.9238795325112867561, 0.8314696123025452371,
                              0.7071067811865475244, 0.5555702330196022247, 0.3826834323650897717, 0.1950903220161282678};
    constexpr double QS[8] = {0.0, 0.1950903220161282678, 0.3826834323650897717, 0.5555702330196022247,
                              0.7071067811865475244, 0.8314696123025452371, 0.9238795325112867561, 0.9807852804032304491};
    if (s == 0) return q0;
    return make_double2(q0.x * QC[s] - q0.y * QS[s], q0.x * QS[s] + q0.y * QC[s]);
}

// Forward DFT of N points spread over N/8 threads.  On entry register e holds x[t + e*N/8]; on exit it
// holds X[fft_out_index<N>(e, t)].  sx: this transform's padded exchange row (fft_padded_len(N) double2).
// Synchronises with fft_sync<N>(group): all N/8 threads of the transform must call it (for N < 256 every thread
// of the CTA, the same number of times).
template <int N, int TWMUL = 1>
__device__ __forceinline__ void fft_block(double (&xr)[8], double (&xi)[8], double2* sx, int t, int group,
                                          const double2* __restrict__ tw) {
    static_assert(N >= 8 && (N & (N - 1)) == 0, "power of two >= 8");
    fft_pass_compute<N, 1, 8, TWMUL>(xr, xi, t, tw);
    FftRest<N, 8, fft_num_r8(N) - 1, TWMUL>::run(xr, xi, sx, t, group, tw);
}

}  // namespace s2k
