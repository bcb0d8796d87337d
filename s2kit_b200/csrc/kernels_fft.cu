// kernels_fft.cu -- K1/K6 (longitude FFT) and K2/K5 (DCT-II / DCT-III along latitude) for sm_100a.
//
// K1  replaces fftw_execute_split_dft + the sqrt(2 pi)/2bw scaling          src/FST_semi_memo.c:81-90
// K2  replaces the weighting + REDFT10 + orthonormal scaling of DLTSemi    src/legendre_transform/seminaive.c:162-176
// K5  replaces the scaling + REDFT01 + sin(theta) of InvDLTSemi            src/legendre_transform/seminaive.c:92-114
//     plus the (-1)^m of negative orders and 1/sqrt(2 pi)                  src/FST_semi_memo.c:294-348
// K6  replaces the re/im-swapped fftw_execute_split_dft                    src/FST_semi_memo.c:350
//
// Layouts (private workspace): S / G spectral planes [f][part][order row m'][latitude j] (the reference's
// own transposed layout), X / V cosine planes [f][order row m'][part][k < bw].
// A DCT pair (real and imaginary column of one order) rides on ONE complex FFT of length 2bw: even/odd
// reordering v[i] = x[2i], v[n-1-i] = x[2i+1] packed as v_re + i v_im, then the two spectra are separated
// by conjugate symmetry.  Non-power-of-two bandwidths use the direct O(n^2) kernels at the end.
#include <stdlib.h>

#include <algorithm>

#include "s2k_fft.cuh"
#include "s2k_internal.cuh"

namespace s2k {

// K1 / K6 keep LT transforms in shared memory and read / write them transposed (lanes run over the LT rows first):
// a double2 row stride == 8/LT (mod 8) makes those 128-bit accesses conflict-free.
__host__ __device__ constexpr int phi_row_stride(int n, int lt) { return ((fft_padded_len(n) + 7) / 8) * 8 + (8 / lt); }


// fft_block, or for the all-radix-8 lengths its variant whose pass twiddles were requested before the data loads
template <int N>
struct FftPlan {
    static constexpr bool R8 = (ilog2(N) % 3 == 0) && N >= 512;
    FftTw8<R8 ? N : 512> f;
    __device__ __forceinline__ void prefetch(int t, const double2* __restrict__ tw) {
        if constexpr (R8) f = fft_r8_twiddles<N>(t, tw);
    }
    __device__ __forceinline__ void run(double (&xr)[8], double (&xi)[8], double2* sx, int t, int group,
                                        const double2* __restrict__ tw) {
        if constexpr (R8)
            fft_block_r8<N>(xr, xi, sx, t, group, f);
        else
            fft_block<N>(xr, xi, sx, t, group, tw);
    }
};

__device__ __forceinline__ void cp_async8_g2s(double* smem_dst, const double* gsrc) {
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a), "l"(gsrc));
}

// ------------------------------------------------------------------------------------------------ K1
// One CTA transforms LT latitude rows of one function and writes them transposed.
template <int N, int LT>
__global__ void __launch_bounds__(N / 8 * LT) k_phi_fft_fwd(const double* __restrict__ rdata,
                                                            const double* __restrict__ idata, long stride,
                                                            double* __restrict__ S, double scale, int rows_kept,
                                                            const double2* __restrict__ tw, PlaneView pv) {
    constexpr int T8 = N / 8, NT = T8 * LT;
    constexpr int RS = phi_row_stride(N, LT);  // double2 row stride: conflict-free transposed reads
    extern __shared__ double2 smem2[];
    double2* sx = smem2;
    const int tid = threadIdx.x, jj = tid / T8, t = tid % T8;
    const int j0 = blockIdx.x * LT, f = blockIdx.y;
    const double* rrow = rdata + (long)f * stride + (long)(j0 + jj) * N;
    const double* irow = idata + (long)f * stride + (long)(j0 + jj) * N;
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        xr[e] = __ldg(rrow + t + e * T8);
        xi[e] = __ldg(irow + t + e * T8);
    }
    fft_block<N>(xr, xi, sx + jj * RS, t, jj, tw);
    fft_sync<N>(jj);
#pragma unroll
    for (int e = 0; e < 8; ++e)
        sx[jj * RS + fft_pad(fft_out_index<N>(e, t))] = make_double2(xr[e] * scale, xi[e] * scale);
    __syncthreads();
    double* Sr = S + (long)f * 2 * N * N + j0;
    double* Si = Sr + pv.part_stride;
    // lanes run over the LT latitudes first (contiguous in S), then over order rows
    for (int flat = tid; flat < N * LT; flat += NT) {
        int j2 = flat % LT, mp = flat / LT;
        // REAL format never reads rows >= bw; COMPLEX skips only row bw (rows_kept encodes which)
        if (rows_kept == N ? (mp != N / 2) : (mp < rows_kept)) {
            long at = (pv.rowbase ? pv.rowbase[mp] : (long)mp * N) + j2;
            double2 v = sx[j2 * RS + fft_pad(mp)];
            Sr[at] = v.x;
            Si[at] = v.y;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K1, TMA store
// Same transform as k_phi_fft_fwd, but the transposed write goes through the TMA: each thread drops its eight outputs
// into a [order row][LT latitudes] staging tile laid out in CU_TENSOR_MAP_SWIZZLE_64B order (which also keeps those
// stores to 2-way bank conflicts), and one elected thread issues cp.async.bulk.tensor stores of 256-row boxes.  This
// takes the transposed shared-memory read and all global store instructions (half-used 64-byte runs, one request per
// 4 rows) off the LSU pipe that bounds the kernel (profiles/r1_ncu_summary.md).
__device__ __forceinline__ unsigned swz64(unsigned byte_off) { return byte_off ^ (((byte_off >> 7) & 3u) << 4); }

// PlaneView::lat_perm: latitude held at slot `pos` of a spectral-plane row
__device__ __forceinline__ int lat_row(int pos, int n, int lat_perm) {
    return !lat_perm ? pos : (pos < n / 2 ? 2 * pos : 2 * (n - 1 - pos) + 1);
}

template <int N>
__global__ void __launch_bounds__(N) k_phi_fft_fwd_tma(const double* __restrict__ rdata, const double* __restrict__ idata,
                                                        long stride, double scale, int rows_kept, int lat_perm,
                                                        const double2* __restrict__ tw, int f_plane0,
                                                        const __grid_constant__ CUtensorMap tmap) {
    constexpr int LT = 8, T8 = N / 8;
    constexpr int RS = phi_row_stride(N, LT);
    constexpr int ROWS = N < 256 ? N : 256;             // order rows per TMA box
    constexpr int BOX_BYTES = ROWS * LT * 8;            // one box = ROWS x 64 bytes
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2* sx = reinterpret_cast<double2*>(smem_raw);  // exchange rows during the FFT, staging tiles afterwards
    const int tid = threadIdx.x, jj = tid / T8, t = tid % T8;
    const int j0 = blockIdx.x * LT, f = blockIdx.y;
    const int jrow = lat_row(j0 + jj, N, lat_perm);  // grid row whose transform lands at latitude slot j0 + jj
    const double* rrow = rdata + (long)f * stride + (long)jrow * N;
    const double* irow = idata + (long)f * stride + (long)jrow * N;
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        xr[e] = __ldg(rrow + t + e * T8);
        xi[e] = __ldg(irow + t + e * T8);
    }
    fft_block<N>(xr, xi, sx + jj * RS, t, jj, tw);
    __syncthreads();  // every transform is done with the exchange rows: reuse them as staging
    // staging: part p (0 = re, 1 = im), box h: rows [h*ROWS, (h+1)*ROWS); element (row r, latitude jj) at swz64(r*64 + jj*8)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int mp = fft_out_index<N>(e, t);
        const int h = mp / ROWS, r = mp % ROWS;
        const unsigned off = swz64((unsigned)(r * 64 + jj * 8));
        *reinterpret_cast<double*>(smem_raw + (0 * (N / ROWS) + h) * BOX_BYTES + off) = xr[e] * scale;
        *reinterpret_cast<double*>(smem_raw + (1 * (N / ROWS) + h) * BOX_BYTES + off) = xi[e] * scale;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the generic-proxy writes visible to the TMA
    __syncthreads();
    if (tid == 0) {
        const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw));
        const int nbox = (rows_kept == N) ? N / ROWS : (rows_kept + ROWS - 1) / ROWS;  // REAL format: rows < bw only
        for (int part = 0; part < 2; ++part)
            for (int h = 0; h < nbox; ++h)
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                                 reinterpret_cast<uint64_t>(&tmap)),
                             "r"(j0), "r"(h * ROWS), "r"((f_plane0 + f) * 2 + part),
                             "r"(sbase + (unsigned)((part * (N / ROWS) + h) * BOX_BYTES))
                             : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging must outlive the reads
    }
}

// ------------------------------------------------------------------------------------------------ K6
template <int N, int LT>
__global__ void __launch_bounds__(N / 8 * LT) k_phi_fft_inv(const double* __restrict__ G, double* __restrict__ rdata,
                                                            double* __restrict__ idata, long stride, int real_fmt,
                                                            const double2* __restrict__ tw, PlaneView pv) {
    constexpr int T8 = N / 8, NT = T8 * LT;
    constexpr int RS = phi_row_stride(N, LT);
    extern __shared__ double2 smem2[];
    double2* sx = smem2;
    const int tid = threadIdx.x, jj = tid / T8, t = tid % T8;
    const int j0 = blockIdx.x * LT, f = blockIdx.y;
    const double* Gr = G + (long)f * 2 * N * N + j0;
    const double* Gi = Gr + pv.part_stride;
    // inverse DFT through the forward one: feed (im, re), read back (im, re)   (FST_semi_memo.c:350)
    // The transposed gather (runs of LT doubles per order row) goes straight into shared memory with 8-byte
    // cp.async: all N*LT*2 copies of the CTA are in flight at once and no registers wait on them.
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        int flat = tid + it * NT;
        int j2 = flat % LT, mp = flat / LT;
        double2* dst = sx + j2 * RS + fft_pad(mp);
        if (mp == N / 2) {
            *dst = make_double2(0.0, 0.0);  // row bw is zero by definition (FST_semi_memo.c:283-284)
        } else {
            int row = (real_fmt && mp > N / 2) ? N - mp : mp;  // conjugate mirror of row n - m' (FST_semi_memo.c:333-341)
            long at = (pv.rowbase ? pv.rowbase[row] : (long)row * N) + j2;
            cp_async8_g2s(&dst->y, Gr + at);  // real part -> imaginary slot (swap)
            cp_async8_g2s(&dst->x, Gi + at);  // imaginary part -> real slot; the mirror's sign flip happens on read
        }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int idx = fft_in_index<N>(e, t);
        double2 v = sx[jj * RS + fft_pad(idx)];
        xr[e] = (real_fmt && idx > N / 2) ? -v.x : v.x;
        xi[e] = v.y;
    }
    fft_block<N>(xr, xi, sx + jj * RS, t, jj, tw);
    double* rrow = rdata + (long)f * stride + (long)(j0 + jj) * N;
    double* irow = idata + (long)f * stride + (long)(j0 + jj) * N;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int k = fft_out_index<N>(e, t);
        irow[k] = xr[e];
        rrow[k] = xi[e];
    }
}

// ------------------------------------------------------------------------------------------------ K6, TMA load
// Mirror image of k_phi_fft_fwd_tma: the transposed gather of the LT = 8 latitude columns of every order row is done
// by cp.async.bulk.tensor loads of 256-row boxes into 64B-swizzled staging tiles (completion on an mbarrier), the
// transforms read their inputs from the staging tiles and then reuse the same shared memory for the FFT exchange.
__device__ __forceinline__ void mbar_wait(unsigned mbar_addr, unsigned phase) {
    unsigned done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(mbar_addr), "r"(phase)
            : "memory");
        if (++spins > (1ull << 28)) __trap();  // a lost TMA completion must not hang the device
    }
}

template <int N>
__global__ void __launch_bounds__(N) k_phi_fft_inv_tma(double* __restrict__ rdata, double* __restrict__ idata, long stride,
                                                        int real_fmt, int lat_perm, const double2* __restrict__ tw, int f_plane0,
                                                        const __grid_constant__ CUtensorMap tmap) {
    constexpr int LT = 8, T8 = N / 8;
    constexpr int RS = phi_row_stride(N, LT);
    constexpr int ROWS = N < 256 ? N : 256;
    constexpr int NBOX = N / ROWS;
    constexpr int BOX_BYTES = ROWS * LT * 8;
    constexpr int WORK_BYTES = (int)(sizeof(double2) * LT * RS) > 2 * NBOX * BOX_BYTES ? (int)(sizeof(double2) * LT * RS)
                                                                                       : 2 * NBOX * BOX_BYTES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2* sx = reinterpret_cast<double2*>(smem_raw);
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw));
    const unsigned mbar = sbase + WORK_BYTES;  // 8-byte mbarrier behind the work area
    const int tid = threadIdx.x, jj = tid / T8, t = tid % T8;
    const int j0 = blockIdx.x * LT, f = blockIdx.y;
    // REAL format only needs rows < bw (the rest are conjugate mirrors): one box when N = 512
    const int nbox = (real_fmt && NBOX > 1) ? NBOX / 2 : NBOX;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(2 * nbox * BOX_BYTES)
                     : "memory");
        for (int part = 0; part < 2; ++part)
            for (int h = 0; h < nbox; ++h)
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                        "r"(sbase + (unsigned)((part * NBOX + h) * BOX_BYTES)),
                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(j0), "r"(h * ROWS), "r"((f_plane0 + f) * 2 + part), "r"(mbar)
                    : "memory");
    }
    mbar_wait(mbar, 0);
    // inverse DFT through the forward one: feed (im, re), read back (im, re)   (FST_semi_memo.c:350)
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int idx = fft_in_index<N>(e, t);
        const bool mirror = real_fmt && idx > N / 2;  // conjugate mirror of row n - m' (FST_semi_memo.c:333-341)
        const int row = mirror ? N - idx : idx;
        const int h = row / ROWS, r = row % ROWS;
        const unsigned off = swz64((unsigned)(r * 64 + jj * 8));
        double vr = *reinterpret_cast<const double*>(smem_raw + (0 * NBOX + h) * BOX_BYTES + off);
        double vi = *reinterpret_cast<const double*>(smem_raw + (1 * NBOX + h) * BOX_BYTES + off);
        if (idx == N / 2) vr = vi = 0.0;  // row bw is zero by definition (FST_semi_memo.c:283-284)
        xr[e] = mirror ? -vi : vi;
        xi[e] = vr;
    }
    __syncthreads();  // every transform has its inputs: the staging tiles become the exchange rows
    fft_block<N>(xr, xi, sx + jj * RS, t, jj, tw);
    const int jrow = lat_row(j0 + jj, N, lat_perm);
    double* rrow = rdata + (long)f * stride + (long)jrow * N;
    double* irow = idata + (long)f * stride + (long)jrow * N;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int k = fft_out_index<N>(e, t);
        irow[k] = xr[e];
        rrow[k] = xi[e];
    }
}


// ------------------------------------------------------------------------------------------------ K1 / K6, large n
// At n >= 2048 a CTA holds one or two rings, so the transposed write of k_phi_fft_fwd (and the gather of k_phi_fft_inv)
// moves isolated 8- or 16-byte elements 8n bytes apart: 0.34 ms for a 268 MB pass at n = 4096, 0.8 TB/s.  Here the
// longitude transform of a ring goes to / comes from a ring-major staging plane T[part][ring][order row] with fully
// coalesced accesses straight from the FFT registers, and a separate tiled transpose moves 256-byte runs between T and
// the spectral planes (also the exchange blocks of the sharded single-field path: PlaneView::rowbase).
template <int N>
__global__ void __launch_bounds__(N / 8) k_phi_rows_fwd(const double* __restrict__ rdata, const double* __restrict__ idata,
                                                        long stride, double* __restrict__ T, double scale, int nrings,
                                                        const double2* __restrict__ tw) {
    constexpr int T8 = N / 8;
    extern __shared__ double2 smem2[];
    const int t = threadIdx.x, j = blockIdx.x, f = blockIdx.y;
    const double* rrow = rdata + (long)f * stride + (long)j * N;
    const double* irow = idata + (long)f * stride + (long)j * N;
    FftPlan<N> fp;
    fp.prefetch(t, tw);
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        xr[e] = __ldg(rrow + t + e * T8);
        xi[e] = __ldg(irow + t + e * T8);
    }
    fp.run(xr, xi, smem2, t, 0, tw);
    double* Tr = T + ((long)f * 2 * nrings + j) * N;
    double* Ti = Tr + (long)nrings * N;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = fft_out_index<N>(e, t);
        Tr[k] = xr[e] * scale;
        Ti[k] = xi[e] * scale;
    }
}

template <int N>
__global__ void __launch_bounds__(N / 8) k_phi_rows_inv(const double* __restrict__ T, double* __restrict__ rdata,
                                                        double* __restrict__ idata, long stride, int real_fmt, int nrings,
                                                        const double2* __restrict__ tw) {
    constexpr int T8 = N / 8;
    extern __shared__ double2 smem2[];
    const int t = threadIdx.x, j = blockIdx.x, f = blockIdx.y;
    const double* Tr = T + ((long)f * 2 * nrings + j) * N;
    const double* Ti = Tr + (long)nrings * N;
    // inverse DFT through the forward one: feed (im, re), read back (im, re)   (FST_semi_memo.c:350)
    FftPlan<N> fp;
    fp.prefetch(t, tw);
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int idx = fft_in_index<N>(e, t);
        const bool mirror = real_fmt && idx > N / 2;  // conjugate mirror of row n - m' (FST_semi_memo.c:333-341)
        const int row = mirror ? N - idx : idx;
        double vr = __ldg(Tr + row), vi = __ldg(Ti + row);
        if (idx == N / 2) vr = vi = 0.0;  // row bw is zero by definition (FST_semi_memo.c:283-284)
        xr[e] = mirror ? -vi : vi;
        xi[e] = vr;
    }
    fp.run(xr, xi, smem2, t, 0, tw);
    double* rrow = rdata + (long)f * stride + (long)j * N;
    double* irow = idata + (long)f * stride + (long)j * N;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = fft_out_index<N>(e, t);
        irow[k] = xr[e];
        rrow[k] = xi[e];
    }
}

// 32 x 32 tiles between T[part][ring j][order row m'] and the spectral planes S[part][row m' (rowbase)][ring j].
// to_planes = 1: T -> S (forward), 0: S -> T (inverse).  Rows a format does not use are skipped (and never read).
__global__ void __launch_bounds__(256) k_plane_transpose(double* __restrict__ T, double* __restrict__ S, int n, int nrings,
                                                         int rows_kept, int to_planes, PlaneView pv) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int mp0 = blockIdx.x * 32, j0 = blockIdx.y * 32, part = blockIdx.z & 1, f = blockIdx.z >> 1;
    double* Tp = T + ((long)f * 2 + part) * nrings * n;
    double* Sp = S + (long)f * 2 * n * n + (long)part * pv.part_stride;
    auto kept = [&](int mp) { return rows_kept == n ? (mp != n / 2) : (mp < rows_kept); };
    if (to_planes) {
#pragma unroll
        for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = Tp[(long)(j0 + ty + 8 * r) * n + mp0 + tx];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int mp = mp0 + ty + 8 * r;
            if (kept(mp)) Sp[(pv.rowbase ? pv.rowbase[mp] : (long)mp * n) + j0 + tx] = tile[tx][ty + 8 * r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int mp = mp0 + ty + 8 * r;
            tile[ty + 8 * r][tx] = kept(mp) ? Sp[(pv.rowbase ? pv.rowbase[mp] : (long)mp * n) + j0 + tx] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) Tp[(long)(j0 + ty + 8 * r) * n + mp0 + tx] = tile[tx][ty + 8 * r];
    }
}

// ------------------------------------------------------------------------------------------------ K2
__device__ __forceinline__ int ridx_to_row(int ridx, int bw) { return ridx < bw ? ridx : ridx + 1; }

// One CTA handles FPB (function, order row) pairs; each pair = re and im column -> one complex FFT.
template <int N, int FPB, bool PEER>
__global__ void __launch_bounds__(N / 8 * FPB) k_dct_fwd(const double* __restrict__ S, double* __restrict__ X,
                                                         const double* __restrict__ weights, int ridx_lo, int ridx_hi,
                                                         const double2* __restrict__ tw,
                                                         const double2* __restrict__ qtab, PlaneView pv, const __grid_constant__ PeerSegs peers) {
    constexpr int T8 = N / 8, B = N / 2, NP = fft_padded_len(N);
    extern __shared__ double2 smem2[];
    const int tid = threadIdx.x, g = tid / T8, t = tid % T8;
    double2* sx = smem2 + g * NP;
    const int ridx = ridx_lo + blockIdx.x * FPB + g, f = blockIdx.y;
    const bool live = ridx < ridx_hi;
    const int rsel = live ? ridx : ridx_lo;
    const int mp = pv.rowlist ? pv.rowlist[rsel] : ridx_to_row(rsel, B);
    const int m = mp < B ? mp : N - mp;
    const double* w = weights + ((m & 1) ? N : 0);
    const long rowoff = (long)(pv.rowlist ? rsel : mp) * pv.lrow_stride;
    const double* Sr = S + (long)f * 2 * N * N + rowoff;
    const double* Si = Sr + pv.part_stride;
    FftPlan<N> fp;
    fp.prefetch(t, tw);
    double xr[8], xi[8];
    if constexpr (PEER) {
        // Single field over the GPUs of one process: segment s of the row lives in peer s's memory -- the ring -> order
        // exchange IS these loads.  The row (2 x N doubles) is pulled into the transform's exchange buffer with coalesced
        // 16-byte cp.async copies (512 contiguous bytes per warp and instruction: NVLink-friendly requests, all of a
        // thread's copies in flight at once), then read from there in the DCT's even/odd order.  (Loading the reordered
        // 8-byte elements straight from peer memory made the 2-GPU transform slower than one GPU: 2.7 vs 2.3 ms.)
        double* st = reinterpret_cast<double*>(sx);
        for (int q = t; q < N; q += T8) {
            const int part = q / (N / 2), j = (q % (N / 2)) * 2;
            const double* src = peers.ptr[j >> pv.seg_shift] + rowoff + (j & pv.seg_mask) + (long)part * pv.part_stride;
            unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(st + part * N + j));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        fft_sync<N>(g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int p = t + e * T8;
            const double wj = __ldg(w + p);
            const int j = (p < B) ? 2 * p : 2 * (N - 1 - p) + 1;
            xr[e] = st[j] * wj;
            xi[e] = st[N + j] * wj;
        }
        fft_sync<N>(g);  // the staging area becomes the exchange buffer
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int p = t + e * T8;
            double wj = __ldg(w + p);  // weights are stored in load order (s2k_host_reordered)
            long at = p;               // lat_perm: the row already is in load order
            if (!pv.lat_perm) {
                int j = (p < B) ? 2 * p : 2 * (N - 1 - p) + 1;
                at = seg_offset(pv, j);
            }
            xr[e] = __ldg(Sr + at) * wj;
            xi[e] = __ldg(Si + at) * wj;
        }
    }
    fp.run(xr, xi, sx, t, g, tw);
    // Separation of the two real spectra needs Z[k] and Z[n-k], k < bw: Z[k] is still in this thread's registers, so
    // only the upper half of the spectrum (indices > bw) goes through shared memory -- half a write and half a read
    // per point instead of a full write and two reads.
    constexpr int R = fft_last_radix(N);
    fft_sync<N>(g);
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (fft_slot<R>(e) >= 4) sx[fft_pad(fft_out_index<N>(e, t) - B)] = make_double2(xr[e], xi[e]);
    fft_sync<N>(g);
    if (!live) return;
    double* Xr = X + (((long)f * N + mp) * 2) * B;
    double* Xi = Xr + B;
    const double s_all = 1.0 / sqrt(2.0 * (double)N);  // 1/sqrt(2*size), seminaive.c:174
    // consecutive lanes hold consecutive k: even and odd k go to slot k/2 of their half of the parity-split plane, so
    // a warp writes two contiguous 128-byte runs per store
    constexpr int HALF = B / 2;
    const double2 q0 = __ldg(qtab + t);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (fft_slot<R>(e) >= 4) continue;
        const int k = fft_out_index<N>(e, t);
        const double ar = xr[e], ai = xi[e];
        double br = ar, bi = ai;  // k = 0: Z[n] = Z[0]
        if (k != 0) {
            const double2 zb = sx[fft_pad(B - k)];  // Z[n - k] sits at (n - k) - bw
            br = zb.x;
            bi = zb.y;
        }
        // V1 = (Z[k] + conj Z[n-k]) / 2 ; V2 = (Z[k] - conj Z[n-k]) / 2i ; REDFT10 = 2 Re(e^{-i pi k/2n} V)
        const double2 q = quarter_rot(q0, fft_slot<R>(e));  // k = t + slot n/8
        double y1 = q.x * (ar + br) + q.y * (ai - bi);
        double y2 = q.x * (ai + bi) - q.y * (ar - br);
        if (k == 0) {
            y1 *= 0.70710678118654752440;  // M_SQRT1_2, seminaive.c:173
            y2 *= 0.70710678118654752440;
        }
        const int slot = (k & 1) * HALF + (k >> 1);
        Xr[slot] = y1 * s_all;
        Xi[slot] = y2 * s_all;
    }
}

// ------------------------------------------------------------------------------------------------ K5
template <int N, int FPB, bool PEER>
__global__ void __launch_bounds__(N / 8 * FPB) k_dct_inv(const double* __restrict__ V, double* __restrict__ G,
                                                         const double* __restrict__ sinv, int ridx_lo, int ridx_hi,
                                                         double out_scale, const double2* __restrict__ tw,
                                                         const double2* __restrict__ qtab, PlaneView pv, const __grid_constant__ PeerSegs peers) {
    constexpr int T8 = N / 8, B = N / 2, NP = fft_padded_len(N);
    extern __shared__ double2 smem2[];
    const int tid = threadIdx.x, g = tid / T8, t = tid % T8;
    double2* sx = smem2 + g * NP;
    const int ridx = ridx_lo + blockIdx.x * FPB + g, f = blockIdx.y;
    const bool live = ridx < ridx_hi;
    const int rsel = live ? ridx : ridx_lo;
    const int mp = pv.rowlist ? pv.rowlist[rsel] : ridx_to_row(rsel, B);
    const int m = mp < B ? mp : N - mp;
    const double* Va = V + (((long)f * N + mp) * 2) * B;
    const double* Vb = Va + B;
    const double c_rest = 1.0 / sqrt(2.0 * (double)N);  // 0.5/sqrt(bw), seminaive.c:72
    const double c_zero = 1.0 / sqrt((double)N);        // fcos[0] / sqrt(2 bw), seminaive.c:98
    // e^{i pi k / 2n} for k = t + e n/8 is e^{i pi t / 2n} e^{i pi e / 16}: one table load and seven constant rotations
    // instead of eight 16-byte loads -- K5 sits at 95 % of the LSU pipe with the FP64 pipe half idle
    // (profiles/r1_ncu_full_metrics_final2.csv)
    const double2 q0 = __ldg(qtab + t);
    FftPlan<N> fp;
    fp.prefetch(t, tw);
    double xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int k = t + e * T8;
        // W[k] = e^{i pi k/2n} (Xa[k] - i Xa[n-k]) + i (same for b), X[k >= bw] = 0
        double wr = 0.0, wi = 0.0;
        if (k != B) {
            int src = k < B ? k : N - k;
            double sc = (src == 0) ? c_zero : c_rest;
            double a = __ldg(Va + cos_slot(src, B)) * sc, b = __ldg(Vb + cos_slot(src, B)) * sc;
            const double2 q = quarter_rot(q0, e);
            double ur = (k < B) ? a : b, ui = (k < B) ? b : -a;  // (a + ib) or -i (a + ib)
            wr = q.x * ur - q.y * ui;
            wi = q.x * ui + q.y * ur;
        }
        xr[e] = wi;  // swapped: inverse DFT through the forward transform
        xi[e] = wr;
    }
    fp.run(xr, xi, sx, t, g, tw);
    double sign = ((mp > B) && (m & 1)) ? -out_scale : out_scale;  // (-1)^m for negative orders
    const long rowoff = (long)(pv.rowlist ? rsel : mp) * pv.lrow_stride;
    if constexpr (PEER) {
        // order -> ring exchange as NVLink stores into the ring owners' receive blocks (multi.cu): the row is put into
        // natural latitude order in the exchange buffer first, so the stores are coalesced 16-byte pieces of each peer's
        // contiguous run instead of scattered 8-byte elements
        double* st = reinterpret_cast<double*>(sx);
        fft_sync<N>(g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = fft_out_index<N>(e, t);
            const double sc = (m & 1) ? __ldg(sinv + i) * sign : sign;
            const int j = (i < B) ? 2 * i : 2 * (N - 1 - i) + 1;
            st[j] = xi[e] * sc;      // Re z -> column a (real part)
            st[N + j] = xr[e] * sc;  // Im z -> column b (imaginary part)
        }
        fft_sync<N>(g);
        if (live)
            for (int q = t; q < N; q += T8) {
                const int part = q / (N / 2), j = (q % (N / 2)) * 2;
                double* dst = const_cast<double*>(peers.ptr[j >> pv.seg_shift]) + rowoff + (j & pv.seg_mask) +
                              (long)part * pv.part_stride;
                *reinterpret_cast<double2*>(dst) = *reinterpret_cast<const double2*>(st + part * N + j);
            }
        return;
    }
    if (!live) return;
    double* Gr = G + (long)f * 2 * N * N + rowoff;
    double* Gi = Gr + pv.part_stride;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int i = fft_out_index<N>(e, t);
        double s = (m & 1) ? __ldg(sinv + i) * sign : sign;  // sines stored in output order (s2k_host_reordered)
        long at = i;  // lat_perm: the row is kept in output order
        if (!pv.lat_perm) {
            int j = (i < B) ? 2 * i : 2 * (N - 1 - i) + 1;
            at = seg_offset(pv, j);
        }
        Gr[at] = xi[e] * s;  // Re z -> column a (real part)
        Gi[at] = xr[e] * s;  // Im z -> column b (imaginary part)
    }
}

// ------------------------------------------------------------------------------------------------ direct kernels
// Any bandwidth: one thread per output, exact-index twiddles.  Used when bw is not a power of two (or < 16).
__global__ void k_direct_phi_fwd(const double* __restrict__ rdata, const double* __restrict__ idata, long stride,
                                 double* __restrict__ S, int n, double scale, const double2* __restrict__ tw) {
    int j = blockIdx.x, f = blockIdx.y;
    const double* rr = rdata + (long)f * stride + (long)j * n;
    const double* ii = idata + (long)f * stride + (long)j * n;
    for (int mp = threadIdx.x; mp < n; mp += blockDim.x) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < n; ++k) {
            double2 w = tw[(int)(((long)mp * k) % n)];
            double a = rr[k], b = ii[k];
            sr += a * w.x - b * w.y;
            si += a * w.y + b * w.x;
        }
        S[((long)f * 2 * n + mp) * n + j] = sr * scale;
        S[((long)f * 2 * n + n + mp) * n + j] = si * scale;
    }
}

__global__ void k_direct_phi_inv(const double* __restrict__ G, double* __restrict__ rdata, double* __restrict__ idata,
                                 long stride, int n, int real_fmt, const double2* __restrict__ tw) {
    int j = blockIdx.x, f = blockIdx.y;
    const double* Gr = G + (long)f * 2 * n * n;
    const double* Gi = Gr + (long)n * n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double sr = 0.0, si = 0.0;
        for (int mp = 0; mp < n; ++mp) {
            if (2 * mp == n) continue;
            double a, b;
            if (real_fmt && 2 * mp > n) {
                a = Gr[(long)(n - mp) * n + j];
                b = -Gi[(long)(n - mp) * n + j];
            } else {
                a = Gr[(long)mp * n + j];
                b = Gi[(long)mp * n + j];
            }
            double2 w = tw[(int)(((long)mp * k) % n)];  // (cos, -sin): conjugate for e^{+i}
            sr += a * w.x + b * w.y;
            si += b * w.x - a * w.y;
        }
        rdata[(long)f * stride + (long)j * n + k] = sr;
        idata[(long)f * stride + (long)j * n + k] = si;
    }
}

__global__ void k_direct_dct_fwd(const double* __restrict__ S, double* __restrict__ X,
                                 const double* __restrict__ weights, int bw, int ridx_lo,
                                 const double2* __restrict__ qtab) {
    int n = 2 * bw;
    int mp = ridx_to_row(ridx_lo + blockIdx.x, bw), f = blockIdx.y;
    int m = mp < bw ? mp : n - mp;
    const double* w = weights + ((m & 1) ? n : 0);
    double s_all = 1.0 / sqrt(2.0 * (double)n);
    for (int o = threadIdx.x; o < 2 * bw; o += blockDim.x) {
        int part = o / bw, k = o % bw;
        const double* col = S + (((long)f * 2 + part) * n + mp) * n;
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += col[j] * w[j] * qtab[(int)(((long)(2 * j + 1) * k) % (4 * n))].x;
        acc *= 2.0;
        if (k == 0) acc *= 0.70710678118654752440;
        X[(((long)f * n + mp) * 2 + part) * bw + cos_slot(k, bw)] = acc * s_all;
    }
}

__global__ void k_direct_dct_inv(const double* __restrict__ V, double* __restrict__ G, const double* __restrict__ sinv,
                                 int bw, int ridx_lo, double out_scale, const double2* __restrict__ qtab) {
    int n = 2 * bw;
    int mp = ridx_to_row(ridx_lo + blockIdx.x, bw), f = blockIdx.y;
    int m = mp < bw ? mp : n - mp;
    double c_rest = 0.5 / sqrt((double)bw), c_zero = 1.0 / sqrt((double)n);
    double sign = ((mp > bw) && (m & 1)) ? -out_scale : out_scale;
    for (int o = threadIdx.x; o < 2 * n; o += blockDim.x) {
        int part = o / n, j = o % n;
        const double* v = V + (((long)f * n + mp) * 2 + part) * bw;
        double acc = v[0] * c_zero;
        for (int k = 1; k < bw; ++k) acc += 2.0 * (v[cos_slot(k, bw)] * c_rest) * qtab[(int)(((long)(2 * j + 1) * k) % (4 * n))].x;
        double s = (m & 1) ? sinv[j] * sign : sign;
        G[(((long)f * 2 + part) * n + mp) * n + j] = acc * s;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    return ensure_smem(reinterpret_cast<const void*>(kernel), bytes);
}

// S2KIT_CUDA_PHI_ROWS: 0 = always the transposing single-kernel K1 / K6, otherwise the smallest n that takes the staged path
static bool phi_rows_enabled(int n) {
    static const int from = [] {
        const char* e = getenv("S2KIT_CUDA_PHI_ROWS");
        return e ? atoi(e) : 4096;  // measured: n = 4096 K1 0.35 -> 0.21 ms, K6 0.30 -> 0.22 ms; n = 2048 no gain, n = 1024 slower
    }();
    return from > 0 && n >= from;
}
// ring-major staging plane of the large-n longitude transforms, as large as the spectral workspace; created on first use
static cudaError_t ensure_phi_stage(s2kit_cuda_plan* p) {
    if (p->d_T) return cudaSuccess;
    return cudaMalloc((void**)&p->d_T, sizeof(double) * (size_t)p->chunk * 2 * p->n * p->n);
}

template <int N>
static cudaError_t phi_fwd_n(s2kit_cuda_plan* p, const double* rdata, const double* idata, long stride, double* S,
                             int nfun, int rows_kept, const PlaneView& pv, int nrings) {
    constexpr int LT = (4096 / N) < 8 ? (4096 / N) : 8;
    if constexpr (N <= 512) {
        // ordinary plane in the plan's own workspace: transposed write through the TMA
        const long plane2 = 2L * N * N, offS = S - p->d_S;
        if (tma_planes_ok(p, nfun) && offS >= 0 && offS % plane2 == 0 && offS / plane2 + nfun <= p->chunk && !pv.rowbase &&
            nrings == N) {
            constexpr int RSX = phi_row_stride(N, 8);
            size_t smem = std::max(sizeof(double2) * 8 * RSX, (size_t)2 * N * 8 * 8);
            cudaError_t e = set_smem(k_phi_fft_fwd_tma<N>, smem);
            if (e != cudaSuccess) return e;
            double scale = sqrt(2.0 * M_PI) / (double)N;
            k_phi_fft_fwd_tma<N><<<dim3(N / 8, nfun), N, smem, p->stream>>>(rdata, idata, stride, scale, rows_kept,
                                                                              pv.lat_perm, p->d_tw_n, (int)(offS / plane2),
                                                                              p->tma_S);
            return cudaGetLastError();
        }
    }
    if (pv.lat_perm) return cudaErrorInvalidValue;  // only the TMA variant writes the reordered latitude layout
    if constexpr (N >= 1024) {
        if (phi_rows_enabled(N) && nrings % 32 == 0 && nfun <= p->chunk) {
            // ring-major staging plane + tiled transpose (see k_phi_rows_fwd)
            cudaError_t e = ensure_phi_stage(p);
            if (e != cudaSuccess) return e;
            const size_t smem_r = sizeof(double2) * fft_padded_len(N);
            e = set_smem(k_phi_rows_fwd<N>, smem_r);
            if (e != cudaSuccess) return e;
            k_phi_rows_fwd<N><<<dim3(nrings, nfun), N / 8, smem_r, p->stream>>>(rdata, idata, stride, p->d_T,
                                                                                 sqrt(2.0 * M_PI) / (double)N, nrings, p->d_tw_n);
            k_plane_transpose<<<dim3(N / 32, nrings / 32, 2 * nfun), 256, 0, p->stream>>>(p->d_T, S, N, nrings, rows_kept, 1, pv);
            return cudaGetLastError();
        }
    }
    constexpr int RS = phi_row_stride(N, LT);
    size_t smem = sizeof(double2) * LT * RS;
    cudaError_t e = set_smem(k_phi_fft_fwd<N, LT>, smem);
    if (e != cudaSuccess) return e;
    double scale = sqrt(2.0 * M_PI) / (double)N;  // FST_semi_memo.c:86
    if (nrings % LT) return cudaErrorInvalidValue;
    k_phi_fft_fwd<N, LT><<<dim3(nrings / LT, nfun), N / 8 * LT, smem, p->stream>>>(rdata, idata, stride, S, scale,
                                                                                   rows_kept, p->d_tw_n, pv);
    return cudaGetLastError();
}

template <int N>
static cudaError_t phi_inv_n(s2kit_cuda_plan* p, const double* G, double* rdata, double* idata, long stride, int nfun,
                             int real_fmt, const PlaneView& pv, int nrings) {
    constexpr int LT = (4096 / N) < 8 ? (4096 / N) : 8;
    if constexpr (N <= 512) {
        const long plane2 = 2L * N * N, offG = G - p->d_S;
        if (tma_planes_ok(p, nfun) && offG >= 0 && offG % plane2 == 0 && offG / plane2 + nfun <= p->chunk && !pv.rowbase &&
            nrings == N) {
            constexpr int RSX = phi_row_stride(N, 8);
            size_t smem = std::max(sizeof(double2) * 8 * RSX, (size_t)2 * N * 8 * 8) + 16;
            cudaError_t e = set_smem(k_phi_fft_inv_tma<N>, smem);
            if (e != cudaSuccess) return e;
            k_phi_fft_inv_tma<N><<<dim3(N / 8, nfun), N, smem, p->stream>>>(rdata, idata, stride, real_fmt, pv.lat_perm,
                                                                              p->d_tw_n, (int)(offG / plane2), p->tma_S);
            return cudaGetLastError();
        }
    }
    if (pv.lat_perm) return cudaErrorInvalidValue;
    if constexpr (N >= 1024) {
        if (phi_rows_enabled(N) && nrings % 32 == 0 && nfun <= p->chunk) {
            cudaError_t e = ensure_phi_stage(p);
            if (e != cudaSuccess) return e;
            const size_t smem_r = sizeof(double2) * fft_padded_len(N);
            e = set_smem(k_phi_rows_inv<N>, smem_r);
            if (e != cudaSuccess) return e;
            k_plane_transpose<<<dim3(N / 32, nrings / 32, 2 * nfun), 256, 0, p->stream>>>(p->d_T, const_cast<double*>(G), N, nrings,
                                                                                          real_fmt ? N / 2 : N, 0, pv);
            k_phi_rows_inv<N><<<dim3(nrings, nfun), N / 8, smem_r, p->stream>>>(p->d_T, rdata, idata, stride, real_fmt, nrings,
                                                                                 p->d_tw_n);
            return cudaGetLastError();
        }
    }
    constexpr int RS = phi_row_stride(N, LT);
    size_t smem = sizeof(double2) * LT * RS;
    cudaError_t e = set_smem(k_phi_fft_inv<N, LT>, smem);
    if (e != cudaSuccess) return e;
    if (nrings % LT) return cudaErrorInvalidValue;
    k_phi_fft_inv<N, LT><<<dim3(nrings / LT, nfun), N / 8 * LT, smem, p->stream>>>(G, rdata, idata, stride, real_fmt,
                                                                                   p->d_tw_n, pv);
    return cudaGetLastError();
}

template <int N>
static cudaError_t dct_fwd_n(s2kit_cuda_plan* p, const double* S, double* X, int nfun, int lo, int hi,
                             const PlaneView& pv) {
    constexpr int T8 = N / 8;
    constexpr int FPB = (256 / T8) < 1 ? 1 : ((256 / T8) > 8 ? 8 : (256 / T8));
    size_t smem = sizeof(double2) * FPB * fft_padded_len(N);
    const dim3 grid((hi - lo + FPB - 1) / FPB, nfun);
    if (pv.peers) {
        cudaError_t e = set_smem(k_dct_fwd<N, FPB, true>, smem);
        if (e != cudaSuccess) return e;
        k_dct_fwd<N, FPB, true><<<grid, T8 * FPB, smem, p->stream>>>(S, X, p->d_wv, lo, hi, p->d_tw_n, p->d_q_n, pv,
                                                                     *pv.peers);
        return cudaGetLastError();
    }
    cudaError_t e = set_smem(k_dct_fwd<N, FPB, false>, smem);
    if (e != cudaSuccess) return e;
    k_dct_fwd<N, FPB, false><<<grid, T8 * FPB, smem, p->stream>>>(S, X, p->d_wv, lo, hi, p->d_tw_n, p->d_q_n, pv,
                                                                  PeerSegs());
    return cudaGetLastError();
}

template <int N>
static cudaError_t dct_inv_n(s2kit_cuda_plan* p, const double* V, double* G, int nfun, int lo, int hi,
                             const PlaneView& pv) {
    if constexpr (N == 512) {
        if (fft16_enabled() && !pv.peers) return launch_dct_inv16(p, V, G, nfun, lo, hi, pv);
    }
    constexpr int T8 = N / 8;
    constexpr int FPB = (256 / T8) < 1 ? 1 : ((256 / T8) > 8 ? 8 : (256 / T8));
    size_t smem = sizeof(double2) * FPB * fft_padded_len(N);
    double out_scale = 1.0 / sqrt(2.0 * M_PI);  // FST_semi_memo.c:344
    const dim3 grid((hi - lo + FPB - 1) / FPB, nfun);
    if (pv.peers) {
        cudaError_t e = set_smem(k_dct_inv<N, FPB, true>, smem);
        if (e != cudaSuccess) return e;
        k_dct_inv<N, FPB, true><<<grid, T8 * FPB, smem, p->stream>>>(V, G, p->d_sv, lo, hi, out_scale, p->d_tw_n,
                                                                     p->d_q_n, pv, *pv.peers);
        return cudaGetLastError();
    }
    cudaError_t e = set_smem(k_dct_inv<N, FPB, false>, smem);
    if (e != cudaSuccess) return e;
    k_dct_inv<N, FPB, false><<<grid, T8 * FPB, smem, p->stream>>>(V, G, p->d_sv, lo, hi, out_scale, p->d_tw_n, p->d_q_n,
                                                                  pv, PeerSegs());
    return cudaGetLastError();
}

#define S2K_DISPATCH_N(n, CALL)                  \
    switch (n) {                                 \
        case 32: return CALL(32);                \
        case 64: return CALL(64);                \
        case 128: return CALL(128);              \
        case 256: return CALL(256);              \
        case 512: return CALL(512);              \
        case 1024: return CALL(1024);            \
        case 2048: return CALL(2048);            \
        case 4096: return CALL(4096);            \
        default: return cudaErrorInvalidValue;   \
    }

bool tma_planes_ok(const s2kit_cuda_plan* p, int nfun) {
    return p->tma_S_ok && p->fast && p->n >= 64 && p->n <= 512 && nfun <= p->chunk;
}

PlaneView default_view(int n) {
    PlaneView v;
    v.rowbase = nullptr;
    v.rowlist = nullptr;
    v.part_stride = (long)n * n;
    v.lrow_stride = n;
    v.seg_stride = 0;
    v.seg_shift = 30;  // j >> 30 == 0: a row is one segment
    v.seg_mask = 0x3fffffff;
    v.nrings = n;
    v.lat_perm = 0;
    return v;
}

cudaError_t launch_phi_fft_fwd(s2kit_cuda_plan* p, const double* rdata, const double* idata, long stride, double* S,
                               int nfun, int data_format, const PlaneView* view) {
    int n = p->n;
    const PlaneView pv = view ? *view : default_view(n);
    int slot = prof_begin(p, S2KIT_K_PHI_FFT_FWD);
    cudaError_t e;
    if (p->fast) {
        int rows_kept = (data_format == S2KIT_REAL) ? p->bw : n;
#define CALL(NN) phi_fwd_n<NN>(p, rdata, idata, stride, S, nfun, rows_kept, pv, pv.nrings)
        e = [&]() -> cudaError_t { S2K_DISPATCH_N(n, CALL) }();
#undef CALL
    } else {
        int nt = n < 256 ? ((n + 31) / 32) * 32 : 256;
        k_direct_phi_fwd<<<dim3(n, nfun), nt, 0, p->stream>>>(rdata, idata, stride, S, n,
                                                              sqrt(2.0 * M_PI) / (double)n, p->d_tw_n);
        e = cudaGetLastError();
    }
    prof_end(p, slot);
    return e;
}

cudaError_t launch_phi_fft_inv(s2kit_cuda_plan* p, const double* G, double* rdata, double* idata, long stride, int nfun,
                               int data_format, const PlaneView* view) {
    int n = p->n;
    const PlaneView pv = view ? *view : default_view(n);
    int real_fmt = data_format == S2KIT_REAL;
    int slot = prof_begin(p, S2KIT_K_PHI_FFT_INV);
    cudaError_t e;
    if (p->fast) {
#define CALL(NN) phi_inv_n<NN>(p, G, rdata, idata, stride, nfun, real_fmt, pv, pv.nrings)
        e = [&]() -> cudaError_t { S2K_DISPATCH_N(n, CALL) }();
#undef CALL
    } else {
        int nt = n < 256 ? ((n + 31) / 32) * 32 : 256;
        k_direct_phi_inv<<<dim3(n, nfun), nt, 0, p->stream>>>(G, rdata, idata, stride, n, real_fmt, p->d_tw_n);
        e = cudaGetLastError();
    }
    prof_end(p, slot);
    return e;
}

// rows are addressed by ridx in [0, 2bw-1): ridx < bw -> order row ridx, else row ridx + 1 (row bw unused)
cudaError_t launch_dct_fwd(s2kit_cuda_plan* p, const double* S, double* X, int nfun, int row_lo, int row_hi,
                           int data_format, const PlaneView* view) {
    (void)data_format;
    const PlaneView pv = view ? *view : default_view(p->n);
    if (row_hi <= row_lo) return cudaSuccess;
    int slot = prof_begin(p, S2KIT_K_DCT_FWD);
    cudaError_t e;
    if (p->fast) {
#define CALL(NN) dct_fwd_n<NN>(p, S, X, nfun, row_lo, row_hi, pv)
        e = [&]() -> cudaError_t { S2K_DISPATCH_N(p->n, CALL) }();
#undef CALL
    } else {
        int nt = 2 * p->bw < 256 ? ((2 * p->bw + 31) / 32) * 32 : 256;
        k_direct_dct_fwd<<<dim3(row_hi - row_lo, nfun), nt, 0, p->stream>>>(S, X, p->d_weights, p->bw, row_lo,
                                                                            p->d_q_n);
        e = cudaGetLastError();
    }
    prof_end(p, slot);
    return e;
}

cudaError_t launch_dct_inv(s2kit_cuda_plan* p, const double* V, double* G, int nfun, int row_lo, int row_hi,
                           int data_format, const PlaneView* view) {
    (void)data_format;
    const PlaneView pv = view ? *view : default_view(p->n);
    if (row_hi <= row_lo) return cudaSuccess;
    int slot = prof_begin(p, S2KIT_K_DCT_INV);
    cudaError_t e;
    if (p->fast) {
#define CALL(NN) dct_inv_n<NN>(p, V, G, nfun, row_lo, row_hi, pv)
        e = [&]() -> cudaError_t { S2K_DISPATCH_N(p->n, CALL) }();
#undef CALL
    } else {
        int nt = 256;
        k_direct_dct_inv<<<dim3(row_hi - row_lo, nfun), nt, 0, p->stream>>>(V, G, p->d_sin, p->bw, row_lo,
                                                                            1.0 / sqrt(2.0 * M_PI), p->d_q_n);
        e = cudaGetLastError();
    }
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
