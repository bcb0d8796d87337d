// kernels_fft16.cu -- K5 (DCT-III along latitude) at bw = 256 on the one-warp 512-point FFT of s2k_fft16.cuh
// (S2KIT_CUDA_FFT16=0 falls back to k_dct_inv).  The arithmetic pieces are checked on the host
// (tests/host_checks/fft16_check.cu), the kernel by the GPU parity tests.
//
// Same mathematics as k_dct_inv (kernels_fft.cu; InvDLTSemi, src/legendre_transform/seminaive.c:92-114, and the
// (-1)^m / 1/sqrt(2 pi) of InvFSTSemiMemo, src/FST_semi_memo.c:294-348), different distribution: one warp per
// (function, order row) pair, 16 points per lane, one shared-memory exchange per transform instead of two and no
// named barriers.  k_dct_inv sits at 95 % of the LSU pipe (profiles/r1_ncu_full_metrics_final2.csv); this layout moves
// about 290 LSU wavefronts per transform instead of 460.
#include <stdlib.h>

#include "s2k_fft16.cuh"
#include "s2k_internal.cuh"

namespace s2k {

constexpr int F16_WARPS = 8;  // transforms per CTA

__global__ void __launch_bounds__(F16_WARPS * 32, 2) k_dct_inv16(const double* __restrict__ V, double* __restrict__ G,
                                                                const double* __restrict__ sinv, int ridx_lo, int ridx_hi,
                                                                double out_scale, const double2* __restrict__ tw,
                                                                const double2* __restrict__ qtab, PlaneView pv) {
    constexpr int N = F16_N, B = N / 2;
    extern __shared__ __align__(16) double2 smem16[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* ex = smem16 + warp * F16_EX_ELEMS;
    const int ridx = ridx_lo + blockIdx.x * F16_WARPS + warp, f = blockIdx.y;
    if (ridx >= ridx_hi) return;  // warp-uniform; nothing below synchronises across warps
    const int mp = pv.rowlist ? pv.rowlist[ridx] : (ridx < B ? ridx : ridx + 1);
    const int m = mp < B ? mp : N - mp;
    const double* Va = V + (((long)f * N + mp) * 2) * B;
    const double* Vb = Va + B;
    const double c_rest = 1.0 / sqrt(2.0 * (double)N);  // 0.5/sqrt(bw), seminaive.c:72
    const double c_zero = 1.0 / sqrt((double)N);        // fcos[0] / sqrt(2 bw), seminaive.c:98
    const double2 q0 = __ldg(qtab + lane);
    double xr[16], xi[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int k = f16_in_index(lane, e);
        // W[k] = e^{i pi k/2n} (Xa[k] - i Xa[n-k]) + i (same for b), X[k >= bw] = 0
        double wr = 0.0, wi = 0.0;
        if (k != B) {
            const int src = k < B ? k : N - k;
            const double sc = (src == 0) ? c_zero : c_rest;
            const double a = __ldg(Va + cos_slot(src, B)) * sc, b = __ldg(Vb + cos_slot(src, B)) * sc;
            double qr, qi;
            f16_quarter_rot(q0.x, q0.y, e, qr, qi);
            const double ur = (k < B) ? a : b, ui = (k < B) ? b : -a;  // (a + ib) or -i (a + ib)
            wr = qr * ur - qi * ui;
            wi = qr * ui + qi * ur;
        }
        xr[e] = wi;  // swapped: inverse DFT through the forward transform
        xi[e] = wr;
    }
    f16_fft512_warp(xr, xi, ex, lane, tw);
    const double sign = ((mp > B) && (m & 1)) ? -out_scale : out_scale;  // (-1)^m for negative orders
    const long rowoff = (long)(pv.rowlist ? ridx : mp) * pv.lrow_stride;
    double* Gr = G + (long)f * 2 * N * N + rowoff;
    double* Gi = Gr + pv.part_stride;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        const int i = f16_out_index(lane, o);
        const double s = (m & 1) ? __ldg(sinv + i) * sign : sign;  // sines stored in output order (s2k_host_reordered)
        long at = i;  // lat_perm: the row is kept in output order
        if (!pv.lat_perm) {
            const int j = (i < B) ? 2 * i : 2 * (N - 1 - i) + 1;
            at = seg_offset(pv, j);
        }
        Gr[at] = xi[o] * s;  // Re z -> column a (real part)
        Gi[at] = xr[o] * s;  // Im z -> column b (imaginary part)
    }
}

bool fft16_enabled() {
    static int on = [] {
        const char* e = getenv("S2KIT_CUDA_FFT16");
        return (e && e[0] == '0') ? 0 : 1;  // default on: 1248 -> 1203 us per 1024 functions at bw = 256 (0.79 -> 0.82 of HBM peak)
    }();
    return on != 0;
}

cudaError_t launch_dct_inv16(s2kit_cuda_plan* p, const double* V, double* G, int nfun, int lo, int hi, const PlaneView& pv) {
    const size_t smem = sizeof(double2) * F16_WARPS * F16_EX_ELEMS;
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_dct_inv16), smem);
    if (e != cudaSuccess) return e;
    k_dct_inv16<<<dim3((hi - lo + F16_WARPS - 1) / F16_WARPS, nfun), F16_WARPS * 32, smem, p->stream>>>(
        V, G, p->d_sv, lo, hi, 1.0 / sqrt(2.0 * M_PI), p->d_tw_n, p->d_q_n, pv);
    return cudaGetLastError();
}

}  // namespace s2k
