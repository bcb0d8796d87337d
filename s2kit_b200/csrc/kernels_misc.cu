// kernels_misc.cu -- K8 (zonal row sums, spectral multiply) and the FP64 / HBM peak micro-benchmarks.
//
// zonal row sums replace the m = 0 loop of FZTSemiMemo   src/FST_semi_memo.c:386-399
// spectral multiply replaces TransMult / ComplexMult     src/util/util.c:23-27,68-103
#include "s2k_internal.cuh"

namespace s2k {

// One warp per latitude row: r0[j] = (sqrt(2 pi)/2bw) sum_k data[j][k], written as order row 0 of S.
__global__ void k_zonal_rowsum(const double* __restrict__ rdata, const double* __restrict__ idata, long stride,
                               double* __restrict__ S, int n, double scale) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int f = blockIdx.y;
    if (warp >= n) return;
    const double* rr = rdata + (long)f * stride + (long)warp * n;
    const double* ii = idata + (long)f * stride + (long)warp * n;
    double sr = 0.0, si = 0.0;
    for (int k = lane; k < n; k += 32) {
        sr += rr[k];
        si += ii[k];
    }
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if (lane == 0) {
        S[((long)f * 2 * n + 0) * n + warp] = sr * scale;
        S[((long)f * 2 * n + n) * n + warp] = si * scale;
    }
}

// res(m,l) = sqrt(4 pi/(2l+1)) * (x u - y v, x v - y u) with (x,y) = filter(l), (u,v) = data(m,l)
__global__ void k_spectral_mul(const double* __restrict__ rd, const double* __restrict__ id, long coef_stride,
                               const double* __restrict__ rf, const double* __restrict__ ifl, long filt_stride,
                               double* __restrict__ rres, double* __restrict__ ires, long res_stride, int bw) {
    int f = blockIdx.y;
    int total = bw * bw;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        // invert IndexOfHarmonicCoeff: non-negative orders occupy the first bw(bw+1)/2 slots
        int npos = bw * (bw + 1) / 2, l;
        if (idx < npos) {
            // m*bw - m(m-1)/2 <= idx: solve by float guess + fix-up
            int m = (int)(((2.0 * bw + 1.0) - sqrt((2.0 * bw + 1.0) * (2.0 * bw + 1.0) - 8.0 * idx)) * 0.5);
            while (m > 0 && m * bw - (m * (m - 1)) / 2 > idx) --m;
            while ((m + 1) * bw - ((m + 1) * m) / 2 <= idx) ++m;
            l = m + (idx - (m * bw - (m * (m - 1)) / 2));
        } else {
            // orders -(bw-1) .. -1: order -a starts at npos + (bw-1-a)(bw-a)/2 and has bw-a entries
            int rel = idx - npos;
            int t = (int)((sqrt(8.0 * rel + 1.0) - 1.0) * 0.5);  // t = bw-1-a
            while (t > 0 && t * (t + 1) / 2 > rel) --t;
            while ((t + 1) * (t + 2) / 2 <= rel) ++t;
            int a = bw - 1 - t;
            l = a + (rel - t * (t + 1) / 2);
        }
        double x = rf[(long)f * filt_stride + l], y = ifl[(long)f * filt_stride + l];
        double u = rd[(long)f * coef_stride + idx], v = id[(long)f * coef_stride + idx];
        double s = sqrt(4.0 * M_PI / (2.0 * l + 1.0));
        rres[(long)f * res_stride + idx] = (x * u - y * v) * s;
        ires[(long)f * res_stride + idx] = (x * v - y * u) * s;  // sign as in util.c:26
    }
}

cudaError_t launch_zonal_rowsum(s2kit_cuda_plan* p, const double* rdata, const double* idata, long stride, double* S,
                                int nfun) {
    int n = p->n;
    int slot = prof_begin(p, S2KIT_K_ZONAL);
    int warps_per_block = 8;
    k_zonal_rowsum<<<dim3((n + warps_per_block - 1) / warps_per_block, nfun), warps_per_block * 32, 0, p->stream>>>(
        rdata, idata, stride, S, n, sqrt(2.0 * M_PI) / (double)n);
    cudaError_t e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

cudaError_t launch_spectral_mul(s2kit_cuda_plan* p, const double* rd, const double* id, long coef_stride,
                                const double* rf, const double* ifl, long filt_stride, double* rres, double* ires,
                                long res_stride, int nfun) {
    int slot = prof_begin(p, S2KIT_K_SPECTRAL_MUL);
    int total = p->bw * p->bw;
    k_spectral_mul<<<dim3((total + 255) / 256, nfun), 256, 0, p->stream>>>(rd, id, coef_stride, rf, ifl, filt_stride,
                                                                           rres, ires, res_stride, p->bw);
    cudaError_t e = cudaGetLastError();
    prof_end(p, slot);
    return e;
}

// ------------------------------------------------------------------------------------------------ naive DLT
// DLTNaive / InvDLTNaive (src/legendre_transform/naive.c:35-60, 77-95): dense products with the caller's theta-space
// table pml[(l - m)][j], j < 2bw.  One warp per degree forward (fixed shuffle-tree summation), one thread per sample
// point inverse (degrees accumulated in the reference's order).
__global__ void k_naive_dlt(const double* __restrict__ data, const double* __restrict__ weights,
                            const double* __restrict__ pml, double* __restrict__ result, int size, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* t = pml + (long)row * size;
    double sum = 0.0;
    for (int j = lane; j < size; j += 32) sum += (data[j] * weights[j]) * t[j];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) result[row] = sum;
}

__global__ void k_naive_inv_dlt(const double* __restrict__ coeffs, const double* __restrict__ pml,
                                double* __restrict__ result, int size, int rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= size) return;
    double acc = 0.0;
    for (int i = 0; i < rows; ++i) acc += coeffs[i] * pml[(long)i * size + j];
    result[j] = acc;
}

cudaError_t launch_naive_dlt(const double* data, const double* weights, const double* pml, double* result, int size,
                             int rows, cudaStream_t st) {
    k_naive_dlt<<<(rows + 7) / 8, 256, 0, st>>>(data, weights, pml, result, size, rows);
    return cudaGetLastError();
}

cudaError_t launch_naive_inv_dlt(const double* coeffs, const double* pml, double* result, int size, int rows,
                                 cudaStream_t st) {
    k_naive_inv_dlt<<<(size + 127) / 128, 128, 0, st>>>(coeffs, pml, result, size, rows);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ peaks
__global__ void k_peak_dfma(double* out, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}

__global__ void k_peak_dmma(double* out, int iters) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void k_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) dst[i] = src[i];
}

cudaError_t measure_fp64(double* fma_tflops, double* dmma_tflops) {
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    int sms = prop.multiProcessorCount;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 20000, threads = 512, blocks = sms * 4;
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(t0);
        k_peak_dfma<<<blocks, threads>>>(d, iters);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
    }
    *fma_tflops = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(t0);
        k_peak_dmma<<<blocks, threads>>>(d, iters);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
    }
    // one m8n8k4 = 8*8*4 FMAs per warp
    *dmma_tflops = 2.0 * 256.0 * 8.0 * iters * (double)(threads / 32) * blocks / (ms * 1e-3) / 1e12;
    e = cudaGetLastError();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    return e;
}

cudaError_t measure_copy(size_t bytes, double* gbs) {
    double2 *a = nullptr, *b = nullptr;
    size_t n = bytes / sizeof(double2);
    cudaError_t e = cudaMalloc(&a, n * sizeof(double2));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&b, n * sizeof(double2));
    if (e != cudaSuccess) {
        cudaFree(a);
        return e;
    }
    cudaMemset(a, 1, n * sizeof(double2));
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    float best = 1e30f, ms = 0.f;
    for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(t0);
        k_copy<<<148 * 16, 512>>>(a, b, n);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    *gbs = 2.0 * n * sizeof(double2) / (best * 1e-3) / 1e9;
    e = cudaGetLastError();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(a);
    cudaFree(b);
    return e;
}

}  // namespace s2k
