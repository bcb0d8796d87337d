// dmma_probe.cu -- how many warps per SM / independent accumulators does DMMA.8x8x4 need to reach its peak on sm_100a?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/dmma_probe tools/probes/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC>
__global__ void k(double* out, int iters, double a0, double b0) {
    double acc[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int ACC>
void run(int warps_per_sm, int sms, double* d) {
    int threads = warps_per_sm >= 4 ? 32 * (warps_per_sm / 4 > 32 ? 32 : warps_per_sm) : 32 * warps_per_sm;
    int blocks_per_sm = 1;
    if (warps_per_sm > 32) { threads = 1024; blocks_per_sm = warps_per_sm / 32; }
    else threads = 32 * warps_per_sm;
    const int iters = 20000;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(t0);
        k<ACC><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0, 2.0);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
    }
    double tf = 2.0 * 256.0 * ACC * iters * (double)(threads / 32) * sms * blocks_per_sm / (ms * 1e-3) / 1e12;
    printf("warps/SM %3d  acc %2d  %.2f TFLOP/s   (%.1f clk per DMMA per warp at 1.965 GHz)\n", warps_per_sm, ACC, tf,
           ms * 1e-3 * 1.965e9 / ((double)ACC * iters));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    double* d;
    cudaMalloc(&d, 64);
    for (int w : {4, 8, 12, 16, 24, 32, 64}) {
        run<1>(w, sms, d);
        run<2>(w, sms, d);
        run<4>(w, sms, d);
        run<8>(w, sms, d);
        run<16>(w, sms, d);
    }
    return 0;
}
