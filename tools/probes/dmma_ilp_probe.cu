// DMMA (mma.sync.m8n8k4.f64) throughput as a function of independent accumulator chains per warp and resident warps per
// SM: tells the latency of a dependent DMMA and how many warps / chains it takes to saturate the FP64 tensor pipe.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_ilp_probe dmma_ilp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void probe(double* out, int iters, double a, double b) {
    double c[2 * CH];
#pragma unroll
    for (int k = 0; k < 2 * CH; ++k) c[k] = 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32 / CH; ++u)
#pragma unroll
            for (int k = 0; k < CH; ++k)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[2 * k]), "+d"(c[2 * k + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 2 * CH; ++k) s += c[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(double* d, int warps_per_sm, int sms) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<CH><<<sms, 32 * warps_per_sm>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double dmma_per_warp = (double)iters * 32;
    const double tf = dmma_per_warp * warps_per_sm * sms * 512.0 / (best * 1e-3) / 1e12;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per DMMA per SM sub-partition (4 per SM) at the nominal clock
    const double cyc = best * 1e-3 * clk * 1e3 / (dmma_per_warp * (warps_per_sm / 4.0 < 1 ? 1 : warps_per_sm / 4.0));
    printf("chains %2d warps/SM %2d : %7.2f TFLOP/s   %6.1f cycles per DMMA per sub-partition\n", CH, warps_per_sm, tf, cyc);
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, sizeof(double) * 1024 * 1024 * 4);
    for (int w : {4, 8, 16, 32}) {
        run<1>(d, w, sms); run<2>(d, w, sms); run<4>(d, w, sms); run<8>(d, w, sms); run<16>(d, w, sms);
    }
    return 0;
}
