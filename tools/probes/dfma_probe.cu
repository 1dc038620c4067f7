// Stand-alone DFMA/DMMA issue-rate probes (not part of the library): which operand patterns reach the FP64 peak?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

// P0: one varying operand (accumulator as multiplicand), a and b constant
__global__ void __launch_bounds__(256) p0(double* out, int iters, double a, double b) {
    double x[16];
    for (int k = 0; k < 16; ++k) x[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    double s = 0; for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}
// P1: like the history kernel: acc[r] += c[s] * w[(r - s) & 15]: c reused 8x, w and acc vary
template <int NACC>
__global__ void __launch_bounds__(256) p1(double* out, int iters, double a) {
    double acc[NACC], w[16], c[8];
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int k = 0; k < 16; ++k) w[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int k = 0; k < 8; ++k) c[k] = a + 1e-3 * k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int r = 0; r < NACC; ++r) acc[r] = fma(c[s], w[(r - s + 16) & 15], acc[r]);
        // keep the compiler from hoisting: perturb one window value per iteration (1 extra op per 8*NACC)
        w[it & 15] += 1e-9;
    }
    double s = 0; for (int k = 0; k < NACC; ++k) s += acc[k];
    if (s == 123.456) out[0] = s;
}
// P2: DMMA m8n8k4, 8 independent accumulator pairs
__global__ void __launch_bounds__(256) p2(double* out, int iters, double a, double b) {
    double c[16];
    for (int k = 0; k < 16; ++k) c[k] = 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 16; k += 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[k]), "+d"(c[k + 1]) : "d"(a), "d"(b));
    double s = 0; for (int k = 0; k < 16; ++k) s += c[k];
    if (s == 123.456) out[0] = s;
}
template <class F>
double timeit(F f, double flop) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = flop / (ms * 1e-3) / 1e12; if (r && tf > best) best = tf;
    }
    return best;
}
int main() {
    double* out; CK(cudaMalloc(&out, 1024));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 4096;
    for (int bps : {1, 2, 4, 8}) {      // blocks of 256 threads per SM -> 2, 4, 8, 16 warps per SMSP
        int blocks = sms * bps;
        double n = (double)blocks * 256 * iters;
        printf("blocks/SM %d (warps/SMSP %d):", bps, bps * 2);
        printf("  P0 %.2f", timeit([&] { p0<<<blocks, 256>>>(out, iters, 0.999999, 1e-7); }, n * 64 * 2));
        printf("  P1<8> %.2f", timeit([&] { p1<8><<<blocks, 256>>>(out, iters, 0.5); }, n * 64 * 2));
        printf("  P1<16> %.2f", timeit([&] { p1<16><<<blocks, 256>>>(out, iters, 0.5); }, n * 128 * 2));
        printf("  DMMA %.2f TFLOP/s\n", timeit([&] { p2<<<blocks, 256>>>(out, iters, 0.999999, 1e-7); }, (double)blocks * 8 * iters * 4 * 8 * 512));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
