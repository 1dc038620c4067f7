// Do the FP64 ALU pipe (DFMA) and the FP64 tensor path (DMMA) run concurrently on B200?  Three kernels with the same
// loop count: DMMA only, DFMA only, both interleaved in every warp.  If the pipes are independent the mixed kernel
// takes max(t_dmma, t_dfma); if they share the FP64 units it takes the sum.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 1 = DMMA, 2 = DFMA, 3 = both
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double c[16], x[16];
    for (int i = 0; i < 16; ++i) { c[i] = 1e-3 * (threadIdx.x + i); x[i] = 1.0 + 1e-3 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE & 1) {
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[i]), "+d"(c[i + 1]) : "d"(a), "d"(b));
            }
            if (MODE & 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
            }
        }
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += c[i] + x[i];
    if (s == 123.456) out[0] = s;
}
template <int MODE> float run(double* out, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-7); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    return best;
}
int main() {
    double* out; cudaMalloc(&out, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096, blocks = sms * 8;
    float t1 = run<1>(out, blocks, iters), t2 = run<2>(out, blocks, iters), t3 = run<3>(out, blocks, iters);
    double fl_mma = (double)blocks * 8 * iters * 4 * 8 * 512, fl_fma = (double)blocks * 256 * iters * 4 * 16 * 2;
    printf("DMMA only %.3f ms (%.1f TF)  DFMA only %.3f ms (%.1f TF)  both %.3f ms (%.1f TF combined)  sum %.3f  max %.3f\n",
           t1, fl_mma / t1 / 1e9, t2, fl_fma / t2 / 1e9, t3, (fl_mma + fl_fma) / t3 / 1e9, t1 + t2, t1 > t2 ? t1 : t2);
    return 0;
}
