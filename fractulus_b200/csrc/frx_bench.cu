// frx_bench.cu -- register-only FP64 peak probes (DFMA pipe and DMMA tensor pipe).  MEASURED_PEAKS.json
// has no FP64 figure, so bench.py measures the denominators of the FP64 rooflines in the same run.
#include "frx_internal.h"

namespace {

__global__ void __launch_bounds__(256) frx_probe_dfma(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) frx_probe_dmma(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[k]), "+d"(c[k + 1])
                             : "d"(a), "d"(b));
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += c[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// best-of-`repeats` sustained FP64 throughput in TFLOP/s of (a) the DFMA pipe, (b) mma.sync m8n8k4 f64
extern "C" int frx_measure_fp64_peak(frx_ctx* ctx, int repeats, double* tflops_dfma, double* tflops_dmma) {
    FRX_REQUIRE(ctx && tflops_dfma && tflops_dmma && repeats > 0, FRX_ERR_INVALID_ARG, "bad argument");
    FRX_ON_DEVICE(ctx);
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    int rc = frx_ws_reserve(ctx, sizeof(double) * (size_t)blocks * threads);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    FRX_CUDA(cudaEventCreate(&e0));
    FRX_CUDA(cudaEventCreate(&e1));
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        for (int r = 0; r < repeats + 1; ++r) {
            FRX_CUDA(cudaEventRecord(e0, ctx->stream));
            if (which == 0)
                frx_probe_dfma<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->ws, iters, 0.999999, 1e-7);
            else
                frx_probe_dmma<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->ws, iters, 0.999999, 1e-7);
            ctx->launches += 1;
            FRX_CUDA(cudaGetLastError());
            FRX_CUDA(cudaEventRecord(e1, ctx->stream));
            FRX_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            FRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            double flop = which == 0 ? (double)blocks * threads * iters * 4.0 * 16.0 * 2.0
                                     : (double)blocks * (threads / 32) * iters * 4.0 * 8.0 * 512.0;
            double tf = flop / (ms * 1e-3) / 1e12;
            if (r > 0 && tf > best[which]) best[which] = tf;  // r == 0 is the warm-up
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_dfma = best[0];
    *tflops_dmma = best[1];
    return FRX_OK;
}
