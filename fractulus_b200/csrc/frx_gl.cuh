// frx_gl.cuh -- EXTENSION (not in the reference): Gruenwald-Letnikov weights
//     w_0 = 1,  w_x = w_{x-1} * (1 - (alpha + 1) / x)
// The recurrence is a running product, so a CTA evaluates it as a parallel prefix product: every thread multiplies
// its own GL_E consecutive factors in registers, a shuffle/shared-memory scan combines the per-thread products, and
// a carry links successive chunks.  Products are accumulated in double-double and rounded once, so the table does
// not drift with x (a plain double recurrence loses ~x/2 ulp by x = 1e5).
#pragma once
#include "frx_rules.cuh"

constexpr int FRX_GL_E = 4;   // consecutive entries per thread and chunk

template <int NTH>
struct frx_gl_smem {
    frx_dd warp_tot[NTH / 32];
    frx_dd carry;
};

// calls store(x, w_x) for x in [0, x_end), all NTH threads of the CTA must call it
template <int NTH, class Store>
__device__ __forceinline__ void frx_gl_block(frx_gl_smem<NTH>& S, double alpha, long long x_end, const Store& store) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) S.carry = frx_dd{1.0, 0.0};
    __syncthreads();
    for (long long chunk = 0; chunk < x_end; chunk += (long long)NTH * FRX_GL_E) {
        const long long x0 = chunk + (long long)t * FRX_GL_E;
        frx_dd p[FRX_GL_E];
#pragma unroll
        for (int e = 0; e < FRX_GL_E; ++e) {
            const long long x = x0 + e;
            const frx_dd f = (x >= 1 && x < x_end) ? frx_gl_factor(alpha, (double)x) : frx_dd{1.0, 0.0};
            p[e] = e == 0 ? f : frx_dd_mul(p[e - 1], f);
        }
        // inclusive scan of the per-thread totals over the warp, then over the warps
        frx_dd inc = p[FRX_GL_E - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double oh = __shfl_up_sync(0xffffffffu, inc.hi, d), ol = __shfl_up_sync(0xffffffffu, inc.lo, d);
            if (lane >= d) inc = frx_dd_mul(frx_dd{oh, ol}, inc);
        }
        if (lane == 31) S.warp_tot[warp] = inc;
        __syncthreads();
        frx_dd pre = S.carry;                                   // product of everything before this warp
        for (int w = 0; w < warp; ++w) pre = frx_dd_mul(pre, S.warp_tot[w]);
        // exclusive prefix of this thread = pre * (inclusive of previous lane)
        const double eh = __shfl_up_sync(0xffffffffu, inc.hi, 1), el = __shfl_up_sync(0xffffffffu, inc.lo, 1);
        if (lane > 0) pre = frx_dd_mul(pre, frx_dd{eh, el});
#pragma unroll
        for (int e = 0; e < FRX_GL_E; ++e) {
            const long long x = x0 + e;
            if (x < x_end) {
                const frx_dd w = frx_dd_mul(pre, p[e]);
                store(x, FRX_ADD(w.hi, w.lo));
            }
        }
        __syncthreads();
        if (t == NTH - 1) S.carry = frx_dd_mul(pre, p[FRX_GL_E - 1]);
        __syncthreads();
    }
}
