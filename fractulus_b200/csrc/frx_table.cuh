// frx_table.cuh -- kernel K1: Toeplitz weight tables of the growing-history operator, one CTA (128 + 64 threads) per
// 124 consecutive table entries of one alpha.
//
// Entry x serves as a distance d (cA/cB: raw weights, reference equation.py:54,68,95,119-139) and as a row
// number i (w0: first-column raw weight, equation.py:49,90,113-117; mult: multiplier of
// Settings(alpha, fl(i*h), i), equation.py:56-57,72,99,105,172).
//
// Every raw weight is a combination of powers (x+k)**e, k in -2..2, with at most three distinct exponents e, so
// the CTA evaluates log(x) once per base (double-double), the <= 3 powers per base once, and shares them through
// shared memory: ~4 double-double exp() per entry instead of ~12, with the same bits as calling frx_pow_cr for
// every power (frx_pow_from_log).  The per-alpha scalars (Lanczos gammas, the multiplier of the three step values
// fl(fl(i*h)/i) can take) are evaluated on separate lanes of warp 0 while the other threads do their powers.
#pragma once
#include "frx_rules.cuh"

constexpr int FRX_TAB_THREADS = 128;              // threads that evaluate powers / entries
constexpr int FRX_TAB_BLOCK = FRX_TAB_THREADS + 64; // + warp 4: gammas, warp 5: multiplier candidates
constexpr int FRX_TAB_HALO = 2;
constexpr int FRX_TAB_ENTRIES = FRX_TAB_THREADS - 2 * FRX_TAB_HALO;  // entries per CTA

struct frx_table_smem {
    frx_par P;
    double pw[3][FRX_TAB_THREADS];  // pw[slot][t] = (x0 + t) ** expo[slot]
    double step[3], mult[3];        // candidate steps h-, h, h+ and their final multipliers
};

struct frx_pow_shared {
    const frx_table_smem* S;
    long long x0;
    __device__ __forceinline__ double operator()(double base, int slot) const {
        const long long t = (long long)base - x0;
        if (t >= 0 && t < FRX_TAB_THREADS) return S->pw[slot][t];
        return frx_pow_cr(base, S->P.expo[slot]);  // never taken for in-range entries; keeps the provider total
    }
};

// X0: storage index of the CTA's first entry; logical distance / row x = X - c_pad.
__device__ __forceinline__ void frx_table_block(frx_table_smem& S, int rule, double alpha, double h, int64_t N, int64_t X0,
                                                int c_pad, int c_perm, double* __restrict__ cA, double* __restrict__ cB,
                                                int64_t ldc, double* __restrict__ w0, double* __restrict__ mult,
                                                double* __restrict__ cm1) {
    const int t = threadIdx.x;
    const frx_par E = frx_par_exponents(rule, alpha);  // cheap: every thread has the exponents at once
    const long long x0 = X0 - c_pad - FRX_TAB_HALO;    // base of thread 0
    if (t == 0) S.P = E;
    __syncthreads();
    if (t >= FRX_TAB_THREADS) {
        // the per-alpha scalars run on their own warps so that they overlap the power evaluation below
        const int k = t & 31;
        if (t < FRX_TAB_THREADS + 32) {
            if (k < frx_par_gamma_count(rule)) frx_par_set_gamma(S.P, k, frx_gamma_py(frx_par_gamma_arg(E, k)));
        } else if (k < 3) {
            const double s = k == 1 ? h : nextafter(h, k == 0 ? 0.0 : 2.0 * h + 1.0);
            S.step[k] = s;
            S.mult[k] = frx_pow_cr(s, E.e_mult);  // divided by the gamma after the barrier
        }
    } else {
        const long long xs = x0 + t;
        if (xs >= 0) {
            const double xd = (double)xs;
            frx_dd L{0.0, 0.0};
            if (xs > 1) L = frx_log_dd(xd);
            const int ns = rule == FRX_R_SIMPSON ? 3 : (rule == FRX_R_RECTANGLE ? 1 : 2);
            for (int s = 0; s < ns; ++s) S.pw[s][t] = frx_pow_from_log(xd, E.expo[s], L);
        }
    }
    __syncthreads();
    if (t < FRX_TAB_HALO || t >= FRX_TAB_THREADS - FRX_TAB_HALO) return;  // halo threads and the scalar warps
    const int64_t X = X0 + (t - FRX_TAB_HALO);
    if (X >= ldc) return;
    const int64_t x = X - c_pad;
    if (x < 0) {
        cA[X] = 0.0;
        if (cB) cB[X] = 0.0;
        return;
    }
    const frx_par& P = S.P;
    const frx_pow_shared pw{&S, x0};
    const double xd = (double)x;
    double ca = 0.0, cb = 0.0;
    if (x < N) {
        ca = frx_raw_toeplitz_pw(P, xd, true, pw);
        if (rule == FRX_R_SIMPSON) cb = frx_raw_toeplitz_pw(P, xd, false, pw);
    }
    // the DMMA history kernel wants every aligned group of 8 distances stored as [0 4 1 5 2 6 3 7]
    const int64_t slot = c_perm ? c_pad + (x & ~(int64_t)7) + ((x & 3) * 2 + ((x >> 2) & 1)) : X;
    cA[slot] = ca;
    if (cB) cB[slot] = cb;
    if (x <= N) {
        double w = 0.0, m = 0.0;
        if (x >= 1) {
            const double step = FRX_DIV(FRX_MUL(xd, h), xd);  // (lf = i * h) / i at the reference's call site
            double pm;
            if (step == S.step[1]) pm = S.mult[1];
            else if (step == S.step[0]) pm = S.mult[0];
            else if (step == S.step[2]) pm = S.mult[2];
            else pm = frx_pow_cr(step, P.e_mult);
            m = rule == FRX_R_SIMPSON ? pm : FRX_DIV(pm, P.g_mult);
            if (rule != FRX_R_RECTANGLE && !(rule == FRX_R_SIMPSON && x < 2)) w = frx_raw_first_pw(P, xd, pw);
        }
        w0[x] = w;
        mult[x] = m;
    }
    if (x == 0 && cm1) *cm1 = rule == FRX_R_SIMPSON ? frx_raw_right_of_point(P) : 0.0;
}
