// frx_table.cuh -- kernel K1: Toeplitz weight tables of the growing-history operator, one CTA of 128 threads per
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
// fl(fl(i*h)/i) can take) come from a one-warp-per-alpha kernel (K0) that runs ahead; the table kernel is launched
// as its programmatic dependent and only waits for it after its own powers are done.
#pragma once
#include "frx_rules.cuh"

// Storage slot of distance d in a weight table.  c_perm = 0: natural order.  c_perm >= 1 (DMMA history kernel): every
// aligned group of 8 distances is stored as [0 4 1 5 2 6 3 7].  c_perm = 2 (bulk-staged geometry): additionally the
// 8-distance rows R and R^1 are swapped where bit 2 of R is set, so that a window copied densely into shared memory
// (64-byte rows, no padding) is read without bank conflicts by the A-fragment LDS.128 pattern (rows 4 apart per lane
// group).  c_pad is a multiple of 64, so the pattern is the same in table and window coordinates.
__host__ __device__ inline long long frx_table_slot(int c_pad, int c_perm, long long d) {
    if (!c_perm) return c_pad + d;
    long long R = (c_pad + d) >> 3;
    if (c_perm == 2) R ^= (R >> 2) & 1;
    return R * 8 + ((d & 3) * 2 + ((d >> 2) & 1));
}

constexpr int FRX_TAB_THREADS = 128;              // threads per CTA: one power base / table entry each
constexpr int FRX_TAB_BLOCK = FRX_TAB_THREADS;
constexpr int FRX_TAB_HALO = 2;
constexpr int FRX_TAB_ENTRIES = FRX_TAB_THREADS - 2 * FRX_TAB_HALO;  // entries per CTA

// per-alpha scalars, produced once per alpha by frx_alpha_scalars_warp (kernel K0) and read by every table CTA
struct frx_alpha_scalars {
    frx_par P;                      // exponents + Lanczos gammas
    double step[3], mult[3];        // the steps fl(fl(i*h)/i) can take (h-, h, h+) and step ** e_mult for each
    double krc, sgn;                // Riesz-Caputo: 1./2.*gamma(2.-alpha)/gamma(2.) and (-1.)**n  (equation.py:22-24)
};

struct frx_table_smem {
    frx_alpha_scalars sc;
    double pw[3][FRX_TAB_THREADS];  // pw[slot][t] = (x0 + t) ** expo[slot]
};

// K0, one CTA of 4 warps per alpha.  The pieces are independent chains of ~200-400 dependent FP64 operations, so
// each kind runs on its own warp (lanes of one warp on different code paths would serialise):
//   warp 0: Lanczos rational sums of the <= 3 gammas      warp 1: exp(y) of each gamma
//   warp 2: y ** (x - 0.5) of each gamma                  warp 3: step ** e_mult for the 3 candidate steps
// then warp 0 combines them exactly as frx_gamma_py does (same operations, same bits).
struct frx_k0_smem {
    double sum[4], ex[4], pw[4];    // slot 3: gamma(2. - alpha) of the Riesz-Caputo constant
};
__device__ __forceinline__ void frx_alpha_scalars_block(frx_k0_smem& W, int rule, double alpha, double h,
                                                        frx_alpha_scalars* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const frx_par E = frx_par_exponents(rule, alpha);
    const int ng = frx_par_gamma_count(rule);
    const double gmh = 5.524680040776729583740234375;
    if (lane < 4) {
        if (warp == 3) {
            if (lane < 3) {
                const double s = lane == 1 ? h : nextafter(h, lane == 0 ? 0.0 : 2.0 * h + 1.0);
                out->step[lane] = s;
                out->mult[lane] = frx_pow_cr(s, E.e_mult);
            }
        } else if (lane < ng || lane == 3) {
            const double x = lane == 3 ? FRX_SUB(2.0, alpha) : frx_par_gamma_arg(E, lane), y = FRX_ADD(x, gmh);
            if (warp == 0) W.sum[lane] = frx_lanczos_sum(x);
            else if (warp == 1) W.ex[lane] = frx_exp_cr(y);
            else W.pw[lane] = frx_pow_cr(y, FRX_SUB(x, 0.5));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        frx_par P = E;
        for (int k = 0; k < ng; ++k)
            frx_par_set_gamma(P, k, frx_gamma_py_combine(frx_par_gamma_arg(E, k), W.sum[k], W.ex[k], W.pw[k]));
        out->P = P;
        // Number(1. / 2. * math.gamma(2. - alpha) / math.gamma(2.)), Number((-1.) ** n)        equation.py:23-24
        const double g2 = frx_gamma_py_combine(FRX_SUB(2.0, alpha), W.sum[3], W.ex[3], W.pw[3]);
        out->krc = FRX_DIV(FRX_MUL(0.5, g2), 1.0);
        out->sgn = (((long long)E.n) & 1) ? -1.0 : 1.0;
    }
}

__device__ __forceinline__ void frx_grid_dependency_wait() {   // no-op unless launched as a programmatic dependent
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void frx_grid_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct frx_pow_shared {
    const frx_table_smem* S;
    long long x0;
    __device__ __forceinline__ double operator()(double base, int slot) const {
        const long long t = (long long)base - x0;
        if (t >= 0 && t < FRX_TAB_THREADS) return S->pw[slot][t];
        return frx_pow_cr(base, S->sc.P.expo[slot]);  // never taken for in-range entries; keeps the provider total
    }
};

// X0: storage index of the CTA's first entry; logical distance / row x = X - c_pad.  `sc` (this alpha's scalars)
// is written by the preceding kernel K0: it is read only after the grid-dependency wait, so under programmatic
// dependent launch the power evaluation overlaps K0.
__device__ __forceinline__ void frx_table_block(frx_table_smem& S, const frx_alpha_scalars* __restrict__ sc, int rule,
                                                double alpha, double h, int64_t N, int64_t X0, int c_pad, int c_perm, int dec,
                                                double* __restrict__ cA, double* __restrict__ cB, int64_t ldc,
                                                double* __restrict__ w0, double* __restrict__ mult,
                                                double* __restrict__ cm1) {
    const int t = threadIdx.x;
    const frx_par E = frx_par_exponents(rule, alpha);  // cheap: every thread has the exponents at once
    const long long x0 = X0 - c_pad - FRX_TAB_HALO;    // base of thread 0
    {
        const long long xs = x0 + t;
        if (xs >= 0) {
            const double xd = (double)xs;
            frx_dd L{0.0, 0.0};
            if (xs > 1) L = frx_log_dd(xd);
            const int ns = rule == FRX_R_SIMPSON ? 3 : (rule == FRX_R_RECTANGLE ? 1 : 2);
            if (xs > 1) {
                // the <= 3 powers of one base are independent chains of ~160 dependent FP64 ops: evaluate them
                // side by side (inlined) so that they overlap
                double r[3];
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const double y = E.expo[s];
                    r[s] = (s < ns && y != 0.0 && y != 1.0) ? frx_pow_core_inline(L, y) : (y == 0.0 ? 1.0 : xd);
                }
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    if (s < ns) S.pw[s][t] = r[s];
            } else {
                for (int s = 0; s < ns; ++s) S.pw[s][t] = frx_pow_from_log(xd, E.expo[s], L);   // bases 0 and 1: exact cases
            }
        }
    }
    frx_grid_dependency_wait();
    if (t < (int)(sizeof(frx_alpha_scalars) / sizeof(double)))
        reinterpret_cast<double*>(&S.sc)[t] = reinterpret_cast<const double*>(sc)[t];
    __syncthreads();
    if (t < FRX_TAB_HALO || t >= FRX_TAB_THREADS - FRX_TAB_HALO) return;  // halo threads and the scalar warps
    const int64_t X = X0 + (t - FRX_TAB_HALO);
    // dec (Simpson inside frx_history): the two base sequences are scattered into four half-length tables, see below;
    // the caller has zeroed them, entries 0 <= x <= N are all this CTA writes
    if (dec ? X > c_pad + N : X >= ldc) return;
    const int64_t x = X - c_pad;
    if (x < 0) {
        if (!dec) {
            cA[X] = 0.0;
            if (cB) cB[X] = 0.0;
        }
        return;
    }
    const frx_par& P = S.sc.P;
    const frx_pow_shared pw{&S, x0};
    const double xd = (double)x;
    double ca = 0.0, cb = 0.0;
    if (x < N) {
        ca = frx_raw_toeplitz_pw(P, xd, true, pw);
        if (rule == FRX_R_SIMPSON) cb = frx_raw_toeplitz_pw(P, xd, false, pw);
    }
    // the DMMA history kernel wants every aligned group of 8 distances stored as [0 4 1 5 2 6 3 7]
    auto slot_of = [&](int64_t d) { return (int64_t)frx_table_slot(c_pad, c_perm, d); };
    if (!dec) {
        cA[slot_of(x)] = ca;
        if (cB) cB[slot_of(x)] = cb;
    } else if (x < N) {
        // Simpson as four half-size Toeplitz problems (shifted coordinates i' = i-1 = 2r + q, j' = j-1 = 2k + e):
        // the coefficient of (i', j') is cE[d] for odd j' (even node j) and cO[d] for even j', d = i' - j' = 2(r-k) + q - e,
        // so table T(q,e)[r-k] takes every second entry of the base sequence:  T(q,e) = cA + (2q+e)*ldc
        //   cE (e=1): d even -> T(1,1)[d/2],   d odd -> T(0,1)[(d+1)/2]  (T(0,1)[0] = 0: column k = r lies right of row r)
        //   cO (e=0): d even -> T(0,0)[d/2],   d odd -> T(1,0)[(d-1)/2]
        if ((x & 1) == 0) {
            cA[3 * ldc + slot_of(x >> 1)] = ca;
            cA[0 * ldc + slot_of(x >> 1)] = cb;
        } else {
            cA[1 * ldc + slot_of((x + 1) >> 1)] = ca;
            cA[2 * ldc + slot_of((x - 1) >> 1)] = cb;
        }
    }
    if (x <= N) {
        double w = 0.0, m = 0.0;
        if (x >= 1) {
            const double step = FRX_DIV(FRX_MUL(xd, h), xd);  // (lf = i * h) / i at the reference's call site
            double pm;
            if (step == S.sc.step[1]) pm = S.sc.mult[1];
            else if (step == S.sc.step[0]) pm = S.sc.mult[0];
            else if (step == S.sc.step[2]) pm = S.sc.mult[2];
            else pm = frx_pow_cr(step, P.e_mult);
            m = rule == FRX_R_SIMPSON ? pm : FRX_DIV(pm, P.g_mult);
            if (rule != FRX_R_RECTANGLE && !(rule == FRX_R_SIMPSON && x < 2)) w = frx_raw_first_pw(P, xd, pw);
        }
        w0[x] = w;
        mult[x] = m;
    }
    if (x == 0 && cm1) *cm1 = rule == FRX_R_SIMPSON ? frx_raw_right_of_point(P) : 0.0;
}
