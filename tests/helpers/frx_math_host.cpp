// Test helper: the product's FP64 scalar math (fractulus_b200/csrc/frx_math.cuh) compiled for the
// HOST, so the CPU suite can check it against mpmath without a GPU.  Every operation in that header
// is an explicitly rounded IEEE op, so these are the same bits the sm_100a build produces.
#include "../../fractulus_b200/csrc/frx_math.cuh"
extern "C" {
double t_pow_cr(double x, double y) { return frx_pow_cr(x, y); }
double t_exp_cr(double x) { return frx_exp_cr(x); }
double t_gamma_py(double x) { return frx_gamma_py(x); }
double t_pow_from_log(double x, double y) { return frx_pow_from_log(x, y, x > 1 ? frx_log_dd(x) : frx_dd{0.0, 0.0}); }
void t_log_dd(double x, double* hi, double* lo) { frx_dd L = frx_log_dd(x); *hi = L.hi; *lo = L.lo; }
void t_pow_cr_many(const double* x, const double* y, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = frx_pow_cr(x[i], y[i]);
}
}

// rule functions (fractulus_b200/csrc/frx_rules.cuh) on the host: same bits as the kernels
#include "../../fractulus_b200/csrc/frx_rules.cuh"
extern "C" int t_left_weights(int rule, double alpha, double lf, double res, int* first, int* count, double* w) {
    frx_par P = frx_par_init(rule, alpha);
    long long f, c, p = (long long)res;
    frx_left_layout(rule, p, &f, &c);
    double mult = frx_multiplier(P, lf, res);
    for (long long k = 0; k < c; ++k) w[k] = FRX_MUL(mult, frx_raw_at_offset(P, p, f + k));
    *first = (int)f;
    *count = (int)c;
    return 0;
}

// Toeplitz tables exactly as kernel K1 computes them (fractulus_b200/csrc/frx_weights.cu), on the host
extern "C" int t_tables(int rule, double alpha, double h, long N, double* cA, double* cB, double* w0, double* mult,
                        double* cm1) {
    frx_par P = frx_par_init(rule, alpha);
    for (long x = 0; x < N; ++x) {
        cA[x] = frx_raw_toeplitz(P, (double)x, true);
        if (rule == FRX_R_SIMPSON) cB[x] = frx_raw_toeplitz(P, (double)x, false);
    }
    w0[0] = mult[0] = 0.0;
    for (long x = 1; x <= N; ++x) {
        double xd = (double)x, lf = FRX_MUL(xd, h);
        mult[x] = frx_multiplier(P, lf, xd);
        w0[x] = (rule != FRX_R_RECTANGLE && !(rule == FRX_R_SIMPSON && x < 2)) ? frx_raw_first(P, xd) : 0.0;
    }
    *cm1 = rule == FRX_R_SIMPSON ? frx_raw_right_of_point(P) : 0.0;
    return 0;
}
