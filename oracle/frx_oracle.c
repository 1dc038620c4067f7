/* TEST INFRASTRUCTURE -- plain-C CPU restatement of the reference's fractional-operator path.
 *
 * Follows reference fractulus/equation.py expression for expression (operand order, int/float
 * mixing), calling the SAME glibc pow() that CPython's float ** uses, and a restatement of
 * CPython's math.gamma (Modules/mathmodule.c m_tgamma, Lanczos N=13 g=6.0246800407767296) since
 * math.gamma is NOT glibc tgamma.  Compile with -ffp-contract=off (see oracle/Makefile).
 *
 * Pinned by (tests/test_oracle_golden.py): bit-exact agreement with tests/golden/ fixtures, which
 * oracle/gen_golden.py produced by running the UNMODIFIED reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It is the checker, never the product.
 *
 * rule codes: 0 'caputo' (equation.py:42-59), 1 'rectangle' (:62-73), 2 'trapezoidal' (:84-100),
 *             3 'simpson' (:103-176).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* -DORC_CR_POW: the same restatement with the correctly rounded pow / exp of the product's scalar math
 * (fractulus_b200/csrc/frx_math.cuh, host build in oracle/cr_math.cpp) instead of glibc's.  That variant
 * (oracle/_build/libfrx_oracle_cr.so) isolates summation order from pow rounding: the GPU must agree with it to
 * 1e-12 on every row (tests/test_gpu_history.py), while the glibc build stays the restatement of the reference. */
#ifdef ORC_CR_POW
double orc_cr_pow(double x, double y);
double orc_cr_exp(double x);
#define pow orc_cr_pow
#define exp orc_cr_exp
#endif

#define ORC_OK 0
#define ORC_ERR_RULE -1
#define ORC_ERR_DOMAIN -2
#define ORC_ERR_ZERODIV -3

/* ---------------------------------------------------------------- CPython math.gamma ------ */
static const double lz_num[13] = {
    23531376880.410759688572007674451636754734846804940,
    42919803642.649098768957899047001988850926355848959,
    35711959237.355668049440185451547166705960488635843,
    17921034426.037209699919755754458931112671403265390,
    6039542586.3520280050642916443072979210699388420708,
    1439720407.3117216736632230727949123939715485786772,
    248874557.86205415651146038641322942321632125127801,
    31426415.585400194380614231628318205362874684987640,
    2876370.6289353724412254090516208496135991145378768,
    186056.26539522349504029498971604569928220784236328,
    8071.6720023658162106380029022722506138218516325024,
    210.82427775157934587250973392071336271166969580291,
    2.5066282746310002701649081771338373386264310793408};
static const double lz_den[13] = {0.0, 39916800.0, 120543840.0, 150917976.0, 105258076.0, 45995730.0,
                                  13339535.0, 2637558.0, 357423.0, 32670.0, 1925.0, 66.0, 1.0};
static const double lz_g = 6.024680040776729583740234375;
static const double lz_g_minus_half = 5.524680040776729583740234375;

double orc_tgamma(double x) {
    /* positive finite arguments only (the path never leaves (0, 6]) */
    if (x == floor(x) && x >= 1.0 && x <= 23.0) {
        double f = 1.0;
        for (int k = 2; k < (int)x; ++k) f *= k;
        return f;
    }
    double num = 0.0, den = 0.0;
    if (x < 5.0) {
        for (int i = 12; i >= 0; --i) { num = num * x + lz_num[i]; den = den * x + lz_den[i]; }
    } else {
        for (int i = 0; i < 13; ++i) { num = num / x + lz_num[i]; den = den / x + lz_den[i]; }
    }
    double y = x + lz_g_minus_half, z;
    if (x > lz_g_minus_half) { double q = y - x; z = q - lz_g_minus_half; }
    else { double q = y - lz_g_minus_half; z = q - x; }
    z = z * lz_g / y;
    double r = (num / den) / exp(y);
    r += z * r;
    r *= pow(y, x - 0.5);
    return r;
}

/* ---------------------------------------------------------------- per-rule raw weights ---- */
typedef struct {
    int rule;
    double alpha;
    double n, index;               /* caputo */
    double g2, g3, g4;             /* simpson gammas */
} orc_par;

static void par_init(orc_par* P, int rule, double alpha) {
    P->rule = rule; P->alpha = alpha;
    P->n = floor(alpha) + 1.;
    P->index = (P->n - alpha + 1.);
    if (rule == 3) { P->g4 = orc_tgamma(4 - alpha); P->g3 = orc_tgamma(3 - alpha); P->g2 = orc_tgamma(2 - alpha); }
}

/* Toeplitz part: raw weight at distance d (>= 0) left of the evaluation node; for Simpson the
 * weight also depends on the parity of the column j = i - d (col_even). */
static double raw_toeplitz(const orc_par* P, long d, int col_even) {
    double a = P->alpha, D = (double)d;
    switch (P->rule) {
    case 0:
        if (d == 0) return 1.;
        return pow(D + 1., P->index) - 2. * pow(D, P->index) + pow(D - 1., P->index);
    case 1: /* k = -(d+1): (-k)**(1.-alpha) - (-k-1)**(1.-alpha) */
        return pow(D + 1, 1. - a) - pow(D, 1. - a);
    case 2:
        if (d == 0) return 1.;
        return pow(D + 1., 2 - a) - 2 * pow(D, 2 - a) + pow(D - 1., 2 - a);
    default: {
        if (col_even) {
            if (d == 0) { /* even m, k == i: equation.py:130-133 */
                double x = pow(2, 3 - a) / P->g4, y = -pow(2, 2 - a) / (2 * P->g3);
                return x + y;
            }
            if (d == 1) { /* odd m, k == i-1: equation.py:135-139  +  u_k(-1): :157-158 */
                double x = (pow(3, 3 - a) - 1) / P->g4, y = -(pow(3, 2 - a) + 3) / (2 * P->g3), z = -1. / P->g2;
                double w = x + y + z;
                double u = (2. * pow(1. - a, 2.0) + 3. - 3. * a) / (2 * P->g4);
                return w + u;
            }
            double x = (pow(D + 2, 3 - a) - pow(D - 2, 3 - a)) / P->g4;
            double y = (pow(D + 2, 2 - a) + 6 * pow(D, 2 - a) + pow(D - 2, 2 - a)) / (2 * P->g3);
            return x - y;            /* equation.py:124-128 */
        }
        if (d == 0) return 0. + (4. - 2. * a) / P->g4;  /* odd m, k == 0: w = 0., u_k(0): :159-160 */
        double x = -2 * (pow(D + 1, 3 - a) - pow(D - 1, 3 - a)) / P->g4;
        double y = 2 * (pow(D + 1, 2 - a) + pow(D - 1, 2 - a)) / P->g3;
        return x + y;                /* equation.py:119-122 */
    }
    }
}

/* first column (node number 0, distance d == p) */
static double raw_first(const orc_par* P, long p_) {
    double a = P->alpha, p = (double)p_;
    switch (P->rule) {
    case 0: return pow(p - 1., P->index) - (p - P->n + a - 1.) * pow(p, P->n - a);
    case 2: return pow(p - 1., 2 - a) + (2. - a - p) * pow(p, 1 - a);
    case 3: {
        double x = (pow(p, 3 - a) - pow(p - 2, 3 - a)) / P->g4;
        double y = -(3 * pow(p, 2 - a) + pow(p - 2, 2 - a)) / (2 * P->g3);
        double z = pow(p, 1 - a) / P->g2;
        return x + y + z;            /* equation.py:113-117 */
    }
    default: return 0.;
    }
}

static double raw_right_of_point(const orc_par* P) { /* simpson odd m, k == +1: u_k(1), :161-162 */
    return 0. + (-1. + P->alpha) / (2. * P->g4);
}

static double multiplier(const orc_par* P, double lf, double p) {
    double a = P->alpha;
    switch (P->rule) {
    case 0: { double h = (0. - (-lf)) / p; return pow(h, P->n - a) / orc_tgamma(P->n - a + 2.); }
    case 1: return pow(lf / p, 1. - a) / orc_tgamma(2. - a);
    case 2: return pow(lf / p, 1. - a) / orc_tgamma(3. - a);
    default: return pow(lf / p, 1. - a);
    }
}

static int check_domain(int rule, double alpha, double res) {
    if (rule < 0 || rule > 3) return ORC_ERR_RULE;
    if (res == 0.) return ORC_ERR_ZERODIV;
    if (res < 0. || res != floor(res) || !(alpha >= 0.)) return ORC_ERR_DOMAIN;
    if (rule == 1 && alpha > 1.) return ORC_ERR_ZERODIV;   /* 0.0 ** negative, equation.py:68 */
    if (rule == 2 && alpha > 2.) return ORC_ERR_ZERODIV;   /* 0.0 ** negative, equation.py:95 */
    if (rule == 3 && alpha >= 2.) return ORC_ERR_DOMAIN;   /* gamma(2 - alpha) pole / complex weights */
    if (rule == 1 && res < 2.) return ORC_ERR_ZERODIV;  /* zero-interval uniform stencil */
    if (rule == 3 && res < 2.) return ORC_ERR_DOMAIN;   /* (m-2)**x with m=1 is complex in Python */
    return ORC_OK;
}

/* integer layout: consecutive offsets [first, first+count) -- equation.py:59,70-71,98,173-176 */
int orc_left_layout(int rule, double res, int* first, int* count) {
    long p = (long)res;
    if (rule == 1) { *first = (int)-(p - 1); *count = (int)p; }
    else if (rule == 3 && (p & 1)) { *first = (int)-p; *count = (int)p + 2; }
    else { *first = (int)-p; *count = (int)p + 1; }
    return ORC_OK;
}

/* left stencil of Settings(alpha, lf, res): w[k] is the weight at offset first+k */
int orc_left_weights(int rule, double alpha, double lf, double res, int* first, int* count, double* w) {
    int rc = check_domain(rule, alpha, res);
    if (rc) return rc;
    orc_par P; par_init(&P, rule, alpha);
    orc_left_layout(rule, res, first, count);
    long p = (long)res;
    double mult = multiplier(&P, lf, res);
    for (int k = 0; k < *count; ++k) {
        long off = (long)*first + k, d = -off;
        double raw;
        if (d == -1) raw = raw_right_of_point(&P);
        else if (d == p && rule != 1) raw = raw_first(&P, p);
        else raw = raw_toeplitz(&P, d, ((p - d) & 1) == 0);
        w[k] = mult * raw;
    }
    return ORC_OK;
}

/* Toeplitz tables in the layout the GPU weight-table kernel produces (include/frx.h):
 * cA[d], cB[d] (cB: Simpson odd-column sequence, else unused), d = 0..N-1; w0[i], mult[i], i = 0..N
 * (entry 0 unused = 0), row i <-> Settings(alpha, fl(i*h), i).  *cm1 = Simpson's right-of-point raw weight. */
int orc_tables(int rule, double alpha, double h, long N, double* cA, double* cB, double* w0, double* mult, double* cm1) {
    if (rule < 0 || rule > 3) return ORC_ERR_RULE;
    orc_par P; par_init(&P, rule, alpha);
#pragma omp parallel for schedule(static)
    for (long d = 0; d < N; ++d) {
        cA[d] = raw_toeplitz(&P, d, 1);
        if (rule == 3) cB[d] = raw_toeplitz(&P, d, 0);
    }
    w0[0] = 0.; mult[0] = 0.;
#pragma omp parallel for schedule(static)
    for (long i = 1; i <= N; ++i) {
        double lf = (double)i * h;
        w0[i] = (rule == 1 || (rule == 3 && i < 2)) ? 0. : raw_first(&P, i);
        mult[i] = multiplier(&P, lf, (double)i);
    }
    *cm1 = rule == 3 ? raw_right_of_point(&P) : 0.;
    return ORC_OK;
}

/* CPython >= 3.12 builtin sum() of floats: Neumaier compensated summation (Python/bltinmodule.c) */
typedef struct { double total, comp; } nsum;
static inline void nsum_add(nsum* s, double x) {
    double t = s->total + x;
    if (fabs(s->total) >= fabs(x)) s->comp += (s->total - t) + x;
    else s->comp += (x - t) + s->total;
    s->total = t;
}
static inline double nsum_get(const nsum* s) { return s->total + s->comp; }

/* Growing-history application, the reference's own algorithm (test/integration/test_equation.py:51-57):
 * for every requested row i the stencil of Settings(alpha, i*h, i) is REBUILT FROM SCRATCH
 * (O(i) pow calls) and y_i = sum(g[node] * weight).  g has n_g entries (node 0..n_g-1). */
static int history_rows_naive_impl(int rule, double alpha, double h, const double* g, long n_g,
                                   const long* rows, long n_rows, double* y, int n_threads, int abs_terms) {
    if (rule < 0 || rule > 3) return ORC_ERR_RULE;
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long r = 0; r < n_rows; ++r) {
        long i = rows[r];
        int first, count;
        if (check_domain(rule, alpha, (double)i)) { err = 1; y[r] = NAN; continue; }
        orc_left_layout(rule, (double)i, &first, &count);
        if (i + first + count - 1 >= n_g) { err = 1; y[r] = NAN; continue; }
        double* w = (double*)malloc(sizeof(double) * (size_t)count);
        orc_left_weights(rule, alpha, (double)i * h, (double)i, &first, &count, w);
        nsum s = {0., 0.};
        for (int k = 0; k < count; ++k) nsum_add(&s, abs_terms ? fabs(g[i + first + k] * w[k]) : g[i + first + k] * w[k]);
        y[r] = nsum_get(&s);
        free(w);
    }
    return err ? ORC_ERR_DOMAIN : ORC_OK;
}
int orc_history_rows_naive(int rule, double alpha, double h, const double* g, long n_g,
                           const long* rows, long n_rows, double* y, int n_threads) {
    return history_rows_naive_impl(rule, alpha, h, g, n_g, rows, n_rows, y, n_threads, 0);
}
/* S_i = sum_j |g_j * w_ij|: the scale of the rounding error of ANY uncompensated FP64 summation of row i (tests use
 * it for the condition-aware part of the parity tolerance at zero crossings of y) */
int orc_history_rows_naive_abs(int rule, double alpha, double h, const double* g, long n_g,
                               const long* rows, long n_rows, double* y, int n_threads) {
    return history_rows_naive_impl(rule, alpha, h, g, n_g, rows, n_rows, y, n_threads, 1);
}

/* Same values, table-accelerated (raw weights depend on the distance only, so one O(N) table
 * serves every row; identical floating-point operations per term => bit-identical results,
 * asserted against orc_history_rows_naive in tests/test_oracle_golden.py).  Rows 1..N -> y[1..N]
 * (y[0] = 0; rows outside the rule's domain = NaN). n_fields fields of stride ld_g / ld_y. */
#define ORC_T(x) (abs_terms ? fabs(x) : (x))
static int history_full_impl(int rule, double alpha, double h, long N, const double* g, long n_g, long n_fields,
                             long ld_g, double* y, long ld_y, int n_threads, int abs_terms) {
    if (rule < 0 || rule > 3) return ORC_ERR_RULE;
    if (n_g < N + 1 + (rule == 3)) return ORC_ERR_DOMAIN;
    double* cA = (double*)malloc(sizeof(double) * (size_t)(N + 1) * 4);
    double *cB = cA + (N + 1), *w0 = cB + (N + 1), *mult = w0 + (N + 1), cm1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    orc_tables(rule, alpha, h, N, cA, cB, w0, mult, &cm1);
    for (long f = 0; f < n_fields; ++f) {
        const double* gf = g + f * ld_g;
        double* yf = y + f * ld_y;
        yf[0] = 0.;
#pragma omp parallel for schedule(dynamic, 16)
        for (long i = N; i >= 1; --i) {
            if (check_domain(rule, alpha, (double)i)) { yf[i] = NAN; continue; }
            double m = mult[i];
            nsum s = {0., 0.};
            if (rule != 1) nsum_add(&s, ORC_T(gf[0] * (m * w0[i])));             /* node number 0 */
            if (rule == 3) {
                for (long d = i - 1; d >= 0; --d) {
                    long j = i - d;
                    nsum_add(&s, ORC_T(gf[j] * (m * ((j & 1) ? cB[d] : cA[d]))));
                }
                if (i & 1) nsum_add(&s, ORC_T(gf[i + 1] * (m * cm1)));
            } else {
                for (long d = i - 1; d >= 0; --d) nsum_add(&s, ORC_T(gf[i - d] * (m * cA[d])));
            }
            yf[i] = nsum_get(&s);
        }
    }
    free(cA);
    return ORC_OK;
}
int orc_history_full(int rule, double alpha, double h, long N, const double* g, long n_g, long n_fields,
                     long ld_g, double* y, long ld_y, int n_threads) {
    return history_full_impl(rule, alpha, h, N, g, n_g, n_fields, ld_g, y, ld_y, n_threads, 0);
}
int orc_history_full_abs(int rule, double alpha, double h, long N, const double* g, long n_g, long n_fields,
                         long ld_g, double* y, long ld_y, int n_threads) {
    return history_full_impl(rule, alpha, h, N, g, n_g, n_fields, ld_g, y, ld_y, n_threads, 1);
}

/* Inverse of the growing-history operator by forward substitution with the reference's own weights
 * w_ij = mult_i * raw (every weight rounded like the reference's): given f_1..f_N and u_0 find u_1..u_N with
 * sum_j w_ij u_j = f_i.  The reference has no solver; this is the definition the GPU extension
 * frx_history_solve is checked against.  'caputo' (0) and 'trapezoidal' (2). */
int orc_history_solve(int rule, double alpha, double h, long N, const double* f, double u0, double* u) {
    if (rule != 0 && rule != 2) return ORC_ERR_RULE;
    double* cA = (double*)malloc(sizeof(double) * (size_t)(N + 1) * 4);
    double *cB = cA + (N + 1), *w0 = cB + (N + 1), *mult = w0 + (N + 1), cm1;
    orc_tables(rule, alpha, h, N, cA, cB, w0, mult, &cm1);
    u[0] = u0;
    for (long i = 1; i <= N; ++i) {
        const double m = mult[i];
        nsum s = {0., 0.};
        nsum_add(&s, u[0] * (m * w0[i]));
        for (long d = i - 1; d >= 1; --d) nsum_add(&s, u[i - d] * (m * cA[d]));
        u[i] = (f[i] - nsum_get(&s)) / (m * cA[0]);
    }
    free(cA);
    return ORC_OK;
}

/* this build's pow (glibc's, or the correctly rounded one with -DORC_CR_POW): lets tests locate the arguments on
 * which the two differ */
double orc_pow(double x, double y) { return pow(x, y); }
void orc_pow_many(const double* x, const double* y, double* out, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) out[i] = pow(x[i], y[i]);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
