// frx_math.cuh -- FP64 scalar math for the weight generators, written so that every operation is
// an explicitly rounded IEEE-754 op (no compiler FMA contraction): the same source compiled for
// the host (g++ -ffp-contract=off, tests/_build helper) and for sm_100a gives the same bits,
// except through the log() seed of frx_log_dd (libm vs libdevice), which moves results by < 2^-100.
//
// Why this exists: the reference computes its weights with glibc pow() (CPython float **) and
// CPython's Lanczos math.gamma (reference fractulus/equation.py:45-57,68,72,90-99,105-172).  Weights
// such as (d+1)^s - 2 d^s + (d-1)^s cancel catastrophically, so any two pow() implementations that
// differ in the last ulp give visibly different weights.  The only implementation-independent
// target is the CORRECTLY ROUNDED pow, which glibc's (< 1 ULP, not CR) matches on all but a small
// fraction of inputs.  frx_pow_cr evaluates x^y in double-double (~2^-100) and rounds once.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FRX_HD __host__ __device__ __forceinline__
// the big routines (exp/pow/gamma) are real calls on the device: the rule functions use up to a dozen
// pow() each and inlining them all costs minutes of compile time for no measurable speed
#define FRX_HD_CALL static __host__ __device__ __noinline__
#define FRX_HD_MEMBER __host__ __device__ __forceinline__
#else
#define FRX_HD static inline
#define FRX_HD_CALL static __attribute__((noinline))
#define FRX_HD_MEMBER inline
#endif

#if defined(__CUDA_ARCH__)
#define FRX_ADD(a, b) __dadd_rn((a), (b))
#define FRX_SUB(a, b) __dsub_rn((a), (b))
#define FRX_MUL(a, b) __dmul_rn((a), (b))
#define FRX_DIV(a, b) __ddiv_rn((a), (b))
#define FRX_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
// host build: compiled with -ffp-contract=off, so plain operators round once each
#define FRX_ADD(a, b) ((a) + (b))
#define FRX_SUB(a, b) ((a) - (b))
#define FRX_MUL(a, b) ((a) * (b))
#define FRX_DIV(a, b) ((a) / (b))
#define FRX_FMA(a, b, c) fma((a), (b), (c))
#endif

#include "frx_math_tables.inc"

#if defined(__CUDACC__)
static __device__ const double frx_exp2_tab_dev[128] = {FRX_EXP2_TABLE};
#endif
static const double frx_exp2_tab_host[128] = {FRX_EXP2_TABLE};

struct frx_dd {
    double hi, lo;
};

FRX_HD frx_dd frx_two_sum(double a, double b) {
    double s = FRX_ADD(a, b);
    double bb = FRX_SUB(s, a);
    double e = FRX_ADD(FRX_SUB(a, FRX_SUB(s, bb)), FRX_SUB(b, bb));
    return frx_dd{s, e};
}
FRX_HD frx_dd frx_fast_two_sum(double a, double b) {  // |a| >= |b| or a == 0
    double s = FRX_ADD(a, b);
    double e = FRX_SUB(b, FRX_SUB(s, a));
    return frx_dd{s, e};
}
FRX_HD frx_dd frx_two_prod(double a, double b) {
    double p = FRX_MUL(a, b);
    double e = FRX_FMA(a, b, -p);
    return frx_dd{p, e};
}
FRX_HD frx_dd frx_dd_add(frx_dd a, frx_dd b) {  // accurate (Knuth/Dekker) double-double sum
    frx_dd s = frx_two_sum(a.hi, b.hi);
    frx_dd t = frx_two_sum(a.lo, b.lo);
    s.lo = FRX_ADD(s.lo, t.hi);
    s = frx_fast_two_sum(s.hi, s.lo);
    s.lo = FRX_ADD(s.lo, t.lo);
    return frx_fast_two_sum(s.hi, s.lo);
}
FRX_HD frx_dd frx_dd_add_d(frx_dd a, double b) {
    frx_dd s = frx_two_sum(a.hi, b);
    s.lo = FRX_ADD(s.lo, a.lo);
    return frx_fast_two_sum(s.hi, s.lo);
}
FRX_HD frx_dd frx_dd_mul(frx_dd a, frx_dd b) {
    frx_dd p = frx_two_prod(a.hi, b.hi);
    p.lo = FRX_FMA(a.hi, b.lo, p.lo);
    p.lo = FRX_FMA(a.lo, b.hi, p.lo);
    return frx_fast_two_sum(p.hi, p.lo);
}
FRX_HD frx_dd frx_dd_mul_d(frx_dd a, double b) {
    frx_dd p = frx_two_prod(a.hi, b);
    p.lo = FRX_FMA(a.lo, b, p.lo);
    return frx_fast_two_sum(p.hi, p.lo);
}

// exp(zh + zl) as an (unrounded) double-double times 2^(*scale2).  |z| < 700.
// z = (64 q + j) ln2/64 + r, |r| <= ln2/128: exp(z) = 2^q * 2^(j/64) * exp(r); exp(r) is a Taylor
// series whose terms n >= 6 fit in plain double (they are < 2^-54 of the sum) and whose terms
// n <= 5 are accumulated in double-double.
FRX_HD frx_dd frx_exp_dd_inline(double zh, double zl, int* scale2) {
    double kd = rint(FRX_MUL(zh, FRX_64_OVER_LN2));
    double a = FRX_FMA(-kd, FRX_LN2_64_C1, zh);               // exact: C1 has 30 bits
    frx_dd b = frx_two_prod(kd, FRX_LN2_64_C2);               // exact
    frx_dd s = frx_two_sum(a, -b.hi);
    double t = FRX_ADD(FRX_SUB(s.lo, b.lo), FRX_FMA(-kd, FRX_LN2_64_C3, zl));
    frx_dd r = frx_fast_two_sum(s.hi, t);
    double x = r.hi;
    double q = FRX_INV_FACT13;
    q = FRX_FMA(q, x, FRX_INV_FACT12);
    q = FRX_FMA(q, x, FRX_INV_FACT11);
    q = FRX_FMA(q, x, FRX_INV_FACT10);
    q = FRX_FMA(q, x, FRX_INV_FACT9);
    q = FRX_FMA(q, x, FRX_INV_FACT8);
    q = FRX_FMA(q, x, FRX_INV_FACT7);
    q = FRX_FMA(q, x, FRX_INV_FACT6);
    q = FRX_MUL(q, x);                                        // sum_{n>=6} r^(n-5)/n!
    frx_dd acc = frx_dd_add_d(frx_dd{FRX_INV_FACT5_HI, FRX_INV_FACT5_LO}, q);
    acc = frx_dd_add(frx_dd_mul(acc, r), frx_dd{FRX_INV_FACT4_HI, FRX_INV_FACT4_LO});
    acc = frx_dd_add(frx_dd_mul(acc, r), frx_dd{FRX_INV_FACT3_HI, FRX_INV_FACT3_LO});
    acc = frx_dd_add_d(frx_dd_mul(acc, r), 0.5);
    acc = frx_dd_add_d(frx_dd_mul(acc, r), 1.0);
    acc = frx_dd_add_d(frx_dd_mul(acc, r), 1.0);
    long long k = (long long)kd;
    int j = (int)(k & 63);
    *scale2 = (int)((k - j) / 64);
#if defined(__CUDA_ARCH__)
    frx_dd tj{frx_exp2_tab_dev[2 * j], frx_exp2_tab_dev[2 * j + 1]};
#else
    frx_dd tj{frx_exp2_tab_host[2 * j], frx_exp2_tab_host[2 * j + 1]};
#endif
    return frx_dd_mul(acc, tj);
}

// out-of-line copy for the call sites where code size matters more than overlap (see FRX_HD_CALL above)
FRX_HD_CALL frx_dd frx_exp_dd(double zh, double zl, int* scale2) { return frx_exp_dd_inline(zh, zl, scale2); }

FRX_HD double frx_scale2(double v, int e) {  // v * 2^e, exact while the result stays normal
#if defined(__CUDA_ARCH__)
    return scalbn(v, e);
#else
    return ldexp(v, e);
#endif
}

// correctly rounded (with overwhelming probability: ~2^-100 internal error) exp(x)
FRX_HD double frx_exp_cr(double x) {
    int e;
    frx_dd v = frx_exp_dd(x, 0.0, &e);
    return frx_scale2(FRX_ADD(v.hi, v.lo), e);
}

// log(x) as a double-double for finite x > 0: seed y0 = log(x) to ~1 ulp from the platform's libm,
// then one Newton step on exp: log(x) = y0 + log(x * exp(-y0)) = y0 + t - t^2/2, t = x*exp(-y0) - 1
// (|t| ~ 1e-16, so the cubic term is < 2^-150).  The seed only has to be within a few ulp; its
// last-bit difference between libm and libdevice moves the result by < 2^-100.
FRX_HD_CALL frx_dd frx_log_dd(double x) {
    double y0 = log(x);
    int e;
    frx_dd E = frx_exp_dd(-y0, 0.0, &e);
    frx_dd p = frx_dd_mul_d(E, x);
    p.hi = frx_scale2(p.hi, e);
    p.lo = frx_scale2(p.lo, e);
    double th = FRX_SUB(p.hi, 1.0);                           // exact (Sterbenz)
    frx_dd t = frx_two_sum(th, p.lo);
    t.lo = FRX_FMA(FRX_MUL(-0.5, t.hi), t.hi, t.lo);
    frx_dd L = frx_two_sum(y0, t.hi);
    L.lo = FRX_ADD(L.lo, t.lo);
    return frx_fast_two_sum(L.hi, L.lo);
}

// x ** y for the domain the reference reaches: x >= 0 finite, y finite.  Python semantics at the
// edges: x ** 0 == 1, 0 ** y == 0 for y > 0, negative base with an integer-valued exponent = +-|x| ** y;
// 0 ** negative raises ZeroDivisionError and a negative base with a fractional exponent gives a complex number
// in the reference -- both are rejected on the host before launch, the device returns NaN for them.
FRX_HD_CALL double frx_pow_core(frx_dd L, double y) {  // exp(y * L) rounded once, L = log(x) as a double-double
    frx_dd z = frx_dd_mul_d(L, y);
    int e;
    frx_dd v = frx_exp_dd(z.hi, z.lo, &e);
    return frx_scale2(FRX_ADD(v.hi, v.lo), e);
}
// inlined variant of frx_pow_core: independent calls side by side overlap their dependency chains (same bits)
FRX_HD double frx_pow_core_inline(frx_dd L, double y) {
    frx_dd z = frx_dd_mul_d(L, y);
    int e;
    frx_dd v = frx_exp_dd_inline(z.hi, z.lo, &e);
    return frx_scale2(FRX_ADD(v.hi, v.lo), e);
}
// same value as frx_pow_cr(x, y) given L = frx_log_dd(x) (L is ignored for the exact special cases)
FRX_HD double frx_pow_from_log(double x, double y, frx_dd L) {
    if (y == 0.0 || x == 1.0) return 1.0;
    if (x == 0.0) return y > 0.0 ? 0.0 : NAN;
    if (!(x > 0.0)) return NAN;
    if (y == 1.0) return x;
    return frx_pow_core(L, y);
}
FRX_HD_CALL double frx_pow_cr(double x, double y) {
    if (y == 0.0 || x == 1.0) return 1.0;
    if (x == 0.0) return y > 0.0 ? 0.0 : NAN;
    if (x < 0.0 && y == floor(y)) {            // (1.-alpha)**2 with alpha > 1 (reference equation.py:158): a float result
        const double m = frx_pow_core(frx_log_dd(-x), y);
        return fmod(y, 2.0) != 0.0 ? -m : m;
    }
    if (!(x > 0.0)) return NAN;
    if (y == 1.0) return x;
    return frx_pow_core(frx_log_dd(x), y);
}

// CPython's math.gamma for positive arguments (Modules/mathmodule.c m_tgamma: Lanczos sum N = 13,
// g = 6.024680040776729583740234375), which is what the reference's math.gamma calls evaluate
// (reference fractulus/equation.py:23,57,72,99,114-139,158-162); NOT C tgamma.  exp and pow inside
// are the correctly rounded ones above where CPython uses libm's.
// the three independent pieces of CPython's m_tgamma for 0 < x: the Lanczos rational sum, exp(y) and
// y ** (x - 0.5) with y = x + g - 0.5, and the final combination.  frx_gamma_py chains them; kernel K0 evaluates
// the pieces on different warps (fractulus_b200/csrc/frx_table.cuh) -- same operations, same bits.
FRX_HD_CALL double frx_lanczos_sum(double x) {
    const double num_c[13] = {
        23531376880.410759688572007674451636754734846804940, 42919803642.649098768957899047001988850926355848959,
        35711959237.355668049440185451547166705960488635843, 17921034426.037209699919755754458931112671403265390,
        6039542586.3520280050642916443072979210699388420708, 1439720407.3117216736632230727949123939715485786772,
        248874557.86205415651146038641322942321632125127801, 31426415.585400194380614231628318205362874684987640,
        2876370.6289353724412254090516208496135991145378768, 186056.26539522349504029498971604569928220784236328,
        8071.6720023658162106380029022722506138218516325024, 210.82427775157934587250973392071336271166969580291,
        2.5066282746310002701649081771338373386264310793408};
    const double den_c[13] = {0.0,        39916800.0, 120543840.0, 150917976.0, 105258076.0, 45995730.0, 13339535.0,
                              2637558.0,  357423.0,   32670.0,     1925.0,      66.0,        1.0};
    double num = 0.0, den = 0.0;
    if (x < 5.0) {
        for (int i = 12; i >= 0; --i) {
            num = FRX_ADD(FRX_MUL(num, x), num_c[i]);
            den = FRX_ADD(FRX_MUL(den, x), den_c[i]);
        }
    } else {
        for (int i = 0; i < 13; ++i) {
            num = FRX_ADD(FRX_DIV(num, x), num_c[i]);
            den = FRX_ADD(FRX_DIV(den, x), den_c[i]);
        }
    }
    return FRX_DIV(num, den);
}
FRX_HD bool frx_gamma_is_factorial(double x) { return x == floor(x) && x >= 1.0 && x <= 23.0; }
FRX_HD double frx_gamma_factorial(double x) {
    double f = 1.0;
    for (int k = 2; k < (int)x; ++k) f = FRX_MUL(f, (double)k);
    return f;
}
// sum = frx_lanczos_sum(x), ex = exp(y), pw = y ** (x - 0.5), y = x + g - 0.5 (0 < x < 140)
FRX_HD double frx_gamma_py_combine(double x, double sum, double ex, double pw) {
    const double g = 6.024680040776729583740234375, gmh = 5.524680040776729583740234375;
    if (frx_gamma_is_factorial(x)) return frx_gamma_factorial(x);
    double y = FRX_ADD(x, gmh), z;
    if (x > gmh) {
        double q = FRX_SUB(y, x);
        z = FRX_SUB(q, gmh);
    } else {
        double q = FRX_SUB(y, gmh);
        z = FRX_SUB(q, x);
    }
    z = FRX_DIV(FRX_MUL(z, g), y);
    double r = FRX_DIV(sum, ex);
    r = FRX_ADD(r, FRX_MUL(z, r));
    r = FRX_MUL(r, pw);
    return r;
}
FRX_HD_CALL double frx_gamma_py(double x) {
    if (frx_gamma_is_factorial(x)) return frx_gamma_factorial(x);
    const double y = FRX_ADD(x, 5.524680040776729583740234375);
    return frx_gamma_py_combine(x, frx_lanczos_sum(x), frx_exp_cr(y), frx_pow_cr(y, FRX_SUB(x, 0.5)));
}
