// frx_rules.cuh -- the four quadrature rules of reference fractulus/equation.py, stated per integer
// DISTANCE d (grid steps to the left of the evaluation node) instead of per node number, with every
// floating-point operation written out in the reference's operand order (FRX_* = explicitly rounded,
// never contracted).  pow is the correctly rounded frx_pow_cr, gamma is CPython's Lanczos
// (frx_gamma_py); see frx_math.cuh for why.
#pragma once
#include "frx_math.cuh"

enum { FRX_R_CAPUTO = 0, FRX_R_RECTANGLE = 1, FRX_R_TRAPEZOIDAL = 2, FRX_R_SIMPSON = 3, FRX_R_GL = 4 };

struct frx_par {
    int rule;
    double alpha;
    double n;        // floor(alpha) + 1.                      equation.py:22,44
    double e_main;   // caputo: n - alpha + 1. (:45); trapezoidal: 2 - alpha (:90,95); rectangle: 1. - alpha (:68)
    double e_first;  // caputo: n - alpha (:49); trapezoidal: 1 - alpha (:90)
    double e3, e2, e1;   // simpson: 3 - alpha, 2 - alpha, 1 - alpha (:114-138)
    double g4, g3, g2;   // simpson: gamma(4 - alpha), gamma(3 - alpha), gamma(2 - alpha)
    double e_mult;   // exponent of the step in the multiplier
    double g_mult;   // gamma in the multiplier's denominator (1 for simpson)
    double expo[3];  // the distinct exponents of the rule's powers: caputo/trapezoidal {e_main, e_first},
                     // rectangle {e_main}, simpson {e3, e2, e1}; indexed by the FRX_E_* slots below
};

enum { FRX_E_MAIN = 0, FRX_E_FIRST = 1, FRX_E_3 = 0, FRX_E_2 = 1, FRX_E_1 = 2 };

// "pow provider": pw(base, slot) == frx_pow_cr(base, P.expo[slot]).  The direct provider calls pow; the
// table kernel substitutes one that reads powers shared through shared memory (same bits, 4x fewer pows).
struct frx_pow_direct {
    const frx_par* P;
    FRX_HD_MEMBER double operator()(double base, int slot) const { return frx_pow_cr(base, P->expo[slot]); }
};

// exponents only (a handful of flops); the gammas are left at 1 / 0 -- see frx_par_gamma_arg
FRX_HD frx_par frx_par_exponents(int rule, double alpha) {
    frx_par P;
    P.rule = rule;
    P.alpha = alpha;
    P.n = FRX_ADD(floor(alpha), 1.0);
    P.e_main = P.e_first = P.e3 = P.e2 = P.e1 = P.g4 = P.g3 = P.g2 = 0.0;
    P.g_mult = 1.0;
    if (rule == FRX_R_CAPUTO) {
        double nma = FRX_SUB(P.n, alpha);
        P.e_main = FRX_ADD(nma, 1.0);
        P.e_first = nma;
        P.e_mult = nma;                                    // h ** (n - alpha)               :57
    } else if (rule == FRX_R_RECTANGLE) {
        P.e_main = FRX_SUB(1.0, alpha);
        P.e_mult = FRX_SUB(1.0, alpha);                    // ** (1. - alpha) / gamma(2. - alpha)  :72
    } else if (rule == FRX_R_TRAPEZOIDAL) {
        P.e_main = FRX_SUB(2.0, alpha);
        P.e_first = FRX_SUB(1.0, alpha);
        P.e_mult = FRX_SUB(1.0, alpha);                    // ** (1. - alpha) / gamma(3. - alpha)  :99
    } else if (rule == FRX_R_GL) {
        P.e_mult = -alpha;                                 // extension: h ** (-alpha) * w_k
    } else {
        P.e3 = FRX_SUB(3.0, alpha);
        P.e2 = FRX_SUB(2.0, alpha);
        P.e1 = FRX_SUB(1.0, alpha);
        P.e_mult = FRX_SUB(1.0, alpha);                    // dt ** (1. - alpha)             :172
    }
    if (rule == FRX_R_SIMPSON) {
        P.expo[0] = P.e3; P.expo[1] = P.e2; P.expo[2] = P.e1;
    } else {
        P.expo[0] = P.e_main; P.expo[1] = P.e_first; P.expo[2] = 0.0;
    }
    return P;
}

// the rule needs frx_par_gamma_count gamma values; value k is gamma(frx_par_gamma_arg(P, k)) and is
// stored by frx_par_set_gamma.  Split like this so that a warp can evaluate them on different lanes.
FRX_HD int frx_par_gamma_count(int rule) { return rule == FRX_R_SIMPSON ? 3 : (rule == FRX_R_GL ? 0 : 1); }
FRX_HD double frx_par_gamma_arg(const frx_par& P, int k) {
    if (P.rule == FRX_R_CAPUTO) return FRX_ADD(FRX_SUB(P.n, P.alpha), 2.0);   // gamma(n - alpha + 2.)  :57
    if (P.rule == FRX_R_RECTANGLE) return FRX_SUB(2.0, P.alpha);               // gamma(2. - alpha)      :72
    if (P.rule == FRX_R_TRAPEZOIDAL) return FRX_SUB(3.0, P.alpha);             // gamma(3. - alpha)      :99
    return FRX_SUB(4.0 - (double)k, P.alpha);                                  // gamma(4|3|2 - alpha)   :114-139
}
FRX_HD void frx_par_set_gamma(frx_par& P, int k, double g) {
    if (P.rule != FRX_R_SIMPSON) P.g_mult = g;
    else if (k == 0) P.g4 = g;
    else if (k == 1) P.g3 = g;
    else P.g2 = g;
}

FRX_HD frx_par frx_par_init(int rule, double alpha) {
    frx_par P = frx_par_exponents(rule, alpha);
    for (int k = 0; k < frx_par_gamma_count(rule); ++k) frx_par_set_gamma(P, k, frx_gamma_py(frx_par_gamma_arg(P, k)));
    return P;
}

// raw weight at distance d >= 0 (not the first column).  col_even: parity of the node index
// j = i - d, which selects between Simpson's two interleaved sequences; ignored by the other rules.
template <class PW>
FRX_HD double frx_raw_toeplitz_pw(const frx_par& P, double d, bool col_even, const PW& pw) {
    if (P.rule == FRX_R_CAPUTO || P.rule == FRX_R_TRAPEZOIDAL) {
        if (d == 0.0) return 1.0;                          // number == p                    :51,92
        double a = pw(FRX_ADD(d, 1.0), FRX_E_MAIN);
        double b = pw(d, FRX_E_MAIN);
        double c = pw(FRX_SUB(d, 1.0), FRX_E_MAIN);
        return FRX_ADD(FRX_SUB(a, FRX_MUL(2.0, b)), c);    // a - 2.*b + c                   :54,95
    }
    if (P.rule == FRX_R_RECTANGLE) {                       // (-k)**(1.-alpha) - (-k-1)**(1.-alpha), -k = d+1   :68
        return FRX_SUB(pw(FRX_ADD(d, 1.0), FRX_E_MAIN), pw(d, FRX_E_MAIN));
    }
    if (col_even) {
        if (d == 0.0) {                                    // even m, k == i                 :130-133
            double x = FRX_DIV(frx_pow_cr(2.0, P.e3), P.g4);
            double y = FRX_DIV(-frx_pow_cr(2.0, P.e2), FRX_MUL(2.0, P.g3));
            return FRX_ADD(x, y);
        }
        if (d == 1.0) {                                    // odd m, k == i-1 (:135-139) + u_k(-1) (:157-158)
            double x = FRX_DIV(FRX_SUB(frx_pow_cr(3.0, P.e3), 1.0), P.g4);
            double y = FRX_DIV(-FRX_ADD(frx_pow_cr(3.0, P.e2), 3.0), FRX_MUL(2.0, P.g3));
            double z = FRX_DIV(-1.0, P.g2);
            double w = FRX_ADD(FRX_ADD(x, y), z);
            double oma = FRX_SUB(1.0, P.alpha);
            double t = FRX_SUB(FRX_ADD(FRX_MUL(2.0, frx_pow_cr(oma, 2.0)), 3.0), FRX_MUL(3.0, P.alpha));
            double u = FRX_DIV(t, FRX_MUL(2.0, P.g4));
            return FRX_ADD(w, u);
        }
        double x = FRX_DIV(FRX_SUB(pw(FRX_ADD(d, 2.0), FRX_E_3), pw(FRX_SUB(d, 2.0), FRX_E_3)), P.g4);
        double s = FRX_ADD(FRX_ADD(pw(FRX_ADD(d, 2.0), FRX_E_2), FRX_MUL(6.0, pw(d, FRX_E_2))),
                           pw(FRX_SUB(d, 2.0), FRX_E_2));
        double y = FRX_DIV(s, FRX_MUL(2.0, P.g3));
        return FRX_SUB(x, y);                              // even node                      :124-128
    }
    if (d == 0.0) {                                        // odd m, k == 0: 0. + u_k(0)     :159-160,166-170
        return FRX_ADD(0.0, FRX_DIV(FRX_SUB(4.0, FRX_MUL(2.0, P.alpha)), P.g4));
    }
    double x = FRX_DIV(FRX_MUL(-2.0, FRX_SUB(pw(FRX_ADD(d, 1.0), FRX_E_3), pw(FRX_SUB(d, 1.0), FRX_E_3))), P.g4);
    double y = FRX_DIV(FRX_MUL(2.0, FRX_ADD(pw(FRX_ADD(d, 1.0), FRX_E_2), pw(FRX_SUB(d, 1.0), FRX_E_2))), P.g3);
    return FRX_ADD(x, y);                                  // odd node                       :119-122
}
FRX_HD double frx_raw_toeplitz(const frx_par& P, double d, bool col_even) {
    return frx_raw_toeplitz_pw(P, d, col_even, frx_pow_direct{&P});
}

// raw weight of node number 0 (distance d == p) -- the row-dependent first column
template <class PW>
FRX_HD double frx_raw_first_pw(const frx_par& P, double p, const PW& pw) {
    if (P.rule == FRX_R_CAPUTO) {                          // (p-1.)**index - (p-n+alpha-1.)*p**(n-alpha)   :49
        double a = pw(FRX_SUB(p, 1.0), FRX_E_MAIN);
        double f = FRX_SUB(FRX_ADD(FRX_SUB(p, P.n), P.alpha), 1.0);
        return FRX_SUB(a, FRX_MUL(f, pw(p, FRX_E_FIRST)));
    }
    if (P.rule == FRX_R_TRAPEZOIDAL) {                     // (p-1.)**(2-alpha) + (2.-alpha-p)*p**(1-alpha) :90
        double a = pw(FRX_SUB(p, 1.0), FRX_E_MAIN);
        double f = FRX_SUB(FRX_SUB(2.0, P.alpha), p);
        return FRX_ADD(a, FRX_MUL(f, pw(p, FRX_E_FIRST)));
    }
    if (P.rule == FRX_R_SIMPSON) {                         // :113-117
        double pm2 = FRX_SUB(p, 2.0);
        double x = FRX_DIV(FRX_SUB(pw(p, FRX_E_3), pw(pm2, FRX_E_3)), P.g4);
        double y = FRX_DIV(-FRX_ADD(FRX_MUL(3.0, pw(p, FRX_E_2)), pw(pm2, FRX_E_2)), FRX_MUL(2.0, P.g3));
        double z = FRX_DIV(pw(p, FRX_E_1), P.g2);
        return FRX_ADD(FRX_ADD(x, y), z);
    }
    return 0.0;
}
FRX_HD double frx_raw_first(const frx_par& P, double p) { return frx_raw_first_pw(P, p, frx_pow_direct{&P}); }

// Simpson, odd m, the node one step to the RIGHT of the evaluation point: 0. + u_k(1)   :161-162,166-170
FRX_HD double frx_raw_right_of_point(const frx_par& P) {
    return FRX_ADD(0.0, FRX_DIV(FRX_ADD(-1.0, P.alpha), FRX_MUL(2.0, P.g4)));
}

// stencil multiplier for Settings(alpha, lf, p)    :56-57, :72, :99, :105,172
FRX_HD double frx_multiplier(const frx_par& P, double lf, double p) {
    double step = FRX_DIV(lf, p);
    double m = frx_pow_cr(step, P.e_mult);
    return (P.rule == FRX_R_SIMPSON || P.rule == FRX_R_GL) ? m : FRX_DIV(m, P.g_mult);
}

// raw weight of the left stencil of resolution p at integer offset `off` (distance d = -off)
FRX_HD double frx_raw_at_offset(const frx_par& P, long long p, long long off) {
    long long d = -off;
    if (d == -1) return frx_raw_right_of_point(P);
    if (d == p && P.rule != FRX_R_RECTANGLE) return frx_raw_first(P, (double)p);
    return frx_raw_toeplitz(P, (double)d, ((p - d) & 1) == 0);
}

// ---- integer layout (shared by host validation and kernels) ------------------------------------
// left stencil: consecutive offsets [first, first + count)
FRX_HD void frx_left_layout(int rule, long long p, long long* first, long long* count) {
    if (rule == FRX_R_RECTANGLE) {            // res nodes i-(res-1), i = 0..res-1        :70-71
        *first = -(p - 1);
        *count = p;
    } else if (rule == FRX_R_SIMPSON && (p & 1)) {  // odd m: m+2 nodes, one at +1       :176
        *first = -p;
        *count = p + 2;
    } else {                                  // p+1 nodes j-p, j = 0..p                  :59,98,174
        *first = -p;
        *count = p + 1;
    }
}

// ---- Gruenwald-Letnikov extension: w_k = prod_{d=1..k} (1 - (alpha+1)/d), accumulated in double-double ----------
// factor d as a double-double: (alpha+1)/d with the exact remainder, then 1 - q
FRX_HD frx_dd frx_gl_factor(double alpha, double d) {
    const double ap1 = FRX_ADD(alpha, 1.0);                 // rounding of alpha+1 is part of the definition used here
    const double q = FRX_DIV(ap1, d);
    const double r = FRX_FMA(-q, d, ap1);                   // exact: ap1 - q*d
    const double ql = FRX_DIV(r, d);
    frx_dd one_minus = frx_two_sum(1.0, -q);
    one_minus.lo = FRX_SUB(one_minus.lo, ql);
    return frx_fast_two_sum(one_minus.hi, one_minus.lo);
}
