// frx_table.cuh -- one entry of the Toeplitz weight tables of the growing-history operator (kernel K1).
// Index x serves as a distance d (cA/cB: raw weights, reference equation.py:54,68,95,119-139) and as a row
// number i (w0: first-column raw weight, equation.py:49,90,113-117; mult: stencil multiplier of
// Settings(alpha, fl(i*h), i), equation.py:56-57,72,99,105,172).
#pragma once
#include "frx_rules.cuh"

struct frx_table_entry {
    double ca, cb, w0, mult;
};

__device__ __forceinline__ frx_table_entry frx_table_at(const frx_par& P, int64_t x, int64_t N, double h) {
    frx_table_entry e{0.0, 0.0, 0.0, 0.0};
    const double xd = (double)x;
    if (x < N) {
        e.ca = frx_raw_toeplitz(P, xd, true);
        if (P.rule == FRX_R_SIMPSON) e.cb = frx_raw_toeplitz(P, xd, false);
    }
    if (x >= 1 && x <= N) {
        const double lf = FRX_MUL(xd, h);  // lf = i * h at the reference's call site
        e.mult = frx_multiplier(P, lf, xd);
        if (P.rule != FRX_R_RECTANGLE && !(P.rule == FRX_R_SIMPSON && x < 2)) e.w0 = frx_raw_first(P, xd);
    }
    return e;
}
