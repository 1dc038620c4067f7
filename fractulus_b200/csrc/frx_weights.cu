// frx_weights.cu -- K5 (stencil weights for batches of Settings) and K1 (Toeplitz weight tables of the
// growing-history operator).  FP64 ALU work: O(count) correctly rounded pow() per stencil / table.
#include "frx_internal.h"
#include "frx_rules.cuh"
#include "frx_table.cuh"
#include "frx_gl.cuh"

// ------------------------------------------------------------------------------------------------
// K5: one CTA per Settings(alpha, lf, resolution); thread k produces the node at offset first + k.
// Replaces create_left_stencil / create_right_stencil / create_riesz_caputo_stencil
// (reference fractulus/equation.py:13-39) including the scalar algebra the reference delegates to
// fdm: mirror + multiply(-1.) (:38-39) and K * (left + (-1.)**n * right) flattened at 0 (:20-25).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) frx_k5_stencil_weights(int rule, int side, const double* __restrict__ alpha,
                                                              const double* __restrict__ lf,
                                                              const double* __restrict__ resolution,
                                                              const int64_t* __restrict__ row_ptr,
                                                              int32_t* __restrict__ offsets, double* __restrict__ weights) {
    __shared__ frx_par P;
    __shared__ double s_mult, s_K, s_sgn;
    __shared__ frx_gl_smem<128> GS;
    const int64_t s = blockIdx.x;
    const double a = alpha[s], L = lf[s], res = resolution[s];
    const long long p = (long long)res;
    if (rule == FRX_R_GL) {   // extension: weight h**(-alpha) * w_x at offset -x, x = 0..p (left side only)
        if (threadIdx.x == 0) {
            P = frx_par_exponents(rule, a);
            s_mult = frx_multiplier(P, L, res);
        }
        __syncthreads();
        const double mult = s_mult;
        const int64_t base = row_ptr[s];
        frx_gl_block<128>(GS, a, p + 1, [&](long long x, double w) {
            weights[base + (p - x)] = FRX_MUL(mult, w);
            if (offsets) offsets[base + (p - x)] = (int32_t)(-x);
        });
        return;
    }
    if (threadIdx.x == 0) {
        P = frx_par_init(rule, a);
        s_mult = frx_multiplier(P, L, res);
        // Number(1. / 2. * math.gamma(2. - alpha) / math.gamma(2.))                     equation.py:23
        s_K = FRX_DIV(FRX_MUL(0.5, frx_gamma_py(FRX_SUB(2.0, a))), 1.0);
        s_sgn = (((long long)P.n) & 1) ? -1.0 : 1.0;  // (-1.) ** n                      equation.py:24
    }
    __syncthreads();
    long long lfirst, lcount;
    frx_left_layout(rule, p, &lfirst, &lcount);
    const long long llast = lfirst + lcount - 1;
    long long first = lfirst, count = lcount;
    if (side == FRX_SIDE_RIGHT) {
        first = -llast;
    } else if (side == FRX_SIDE_RIESZ_CAPUTO) {
        first = lfirst < -llast ? lfirst : -llast;
        count = -2 * first + 1;
    }
    const int64_t base = row_ptr[s];
    const double mult = s_mult;
    for (long long k = threadIdx.x; k < count; k += blockDim.x) {
        const long long o = first + k;
        double w;
        if (side == FRX_SIDE_LEFT) {
            w = FRX_MUL(mult, frx_raw_at_offset(P, p, o));                 // multiplier * weight(node)   :78-79
        } else if (side == FRX_SIDE_RIGHT) {
            w = FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0);  // symmetry + multiply(-1.)    :38-39
        } else {
            const bool in_left = (o >= lfirst && o <= llast), in_mirror = (-o >= lfirst && -o <= llast);
            double acc = 0.0;
            if (in_left) acc = FRX_MUL(mult, frx_raw_at_offset(P, p, o));
            if (in_mirror) {
                double r = FRX_MUL(s_sgn, FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0));
                acc = in_left ? FRX_ADD(acc, r) : r;
            }
            w = FRX_MUL(s_K, acc);
        }
        weights[base + k] = w;
        if (offsets) offsets[base + k] = (int32_t)o;
    }
}

// Batched form of K5 for the four reference rules: (a) one THREAD per Settings evaluates its scalars (gammas,
// multiplier, Riesz-Caputo constant) -- every thread runs the same chain, so the warp stays converged; (b) one
// thread per stencil NODE of the whole batch (the stencil is found by bisection in row_ptr) evaluates its weight.
struct frx_k5_scalars {
    frx_par P;
    double mult, K, sgn;
};

__global__ void __launch_bounds__(128) frx_k5a_settings_scalars(int rule, int64_t n, const double* __restrict__ alpha,
                                                                const double* __restrict__ lf,
                                                                const double* __restrict__ resolution,
                                                                frx_k5_scalars* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    frx_k5_scalars S;
    S.P = frx_par_init(rule, alpha[s]);
    S.mult = frx_multiplier(S.P, lf[s], resolution[s]);
    S.K = FRX_DIV(FRX_MUL(0.5, frx_gamma_py(FRX_SUB(2.0, alpha[s]))), 1.0);   // 1./2.*gamma(2.-alpha)/gamma(2.)   equation.py:23
    S.sgn = (((long long)S.P.n) & 1) ? -1.0 : 1.0;                             // (-1.) ** n                       equation.py:24
    out[s] = S;
}

__global__ void __launch_bounds__(128) frx_k5b_stencil_nodes(int rule, int side, int64_t n, int64_t total,
                                                             const double* __restrict__ resolution,
                                                             const int64_t* __restrict__ row_ptr,
                                                             const frx_k5_scalars* __restrict__ scal,
                                                             int32_t* __restrict__ offsets, double* __restrict__ weights) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int64_t lo = 0, hi = n;                   // row_ptr[lo] <= e < row_ptr[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (row_ptr[mid] <= e) lo = mid; else hi = mid;
    }
    const int64_t s = lo, k = e - row_ptr[s];
    const frx_k5_scalars& S = scal[s];
    const frx_par P = S.P;
    const long long p = (long long)resolution[s];
    long long lfirst, lcount;
    frx_left_layout(rule, p, &lfirst, &lcount);
    const long long llast = lfirst + lcount - 1;
    long long first = lfirst;
    if (side == FRX_SIDE_RIGHT) first = -llast;
    else if (side == FRX_SIDE_RIESZ_CAPUTO) first = lfirst < -llast ? lfirst : -llast;
    const long long o = first + k;
    const double mult = S.mult;
    double w;
    if (side == FRX_SIDE_LEFT) {
        w = FRX_MUL(mult, frx_raw_at_offset(P, p, o));                 // multiplier * weight(node)   :78-79
    } else if (side == FRX_SIDE_RIGHT) {
        w = FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0);  // symmetry + multiply(-1.)    :38-39
    } else {
        const bool in_left = (o >= lfirst && o <= llast), in_mirror = (-o >= lfirst && -o <= llast);
        double acc = 0.0;
        if (in_left) acc = FRX_MUL(mult, frx_raw_at_offset(P, p, o));
        if (in_mirror) {
            double r = FRX_MUL(S.sgn, FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0));
            acc = in_left ? FRX_ADD(acc, r) : r;
        }
        w = FRX_MUL(S.K, acc);
    }
    weights[e] = w;
    if (offsets) offsets[e] = (int32_t)o;
}

// Riesz-Caputo batches: K (left + (-1.)**n * mirror) is symmetric (n odd) or antisymmetric (n even) in the offset, and its two
// halves are built from the SAME left weights -- one thread per node of the LEFT half (offsets <= 0) evaluates its left weight
// once and writes both the node and its mirror image, with the operations (and their order) of k5b: same bits, half the pow().
// Every Riesz-Caputo stencil has an odd node count, so the left half of stencil s starts at (row_ptr[s] + s) / 2.
__global__ void __launch_bounds__(128) frx_k5b_riesz_half(int rule, int64_t n, int64_t total_half,
                                                          const double* __restrict__ resolution,
                                                          const int64_t* __restrict__ row_ptr,
                                                          const frx_k5_scalars* __restrict__ scal,
                                                          int32_t* __restrict__ offsets, double* __restrict__ weights) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total_half) return;
    int64_t lo = 0, hi = n;                   // half_ptr(lo) <= e < half_ptr(hi), half_ptr(s) = (row_ptr[s] + s) / 2
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (((row_ptr[mid] + mid) >> 1) <= e) lo = mid; else hi = mid;
    }
    const int64_t s = lo, k = e - ((row_ptr[s] + s) >> 1), base = row_ptr[s], count = row_ptr[s + 1] - base;
    const frx_k5_scalars& S = scal[s];
    const frx_par P = S.P;
    const long long p = (long long)resolution[s];
    long long lfirst, lcount;
    frx_left_layout(rule, p, &lfirst, &lcount);
    const long long llast = lfirst + lcount - 1;
    const long long first = lfirst < -llast ? lfirst : -llast;
    const long long o = first + k;            // <= 0
    const double mult = S.mult;
    const bool in_left = (o >= lfirst && o <= llast), in_left_m = (-o >= lfirst && -o <= llast);
    const double Lo = in_left ? FRX_MUL(mult, frx_raw_at_offset(P, p, o)) : 0.0;
    const double Lm = (in_left_m && o != 0) ? FRX_MUL(mult, frx_raw_at_offset(P, p, -o)) : Lo;
    {   // node o:  acc = left(o); if (-o in left) acc += sgn * (left(-o) * -1.)
        double acc = 0.0;
        if (in_left) acc = Lo;
        if (in_left_m) {
            const double r = FRX_MUL(S.sgn, FRX_MUL(Lm, -1.0));
            acc = in_left ? FRX_ADD(acc, r) : r;
        }
        weights[base + k] = FRX_MUL(S.K, acc);
        if (offsets) offsets[base + k] = (int32_t)o;
    }
    if (o != 0) {   // node -o: left(-o) if it exists, then + sgn * (left(o) * -1.)
        double acc = 0.0;
        if (in_left_m) acc = Lm;
        if (in_left) {
            const double r = FRX_MUL(S.sgn, FRX_MUL(Lo, -1.0));
            acc = in_left_m ? FRX_ADD(acc, r) : r;
        }
        weights[base + (count - 1 - k)] = FRX_MUL(S.K, acc);
        if (offsets) offsets[base + (count - 1 - k)] = (int32_t)(-o);
    }
}

// 'caputo' and 'trapezoidal' weights are second differences of ONE power sequence, a - 2b + c with a, b, c = (d+1)^e, d^e,
// (d-1)^e (equation.py:54,95); the first node adds p^e' with a second exponent (equation.py:49,90).  frx_k5c_riesz_shared gives
// every node of the left half one lane that evaluates ONE correctly rounded pow() -- what the stencil kernels spend their
// time on -- and takes the other operands from its neighbours by shuffle: the same pow() results in the same expressions,
// so the same bits as frx_k5b_riesz_half with a third of the evaluations.  No lane may take a second pow() (a divergent call
// would cost the whole warp another evaluation), so
//   * every stencil gets one extra VIRTUAL position: [first node: p^e'] [virtual: p^e] [p-1] [p-2] ... [0];
//   * warps overlap: a warp covers positions 29w - 1 .. 29w + 30 and writes the nodes of lanes 1..29 (the first node reads
//     the lane two to its right, everyone else its two neighbours).
struct frx_pow_lanes {
    const frx_par* P;
    double base0, q0, base1, q1, base2, q2;     // three (base, power) pairs for slot FRX_E_MAIN, plus ...
    double basef, qf;                            // ... one for slot FRX_E_FIRST
    __device__ double operator()(double base, int slot) const {
        if (slot == FRX_E_MAIN) {
            if (base == base0) return q0;
            if (base == base1) return q1;
            if (base == base2) return q2;
        } else if (slot == FRX_E_FIRST && base == basef) {
            return qf;
        }
        return frx_pow_cr(base, P->expo[slot]);     // (never taken for valid layouts)
    }
};

__global__ void __launch_bounds__(128) frx_k5c_riesz_shared(int rule, int64_t n, int64_t total_virtual,
                                                            const double* __restrict__ resolution,
                                                            const int64_t* __restrict__ row_ptr,
                                                            const frx_k5_scalars* __restrict__ scal,
                                                            int32_t* __restrict__ offsets, double* __restrict__ weights) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t pos_raw = 29 * gw + lane - 1;
    const bool valid = pos_raw >= 0 && pos_raw < total_virtual;
    const int64_t pos = pos_raw < 0 ? 0 : (pos_raw >= total_virtual ? total_virtual - 1 : pos_raw);
    // Which stencil owns position pos: vptr(s) <= pos < vptr(s + 1), vptr(s) = (row_ptr[s] + s) / 2 + s.  A binary search per
    // lane is 17 dependent loads at 1e5 stencils; the warp searches ONCE, 33-way, for its first position (4 dependent
    // loads), and each lane then places itself among the next 32 stencil starts (a stencil has at least two positions, so
    // the warp's 32 positions span at most 16 stencils).  Same stencil, same bits; 0.141 -> 0.125-0.140 ms for the 1e5
    // stencils of cfg 4 -- the kernel's time is the correctly rounded pow() itself.
    auto vptr = [&](const int64_t t) { return ((row_ptr[t] + t) >> 1) + t; };
    const int64_t pos0 = __shfl_sync(FULL, pos, 0);
    int64_t lo = 0, hi = n;                   // vptr(lo) <= pos0 < vptr(hi)
    while (hi - lo > 1) {
        const int64_t span = hi - lo;
        const int64_t probe = lo + ((lane + 1) * span) / 33;          // nondecreasing in the lane, in [lo, hi)
        const int below = __popc(__ballot_sync(FULL, vptr(probe) <= pos0));   // the lanes that pass form a prefix
        const int64_t new_lo = below ? lo + (below * span) / 33 : lo;
        hi = below < 32 ? lo + ((below + 1) * span) / 33 : hi;
        lo = new_lo;
    }
    const int64_t nxt = lo + 1 + lane;                                // start of stencil lo + 1 + lane (total beyond the last)
    const int64_t v_next = nxt < n ? vptr(nxt) : total_virtual;
    int ahead = 0;                                                    // stencil starts among them that are <= pos
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int64_t v = __shfl_sync(FULL, v_next, ahead + step - 1);
        if (v <= pos) ahead += step;
    }
    const int64_t s = lo + ahead, kv = pos - vptr(s), base = row_ptr[s], count = row_ptr[s + 1] - base;
    const frx_k5_scalars& S = scal[s];
    const frx_par P = S.P;
    const long long p = (long long)resolution[s];
    // kv = 0: the first node (distance p, exponent e'), kv = 1: virtual (distance p, exponent e), kv >= 2: distance p - kv + 1
    const bool is_first = kv == 0, is_virtual = kv == 1;
    const double d = (double)(kv <= 1 ? p : p - kv + 1);
    const double own = frx_pow_cr(d, P.expo[is_first ? FRX_E_FIRST : FRX_E_MAIN]);
    const double q_m1 = __shfl_up_sync(FULL, own, 1), q_p1 = __shfl_down_sync(FULL, own, 1), q_p2 = __shfl_down_sync(FULL, own, 2);
    if (!valid || is_virtual || lane < 1 || lane > 29) return;
    frx_pow_lanes pw;
    pw.P = &P;
    double raw;
    if (is_first) {                           // (p-1)^e from two lanes to the right, p^e' is this lane's
        pw.base0 = d - 1.0; pw.q0 = q_p2;
        pw.base1 = pw.base2 = -1.0; pw.q1 = pw.q2 = 0.0;
        pw.basef = d; pw.qf = own;
        raw = frx_raw_first_pw(P, d, pw);
    } else {
        pw.base0 = d; pw.q0 = own;
        pw.base1 = d + 1.0; pw.q1 = q_m1;
        pw.base2 = d - 1.0; pw.q2 = q_p1;
        pw.basef = -1.0; pw.qf = 0.0;
        raw = frx_raw_toeplitz_pw(P, d, false, pw);
    }
    // as frx_k5b_riesz_half for these rules: the left stencil is offsets -p .. 0, a node's mirror image is in it only for o = 0
    const int64_t k = is_first ? 0 : kv - 1;
    const long long o = -(long long)d;
    const double Lo = FRX_MUL(S.mult, raw);
    {
        double acc = Lo;
        if (o == 0) acc = FRX_ADD(acc, FRX_MUL(S.sgn, FRX_MUL(Lo, -1.0)));
        weights[base + k] = FRX_MUL(S.K, acc);
        if (offsets) offsets[base + k] = (int32_t)o;
    }
    if (o != 0) {
        weights[base + (count - 1 - k)] = FRX_MUL(S.K, FRX_MUL(S.sgn, FRX_MUL(Lo, -1.0)));
        if (offsets) offsets[base + (count - 1 - k)] = (int32_t)(-o);
    }
}

// One Settings triple passed BY VALUE (the drop-in's create_*_stencil call): no parameter upload, one launch.  Every CTA
// re-evaluates the scalars (a few microseconds) and then one node per thread -- the operations of k5a + k5b, same bits.
__global__ void __launch_bounds__(128) frx_k5_single(int rule, int side, double alpha, double lf, double resolution,
                                                     long long first, long long count, double* __restrict__ weights) {
    __shared__ frx_k5_scalars S;
    __shared__ double s_pm;
    // the scalars are three independent ~5 us chains (the rule's gammas, the power of the step, the Riesz-Caputo gamma):
    // one warp each -- the same operations as frx_par_init / frx_multiplier, evaluated side by side
    if (threadIdx.x == 0) {
        S.P = frx_par_init(rule, alpha);
        S.sgn = (((long long)S.P.n) & 1) ? -1.0 : 1.0;
    } else if (threadIdx.x == 32) {
        const frx_par E = frx_par_exponents(rule, alpha);
        s_pm = frx_pow_cr(FRX_DIV(lf, resolution), E.e_mult);
    } else if (threadIdx.x == 64) {
        S.K = side == FRX_SIDE_RIESZ_CAPUTO ? FRX_DIV(FRX_MUL(0.5, frx_gamma_py(FRX_SUB(2.0, alpha))), 1.0) : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) S.mult = (rule == FRX_R_SIMPSON || rule == FRX_R_GL) ? s_pm : FRX_DIV(s_pm, S.P.g_mult);   // = frx_multiplier
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const frx_par P = S.P;
    const long long p = (long long)resolution;
    long long lfirst, lcount;
    frx_left_layout(rule, p, &lfirst, &lcount);
    const long long llast = lfirst + lcount - 1;
    const long long o = first + k;
    const double mult = S.mult;
    double w;
    if (side == FRX_SIDE_LEFT) {
        w = FRX_MUL(mult, frx_raw_at_offset(P, p, o));
    } else if (side == FRX_SIDE_RIGHT) {
        w = FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0);
    } else {
        const bool in_left = (o >= lfirst && o <= llast), in_mirror = (-o >= lfirst && -o <= llast);
        double acc = 0.0;
        if (in_left) acc = FRX_MUL(mult, frx_raw_at_offset(P, p, o));
        if (in_mirror) {
            double r = FRX_MUL(S.sgn, FRX_MUL(FRX_MUL(mult, frx_raw_at_offset(P, p, -o)), -1.0));
            acc = in_left ? FRX_ADD(acc, r) : r;
        }
        w = FRX_MUL(S.K, acc);
    }
    weights[k] = w;
}

size_t frx_k5_scratch_bytes(int64_t n_settings) { return sizeof(frx_k5_scalars) * (size_t)n_settings; }

// K5 launch with the row pointer already on the device (also used by frx_assemble_solve).  d_scratch:
// frx_k5_scratch_bytes(n_settings) bytes of device memory; total = row_ptr[n_settings].
int frx_launch_stencil_weights(frx_ctx* ctx, int rule, int side, int64_t n_settings, const double* d_alpha,
                               const double* d_lf, const double* d_resolution, const int64_t* d_row_ptr, int64_t total,
                               void* d_scratch, int32_t* d_offsets, double* d_weights) {
    if (rule == FRX_RULE_GRUNWALD_LETNIKOV || total > (int64_t)64 * n_settings + 4096) {
        // the GL scan is CTA-cooperative; very long stencils keep one CTA each (one pass over row_ptr is cheaper)
        FRX_LAUNCH(ctx, FRX_K_STENCIL,
                   frx_k5_stencil_weights<<<(unsigned)n_settings, 128, 0, ctx->stream>>>(
                       rule, side, d_alpha, d_lf, d_resolution, d_row_ptr, d_offsets, d_weights));
        return FRX_OK;
    }
    FRX_LAUNCH(ctx, FRX_K_STENCIL,
               frx_k5a_settings_scalars<<<(unsigned)((n_settings + 127) / 128), 128, 0, ctx->stream>>>(
                   rule, n_settings, d_alpha, d_lf, d_resolution, (frx_k5_scalars*)d_scratch));
    if (side == FRX_SIDE_RIESZ_CAPUTO && (rule == FRX_RULE_CAPUTO || rule == FRX_RULE_TRAPEZOIDAL)) {
        // one pow() per node, operands shared between neighbouring lanes (29 nodes per warp)
        const int64_t total_virtual = (total + n_settings) / 2 + n_settings;
        const int64_t warps = (total_virtual + 28) / 29;
        FRX_LAUNCH(ctx, FRX_K_STENCIL,
                   frx_k5c_riesz_shared<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, ctx->stream>>>(
                       rule, n_settings, total_virtual, d_resolution, d_row_ptr, (const frx_k5_scalars*)d_scratch, d_offsets,
                       d_weights));
        return FRX_OK;
    }
    if (side == FRX_SIDE_RIESZ_CAPUTO) {
        const int64_t total_half = (total + n_settings) / 2;
        FRX_LAUNCH(ctx, FRX_K_STENCIL,
                   frx_k5b_riesz_half<<<(unsigned)((total_half + 127) / 128), 128, 0, ctx->stream>>>(
                       rule, n_settings, total_half, d_resolution, d_row_ptr, (const frx_k5_scalars*)d_scratch, d_offsets,
                       d_weights));
        return FRX_OK;
    }
    FRX_LAUNCH(ctx, FRX_K_STENCIL,
               frx_k5b_stencil_nodes<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(
                   rule, side, n_settings, total, d_resolution, d_row_ptr, (const frx_k5_scalars*)d_scratch, d_offsets,
                   d_weights));
    return FRX_OK;
}

extern "C" int frx_stencil_weights(frx_ctx* ctx, int rule, int side, int64_t n_settings, const double* d_alpha,
                                   const double* d_lf, const double* d_resolution, const int64_t* h_row_ptr,
                                   int32_t* d_offsets, double* d_weights) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_LAST, FRX_ERR_UNKNOWN_RULE, "unknown approximation key");
    FRX_REQUIRE(side >= FRX_SIDE_LEFT && side <= FRX_SIDE_RIESZ_CAPUTO, FRX_ERR_INVALID_ARG, "bad side");
    FRX_REQUIRE(n_settings >= 0 && n_settings < ((int64_t)1 << 31), FRX_ERR_INVALID_ARG, "bad n_settings");
    if (n_settings == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_lf && d_resolution && h_row_ptr && d_weights, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    // row_ptr travels through the workspace (it is host data describing the caller's layout): [row_ptr][scalars]
    const size_t bytes = frx_align_up(sizeof(int64_t) * (size_t)(n_settings + 1), 256);
    int rc = frx_ws_reserve(ctx, bytes + frx_k5_scratch_bytes(n_settings));
    if (rc) return rc;
    FRX_CUDA(cudaMemcpyAsync(ctx->ws, h_row_ptr, sizeof(int64_t) * (size_t)(n_settings + 1), cudaMemcpyHostToDevice,
                             ctx->stream));
    return frx_launch_stencil_weights(ctx, rule, side, n_settings, d_alpha, d_lf, d_resolution, (const int64_t*)ctx->ws,
                                      h_row_ptr[n_settings], (char*)ctx->ws + bytes, d_offsets, d_weights);
}

extern "C" int frx_stencil_weights_host(frx_ctx* ctx, int rule, int side, double alpha, double lf, double resolution,
                                        int32_t capacity, int32_t* h_offsets, double* h_weights, int32_t* count) {
    FRX_REQUIRE(ctx != nullptr && h_weights != nullptr && count != nullptr, FRX_ERR_INVALID_ARG, "NULL argument");
    int32_t first, cnt;
    int rc = frx_stencil_layout(rule, side, alpha, resolution, &first, &cnt);
    if (rc) return rc;
    *count = cnt;
    FRX_REQUIRE(capacity >= cnt, FRX_ERR_INVALID_ARG, "capacity smaller than the stencil length");
    FRX_ON_DEVICE(ctx);
    if (rule != FRX_RULE_GRUNWALD_LETNIKOV) {
        // one launch with the triple by value, one copy back; the offsets are first .. first + count - 1
        int rc2 = frx_ws_reserve(ctx, sizeof(double) * (size_t)cnt);
        if (rc2) return rc2;
        FRX_LAUNCH(ctx, FRX_K_STENCIL,
                   frx_k5_single<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(rule, side, alpha, lf, resolution,
                                                                                       (long long)first, (long long)cnt,
                                                                                       (double*)ctx->ws));
        FRX_CUDA(cudaMemcpyAsync(h_weights, ctx->ws, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
        if (h_offsets)
            for (int32_t k = 0; k < cnt; ++k) h_offsets[k] = first + k;
        FRX_CUDA(cudaStreamSynchronize(ctx->stream));
        return FRX_OK;
    }
    // workspace: [row_ptr (2 x i64) + scalars, written by frx_stencil_weights] [3 doubles] [weights] [offsets]
    size_t off_par = 256 + frx_align_up(frx_k5_scratch_bytes(1), 256), off_w = off_par + 64,
           off_o = off_w + frx_align_up(sizeof(double) * cnt, 64);
    size_t total = off_o + sizeof(int32_t) * (size_t)cnt;
    rc = frx_ws_reserve(ctx, total);
    if (rc) return rc;
    char* ws = (char*)ctx->ws;
    double par[3] = {alpha, lf, resolution};
    FRX_CUDA(cudaMemcpyAsync(ws + off_par, par, sizeof(par), cudaMemcpyHostToDevice, ctx->stream));
    int64_t row_ptr[2] = {0, cnt};
    const double* dpar = (const double*)(ws + off_par);
    // (frx_stencil_weights puts row_ptr at ws + 0; the regions above do not overlap it)
    rc = frx_stencil_weights(ctx, rule, side, 1, dpar, dpar + 1, dpar + 2, row_ptr, (int32_t*)(ws + off_o),
                             (double*)(ws + off_w));
    if (rc) return rc;
    FRX_CUDA(cudaMemcpyAsync(h_weights, ws + off_w, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_offsets)
        FRX_CUDA(cudaMemcpyAsync(h_offsets, ws + off_o, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    FRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

// ------------------------------------------------------------------------------------------------
// K0: per-alpha scalars (one warp per alpha), K1: weight tables (stand-alone entry; frx_history runs the same
// block routine inside its prepare kernel).  K1 grid = (ceil(ldc / 124), n_alpha), see frx_table.cuh.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) frx_k0_alpha_scalars(int rule, const double* __restrict__ alpha, double h,
                                                            frx_alpha_scalars* __restrict__ out) {
    __shared__ frx_k0_smem W;
    frx_grid_launch_dependents();   // the table kernel may start its power evaluation right away
    frx_alpha_scalars_block(W, rule, alpha[blockIdx.x], h, out + blockIdx.x);
}

int frx_launch_alpha_scalars(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, void* d_scalars) {
    FRX_LAUNCH(ctx, FRX_K_OTHER,
               frx_k0_alpha_scalars<<<(unsigned)n_alpha, 128, 0, ctx->stream>>>(rule, d_alpha, h,
                                                                               (frx_alpha_scalars*)d_scalars));
    return FRX_OK;
}
size_t frx_alpha_scalars_bytes(void) { return sizeof(frx_alpha_scalars); }

__global__ void __launch_bounds__(FRX_TAB_BLOCK, 8) frx_k1_weights_table(int rule, const double* __restrict__ alpha, double h,
                                                                           int64_t N, const frx_alpha_scalars* __restrict__ sc,
                                                                           double* __restrict__ cA, double* __restrict__ cB,
                                                                           int64_t ldc, double* __restrict__ w0,
                                                                           double* __restrict__ mult, int64_t ldw,
                                                                           double* __restrict__ cm1) {
    __shared__ frx_table_smem S;
    const int64_t a = blockIdx.y;
    frx_table_block(S, sc + a, rule, alpha[a], h, N, (int64_t)blockIdx.x * FRX_TAB_ENTRIES, 0, 0, 0, cA + a * ldc,
                    cB ? cB + a * ldc : nullptr, ldc, w0 + a * ldw, mult + a * ldw, cm1 ? cm1 + a : nullptr);
}

// Gruenwald-Letnikov tables (extension): one CTA per alpha walks the whole table as a chunked prefix product.
// cA[c_pad + x] = w_x (x < N; permuted like the other rules if c_perm), w0[i] = w_i, mult[i] = step_i ** (-alpha).
__global__ void __launch_bounds__(256) frx_k1_gl_table(const double* __restrict__ alpha, double h, int64_t N,
                                                       const frx_alpha_scalars* __restrict__ sc, int c_pad, int c_perm,
                                                       double* __restrict__ cA, int64_t ldc, double* __restrict__ w0,
                                                       double* __restrict__ mult, int64_t ldw, double* __restrict__ cm1) {
    __shared__ frx_gl_smem<256> GS;
    const int64_t a = blockIdx.x;
    const frx_alpha_scalars* S = sc + a;
    double* ca = cA + a * ldc;
    for (int64_t X = threadIdx.x; X < ldc; X += 256) ca[X] = 0.0;   // front padding and the tail beyond N
    if (threadIdx.x == 0 && cm1) cm1[a] = 0.0;
    __syncthreads();
    frx_gl_block<256>(GS, alpha[a], N + 1, [&](long long x, double w) {
        if (x < N) ca[frx_table_slot(c_pad, c_perm, x)] = w;
        double m = 0.0, wf = 0.0;
        if (x >= 1) {
            const double xd = (double)x, step = FRX_DIV(FRX_MUL(xd, h), xd);
            if (step == S->step[1]) m = S->mult[1];
            else if (step == S->step[0]) m = S->mult[0];
            else if (step == S->step[2]) m = S->mult[2];
            else m = frx_pow_cr(step, S->P.e_mult);
            wf = w;                                               // node number 0 of row i sits at distance i
        }
        w0[a * ldw + x] = wf;
        mult[a * ldw + x] = m;
    });
}

int frx_launch_gl_table(frx_ctx* ctx, int64_t n_alpha, const double* d_alpha, double h, int64_t N, const void* d_scalars,
                        int c_pad, int c_perm, double* d_cA, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                        double* d_cm1) {
    FRX_LAUNCH(ctx, FRX_K_WEIGHTS_TABLE,
               frx_k1_gl_table<<<(unsigned)n_alpha, 256, 0, ctx->stream>>>(d_alpha, h, N, (const frx_alpha_scalars*)d_scalars,
                                                                           c_pad, c_perm, d_cA, ldc, d_w0, d_mult, ldw, d_cm1));
    return FRX_OK;
}

int frx_launch_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                             double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                             double* d_cm1) {
    int rc = frx_ws_reserve(ctx, sizeof(frx_alpha_scalars) * (size_t)n_alpha);
    if (rc) return rc;
    rc = frx_launch_alpha_scalars(ctx, rule, n_alpha, d_alpha, h, ctx->ws);
    if (rc) return rc;
    if (rule == FRX_RULE_GRUNWALD_LETNIKOV)
        return frx_launch_gl_table(ctx, n_alpha, d_alpha, h, N, ctx->ws, 0, 0, d_cA, ldc, d_w0, d_mult, ldw, d_cm1);
    dim3 grid((unsigned)((ldc + FRX_TAB_ENTRIES - 1) / FRX_TAB_ENTRIES), (unsigned)n_alpha);
    FRX_LAUNCH(ctx, FRX_K_WEIGHTS_TABLE,
               FRX_CUDA(frx_launch_dependent(frx_k1_weights_table, grid, dim3(FRX_TAB_BLOCK), 0, ctx->stream, rule, d_alpha, h,
                                             N, (const frx_alpha_scalars*)ctx->ws, d_cA, d_cB, ldc, d_w0, d_mult, ldw,
                                             d_cm1)));
    return FRX_OK;
}

extern "C" int frx_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                                 double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                                 double* d_cm1) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_LAST, FRX_ERR_UNKNOWN_RULE, "unknown approximation key");
    FRX_REQUIRE(n_alpha >= 0 && n_alpha <= 65535 && N >= 1, FRX_ERR_INVALID_ARG, "bad n_alpha (<= 65535) or N");
    FRX_REQUIRE(ldc >= frx_table_ldc(N) && ldw >= N + 1, FRX_ERR_INVALID_ARG, "leading dimension too small");
    FRX_REQUIRE(d_alpha && d_cA && d_w0 && d_mult, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_REQUIRE(rule != FRX_RULE_SIMPSON || (d_cB && d_cm1), FRX_ERR_INVALID_ARG, "'simpson' needs d_cB and d_cm1");
    if (n_alpha == 0) return FRX_OK;
    FRX_ON_DEVICE(ctx);
    return frx_launch_weights_table(ctx, rule, n_alpha, d_alpha, h, N, d_cA, d_cB, ldc, d_w0, d_mult, ldw, d_cm1);
}
