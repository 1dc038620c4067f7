// K4, first path: band LDL^T WITHOUT pivoting -- sm_100a.  Two forms of the same idea live here:
//   frx_k4_band_bldl<T>  (shipped, second half of the file): block LDL^T with 8 x 8 block pivots, tiles in registers;
//   frx_k4_band_ldl<T>   (first form, FRX_K4_LDL=2): scalar pivots, tiles in shared memory; kept for comparison.
//
// The assembled operator (central o Riesz-Caputo o central) is a SYMMETRIC banded Toeplitz matrix;
// about three quarters of the (alpha, resolution) pairs of the cfg-4 sweep give a DEFINITE one.  For a definite matrix
// elimination without row interchanges is unconditionally backward stable (it is the Cholesky / LDL^T factorisation), and
// definiteness shows in the factorisation itself (Sylvester: the matrix is definite iff every pivot has the sign of the
// first).  So every system is first tried here.  A pivot of the wrong sign marks the system indefinite: the first form
// gives up at once, the shipped form keeps the solution only if its residual passes after at most six steps of iterative
// refinement.  What is rejected is appended to the launch's fail list, which the pivoted kernels of frx_solve_dmma.cuh then
// work through (device-side count, no host round trip).  Like LAPACK's dpbtrf(uplo='L') the routines read ONE triangle:
// the lower band a_m = A[i+m][i], m = 0..nb, mirrored where a diagonal tile needs its upper part.
//
// frx_k4_band_ldl<T>: without interchanges nothing fills in beyond the band and no row ever moves, which removes what the
// pivoted kernel spends its time on (ncu: 62 % of its instructions are the pivot search / row retirement chain):
//   * state = the T x T trailing block behind the current panel of 8 columns, lower tiles only: T(T+1)/2 tiles of 8 x 8
//     in shared memory (5 KB at nb <= 32 against 18 KB of pivoted window), every tile stored as its DMMA C fragment reads it;
//   * panel: lanes 0..7 hold the rows of the diagonal tile (+ rhs), lane l < 8T the row 8 + l below it.  Per column: the
//     pivot row is broadcast from a FIXED lane (shuffles), one reciprocal, one multiply, <= 8 FMAs -- the sub-diagonal rows
//     leave the loop as L21 (multipliers) and, one step earlier, as U12^T (the values before the division): both operands of
//     the trailing update come out of the same registers, the forward substitution rides along as a ninth column;
//   * trailing update C(i,j) -= L21(i) U12(j), i >= j: 2 mma.sync.m8n8k4.f64 per tile; the A fragment of row tile i and the
//     B fragment of column tile j have the SAME lane mapping (lane (g,cp) holds M[8t+g][2cp..2cp+1]), so 4T doubles per
//     lane feed all T(T+1)/2 tiles.  Tile (i,j) is written where the NEXT panel expects it (one tile up and left), tiles
//     that enter the band are built from the coefficient vector in the C fragment layout -- no ring, no copies;
//   * rows >= n: the system is padded with identity rows up to the next multiple of 8;
//   * per panel the rows of the diagonal tile (reciprocal pivot on the diagonal) and of U12^T go to the per-warp scratch in
//     L2 as dense blocks; back substitution in blocks of 8 rows (transpose-reduce + 8 x 8 triangle).
//
// Index bookkeeping prototyped in NumPy first: tools/proto_band_ldl.py, tools/proto_band_bldl.py.
#pragma once
#include "frx_solve_dmma.cuh"

namespace frx_k4 {

template <int T_>
struct LdlGeo {
    static constexpr int T = T_;
    static constexpr int NT = T * (T + 1) / 2;          // lower tiles of the trailing block
    static constexpr int OFF = 8 * T + 8;               // apad[OFF + k] = a_|k|, zero for nb < |k| <= OFF
    static constexpr int APAD = 2 * OFF + 2;
    static constexpr int NB_MAX = 8 * T;
    static constexpr int UBLK = (8 * T + 8) * 8;        // doubles of U scratch per panel: U11 rows, then the rows of U12^T
    __host__ __device__ static inline int xs_doubles(int n) { return ((n + 7) & ~7) + 8 * T + 2; }   // zero beyond n
    __host__ __device__ static inline int smem_doubles(int n) { return NT * 64 + 2 * 64 * T + APAD + xs_doubles(n); }
    __host__ __device__ static inline int64_t ustride(int n) { return (int64_t)((n + 7) >> 3) * UBLK; }
    __host__ __device__ static constexpr int tile(int i, int j) { return i * (i + 1) / 2 + j; }
};

// a_k of the assembled operator, k = column - row: the terms and the order of the dict-merging expansion
// (same expression as in frx_k4_band_dmma)
__device__ __forceinline__ double band_coefficient(const int k, const int res, const double ih, const double* rc) {
    double acc = 0.0;
    bool any = false;
    auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
    if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));
    if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));
    if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));
    if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));
    return acc;
}

// Back substitution on the panel blocks the factorisation left in the scratch: block of panel kb/8 =
// [U11: 8 rows of 8, reciprocal pivot on the diagonal][U12^T: 8T rows l of 8, entry q = U[kb + q][kb + 8 + l]].
// xs = transformed rhs on entry, the solution on exit (zero beyond n, allocated up to 8 * n_panels + 8T).
template <int T>
__device__ __forceinline__ void ldl_back_substitution(const int lane, const int n_panels, const double* __restrict__ U, double* xs) {
    using G = LdlGeo<T>;
    constexpr unsigned FULL = 0xffffffffu;
    const int gq = lane >> 2;
    auto load_block = [&](int pnl, double (&uu)[8], double (&dd)[8]) {
        if (pnl >= 0) {
            const double* blk = U + (int64_t)pnl * G::UBLK;
            const double2* dsrc = reinterpret_cast<const double2*>(blk + gq * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 v = __ldcg(dsrc + q);
                dd[2 * q] = v.x;
                dd[2 * q + 1] = v.y;
            }
            if (lane < 8 * T) {
                const double2* usrc = reinterpret_cast<const double2*>(blk + 64 + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v = __ldcg(usrc + q);
                    uu[2 * q] = v.x;
                    uu[2 * q + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) uu[q] = 0.0;
            }
        }
    };
    double un[8], dn[8];
    load_block(n_panels - 1, un, dn);
    for (int pnl = n_panels - 1; pnl >= 0; --pnl) {
        const int kb = pnl * 8;
        double sacc[8], d[8];
        const double xl = lane < 8 * T ? xs[kb + 8 + lane] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            sacc[q] = un[q] * xl;
            d[q] = dn[q];
        }
        load_block(pnl - 1, un, dn);
        // transpose-reduce: 8 sums over 32 lanes -> lane group gq holds the sum of row kb + gq
        double v4[4], v2[2], v1;
        {
            const bool up = lane & 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double keep = up ? sacc[4 + i] : sacc[i], send = up ? sacc[i] : sacc[4 + i];
                v4[i] = keep + __shfl_xor_sync(FULL, send, 16);
            }
            const bool up8 = lane & 8;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double keep = up8 ? v4[2 + i] : v4[i], send = up8 ? v4[i] : v4[2 + i];
                v2[i] = keep + __shfl_xor_sync(FULL, send, 8);
            }
            const bool up4 = lane & 4;
            const double keep = up4 ? v2[1] : v2[0], send = up4 ? v2[0] : v2[1];
            v1 = keep + __shfl_xor_sync(FULL, send, 4);
            v1 += __shfl_xor_sync(FULL, v1, 2);
            v1 += __shfl_xor_sync(FULL, v1, 1);
        }
        double rhsq = xs[kb + gq] - v1;
        double rinvq = 0.0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) rinvq = cc == gq ? d[cc] : rinvq;
        __syncwarp();
#pragma unroll
        for (int qq = 7; qq >= 0; --qq) {
            const double xv = __shfl_sync(FULL, rhsq * rinvq, qq * 4);
            if (gq < qq) rhsq = fma(-d[qq], xv, rhsq);
            if (lane == qq * 4) xs[kb + qq] = xv;
        }
        __syncwarp();
    }
}

template <int T>
__global__ void __launch_bounds__(32) frx_k4_band_ldl(const BandParams p) {
    using G = LdlGeo<T>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int OFF = G::OFF;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, g = lane >> 2, cp = lane & 3, n = p.n;
    double* SL = sm;                        // SL[tile(i,j) * 64 + row * 8 + col], i >= j: trailing block behind the panel
    double* XS = SL + G::NT * 64;           // XS[l * 8 + c] = -L21[l][c]   (A operand)
    double* YS = XS + 64 * T;               // YS[l * 8 + c] = U12[c][l]    (B operand)
    double* apad = YS + 64 * T;
    double* xs = apad + G::APAD;            // rhs, then transformed rhs, then the solution
    double* U = p.uscratch + (int64_t)blockIdx.x * p.ustride_cta;
    const int n_panels = (n + 7) >> 3;
    const int64_t n_sys = p.n_sys_dev ? (int64_t)*p.n_sys_dev : p.n_sys;
    for (int64_t si = blockIdx.x; si < n_sys; si += gridDim.x) {
        const int64_t s = p.order[si];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const double* rhs = p.rhs + s * p.ldb;
        double* xo = p.x + s * p.ldx;
        __syncwarp();
        for (int e = lane; e < G::APAD; e += 32) apad[e] = 0.0;
        for (int i = lane; i < G::xs_doubles(n); i += 32) xs[i] = i < n ? rhs[i] : 0.0;
        __syncwarp();
        for (int m = lane; m <= nb; m += 32) {               // the lower band, mirrored
            const double a = band_coefficient(-m, res, ih, rc);
            apad[OFF + m] = a;
            apad[OFF - m] = a;
        }
        __syncwarp();
        // entries (R, C), (R, C + 1) of the padded matrix, |R - C| <= OFF - 1
        auto fresh2 = [&](const int R, const int C, const bool edge) -> double2 {
            double2 v;
            v.x = apad[OFF + R - C];
            v.y = apad[OFF + R - C - 1];
            if (edge) {
                if (R >= n || C >= n) v.x = R == C ? 1.0 : 0.0;
                if (R >= n || C + 1 >= n) v.y = R == C + 1 ? 1.0 : 0.0;
            }
            return v;
        };
#pragma unroll
        for (int i = 0; i < T; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                *reinterpret_cast<double2*>(SL + G::tile(i, j) * 64 + g * 8 + 2 * cp) = fresh2(8 * i + g, 8 * j + 2 * cp, 8 * T > n);
        const double sgn = apad[OFF] < 0.0 ? -1.0 : 1.0;
        bool bad = false;
        __syncwarp();
        for (int pnl = 0; pnl < n_panels; ++pnl) {
            const int k0 = pnl * 8;
            const bool edge = k0 + 8 * T + 8 > n;
            // ================= panel: lanes 0..7 = rows of the diagonal tile, lane l < 8T = row k0 + 8 + l ==========
            double d[9], r[9], x[8];
            if (lane < 8) {
                const double2* src = reinterpret_cast<const double2*>(SL + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v = src[q];
                    d[2 * q] = v.x;
                    d[2 * q + 1] = v.y;
                }
                d[8] = xs[k0 + lane];
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) d[q] = 0.0;
            }
            const int R = k0 + 8 + lane;
            if (lane < 8 * T) {
                const int t = lane >> 3;
                if (t + 1 <= T - 1) {
                    const double2* src = reinterpret_cast<const double2*>(SL + ((t + 1) * (t + 2) / 2) * 64 + (lane & 7) * 8);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 v = src[q];
                        r[2 * q] = v.x;
                        r[2 * q + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) r[c] = (!edge || (R < n && k0 + c < n)) ? apad[OFF + 8 + lane - c] : 0.0;
                }
                r[8] = xs[R];
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) r[q] = 0.0;
            }
            double myrinv = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double pv[9];
#pragma unroll
                for (int cc = c; cc < 9; ++cc) pv[cc] = __shfl_sync(FULL, d[cc], c);
                const double pval = pv[c];
                if (k0 + c < n && !(pval * sgn > 0.0)) bad = true;      // warp-uniform
                const double rinv = 1.0 / pval;
                if (lane == c) myrinv = rinv;
                if (lane > c) {                                          // rows below the pivot inside the diagonal tile
                    const double l = d[c] * rinv;
#pragma unroll
                    for (int cc = c + 1; cc < 9; ++cc) d[cc] = fma(-l, pv[cc], d[cc]);
                }
                const double l = r[c] * rinv;                            // r[c] stays: it is U12[c][this row]
                x[c] = -l;
#pragma unroll
                for (int cc = c + 1; cc < 9; ++cc) r[cc] = fma(-l, pv[cc], r[cc]);
            }
            if (bad) break;
            // this panel's block of the U scratch: the rows of the diagonal tile (reciprocal pivot on the diagonal; what is
            // left of it are multipliers, never read), then the rows of U12^T; the transformed right-hand side stays in xs
            double* ublk = U + (int64_t)pnl * G::UBLK;
            if (lane < 8) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                    if (lane == cc) d[cc] = myrinv;
                double2* ud = reinterpret_cast<double2*>(ublk + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) ud[q] = make_double2(d[2 * q], d[2 * q + 1]);
                xs[k0 + lane] = d[8];
            }
            if (lane < 8 * T) {
                xs[R] = r[8];
                double2* ud = reinterpret_cast<double2*>(ublk + 64 + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) ud[q] = make_double2(r[2 * q], r[2 * q + 1]);
                double2* xd = reinterpret_cast<double2*>(XS + lane * 8);
                double2* yd = reinterpret_cast<double2*>(YS + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    xd[q] = make_double2(x[2 * q], x[2 * q + 1]);
                    yd[q] = make_double2(r[2 * q], r[2 * q + 1]);
                }
            }
            __syncwarp();
            // ================= trailing update on the tensor cores, written one tile up and left ====================
            double2 xa[T], yb[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                xa[t] = *reinterpret_cast<const double2*>(XS + (8 * t + g) * 8 + 2 * cp);
                yb[t] = *reinterpret_cast<const double2*>(YS + (8 * t + g) * 8 + 2 * cp);
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double2 cf;
                    if (i + 1 <= T - 1) cf = *reinterpret_cast<const double2*>(SL + G::tile(i + 1, j + 1) * 64 + g * 8 + 2 * cp);
                    else cf = fresh2(k0 + 8 + 8 * i + g, k0 + 8 + 8 * j + 2 * cp, edge);
                    dmma884(cf.x, cf.y, xa[i].x, yb[j].x);
                    dmma884(cf.x, cf.y, xa[i].y, yb[j].y);
                    *reinterpret_cast<double2*>(SL + G::tile(i, j) * 64 + g * 8 + 2 * cp) = cf;
                }
            }
            __syncwarp();
        }
        if (!bad) {
            __syncwarp();
            ldl_back_substitution<T>(lane, n_panels, U, xs);
            for (int i = lane; i < n; i += 32) xo[i] = xs[i];
            if (lane == 0 && p.info) p.info[s] = 0;
            __syncwarp();
            discard_l2(U, G::ustride(n) * 8, lane, 32);
        } else if (lane == 0) {
            p.fail_list[atomicAdd(p.fail_count, 1)] = (int32_t)s;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second form of the same attempt: BLOCK LDL^T with 8 x 8 block pivots, everything in registers and on the tensor cores.
//
// ncu on frx_k4_band_ldl (profiles/r02_k4_band_ldl_v1_ncu_full.txt): the shuffle / shared-memory pipe is the limit
// (stalls: short scoreboard 4.6 and MIO throttle 2.8 per issue) -- ~470 shared-memory wavefronts and shuffles per panel
// for the pivot-row broadcasts, the row <-> fragment layout changes and the tile traffic.  Here no tile ever leaves the
// DMMA C-fragment layout (lane (g, cp) holds M[g][2cp], M[g][2cp+1]):
//   * A = L_B D_B L_B^T with unit BLOCK lower L_B: the diagonal tile D_p becomes MINUS its inverse in place, by eight
//     symmetric sweeps without pivoting on the fragment itself (per sweep, T >= 2: four 64-bit shuffles -- the lane's
//     column-c entry, the pivot row's two entries, the pivot --, one reciprocal, three multiplies, two FMAs; T = 1: one
//     shuffle and ONE DMMA on in-lane operands) -- the pivots are those of the scalar LDL^T, so the definiteness test is
//     unchanged;
//   * L21(t) = A21(t) Dinv: two DMMAs per tile with A21(t) AS IT IS (a C fragment is a valid A operand when the B operand
//     holds rows 2cp / 2cp+1 of the right factor, and Dinv is symmetric: its own fragment); the result is again a C fragment;
//   * trailing update C(i,j) += (-L21(i)) A21(j)^T: A = the fragment just computed, B = the fragment of A21(j): two DMMAs,
//     accumulating in the registers of tile (i+1, j+1), which become tile (i, j) of the next panel; tiles entering the band
//     are T + 1 constant Toeplitz fragments kept in registers (masked at the end of the matrix: the system is padded with
//     +-identity rows, of the sign of a_0, up to a multiple of 8);
//   * right-hand side as replicated 8-vectors: b(t) += (-L21(t)) z (two FMAs and two xor-shuffles per tile);
//   * back substitution x_p = Dinv z_p - sum_t L21(t)^T x_{p+t}: two FMAs per tile on the stored fragments and ONE
//     reduction over the row index (three xor-shuffles of two doubles) -- no triangular solve, no dependent chain.
//   * an INDEFINITE system (a pivot of the wrong sign) is kept only on its residual: up to six steps of iterative
//     refinement with the same factors (residual with one lane per row from the coefficient vector, forward substitution
//     with the stored -L21 fragments, the same back substitution accumulating into x), each followed by the test
//     ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||) in max norms; what fails goes to the fail list.
// Shared memory: the coefficient vector, x, the residual and two right-hand-side buffers (the next system's rhs arrives by
// cp.async while this one is solved).  ~100 shuffles per panel (52 at T = 1), 2T + T(T+1) DMMAs (+ 8 at T = 1).  The explicit
// inverse of the 8 x 8 tile leaves residuals of 1e-15 on the operator's definite matrices (tools/proto_band_bldl.py checks
// that on oracle coefficients together with the index bookkeeping).
template <int T_>
struct BldlGeo {
    static constexpr int T = T_;
    static constexpr int OFF = 8 * T + 8;               // apad[OFF + k] = a_|k|, zero for nb < |k| <= OFF
    static constexpr int APAD = 2 * OFF + 2;
    static constexpr int UBLK = (T + 1) * 64 + 16;      // per panel: -Dinv, -L21(1..T) as fragments, z (8) + pad to 128 bytes
    static constexpr int XPAD = 32;                     // zeros in front of the solution vector (the band reaches back nb <= 32)
    __host__ __device__ static inline int npad(int n) { return (n + 31) & ~31; }
    __host__ __device__ static inline int xs_doubles(int n) { return XPAD + npad(n) + 8 * T + 32; }   // x, zero on both sides
    __host__ __device__ static inline int rs_doubles(int n) { return npad(n) + 8 * T + 8; }           // residual, zero beyond n
    __host__ __device__ static inline int smem_doubles(int n) { return APAD + xs_doubles(n) + 3 * rs_doubles(n); }
    __host__ __device__ static inline int64_t ustride(int n) { return (int64_t)((n + 7) >> 3) * UBLK; }
};

// 8-byte asynchronous copy global -> shared (LDGSTS): the NEXT system's right-hand side travels while this one is solved
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 1 / p for the pivots: MUFU seed (upper 20 mantissa bits) and ONE third-order step (three DFMAs), no special-case branch.  p = 0, denormal, Inf or NaN give
// Inf / NaN -- such a pivot has already failed the sign test (p * sgn > 1e-300), and the residual test rejects what follows.
__device__ __forceinline__ double pivot_reciprocal(const double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);            // |e| <= 2^-20: r (1 + e + e^2) is 1 / p to e^3 < 2^-60
    return fma(r, fma(e, e, e), r);
}

// -v by flipping the sign bit of the high word (an integer instruction: the compiler's own negation is a DADD)
__device__ __forceinline__ double flip_sign(const double v) {
    return __hiloint2double(__double2hiint(v) ^ (int)0x80000000, __double2loint(v));
}

// panels per trip of the factorisation loop (the tile renaming at its back edge costs ~40 moves per trip at T = 4)
#ifndef FRX_BLDL_PANEL_UNROLL
#define FRX_BLDL_PANEL_UNROLL 1
#endif
constexpr int kBldlPanelUnroll = FRX_BLDL_PANEL_UNROLL;
#ifndef FRX_BLDL_PREDICT
#define FRX_BLDL_PREDICT 0
#endif
constexpr bool kBldlPredict = FRX_BLDL_PREDICT != 0;      // see the sweeps of frx_k4_band_bldl

// Iterative refinement of an indefinite system: at most this many steps (each ~1/3 of a factorisation; the pivoted LU that
// takes over otherwise costs ~4 factorisations and runs at half the occupancy).  On the cfg-4 matrices every indefinite
// system of n <= 128 passes after one step; at n = 256 two steps accept 67 %, six 89 % (tools/proto_band_bldl.py --refine).
constexpr int kBldlRefineSteps = 6;

// NaN-propagating maximum (fmax would drop a NaN: a broken factorisation must not pass the acceptance test)
__device__ __forceinline__ double nanmax(const double m, const double v) { return (v > m || v != v) ? v : m; }
__device__ __forceinline__ double warp_nanmax(double m) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = nanmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    return m;
}

// Forward substitution L_B z = r for a NEW right-hand side (iterative refinement) with the stored -L21 fragments;
// r in shared memory (zero beyond n), z_p goes to the panel blocks of the scratch.
template <int T>
__device__ __forceinline__ void bldl_forward(const int lane, const int n_panels, double* U, const double* rs) {
    using G = BldlGeo<T>;
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, cp = lane & 3, fo = g * 8 + 2 * cp;
    double bv[T];
#pragma unroll
    for (int t = 0; t < T; ++t) bv[t] = rs[8 * t + g];
    double2 Xnn[T + 1];
#pragma unroll
    for (int t = 1; t <= T; ++t) Xnn[t] = __ldcg(reinterpret_cast<const double2*>(U + t * 64 + fo));
#pragma unroll 1
    for (int pnl = 0; pnl < n_panels; ++pnl) {
        double2 Xn[T + 1];
#pragma unroll
        for (int t = 1; t <= T; ++t) Xn[t] = Xnn[t];
        if (pnl + 1 < n_panels) {
            const double* blk = U + (int64_t)(pnl + 1) * G::UBLK;
#pragma unroll
            for (int t = 1; t <= T; ++t) Xnn[t] = __ldcg(reinterpret_cast<const double2*>(blk + t * 64 + fo));
        }
        const double z = bv[0];
        if (cp == 0) U[(int64_t)pnl * G::UBLK + (T + 1) * 64 + g] = z;
        const double z0 = __shfl_sync(FULL, z, 8 * cp), z1 = __shfl_sync(FULL, z, 8 * cp + 4);
        double nbv[T];
#pragma unroll
        for (int t = 1; t <= T; ++t) {
            double part = fma(Xn[t].x, z0, Xn[t].y * z1);
            part += __shfl_xor_sync(FULL, part, 1);
            part += __shfl_xor_sync(FULL, part, 2);
            nbv[t - 1] = (t <= T - 1 ? bv[t < T ? t : 0] : rs[pnl * 8 + 8 * T + g]) + part;
        }
#pragma unroll
        for (int t = 0; t < T; ++t) bv[t] = nbv[t];
    }
}

// Back substitution x_p = Dinv z_p - sum_t L21(t)^T x_{p+t} on the panel blocks of the scratch; the solution goes to
// shared memory (xs[XPAD + i]), added to what is there when `accumulate` (a refinement step).
template <int T>
__device__ __forceinline__ void bldl_backward(const int lane, const int n_panels, const double* U, double* xs, const bool accumulate) {
    using G = BldlGeo<T>;
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, cp = lane & 3, fo = g * 8 + 2 * cp;
    double xv[T];                                    // x of block rows p+1 .. p+T, entry g (replicated over cp)
#pragma unroll
    for (int t = 0; t < T; ++t) xv[t] = 0.0;
    double2 nDn, Xnn[T + 1];
    double zn;
    {
        const double* blk = U + (int64_t)(n_panels - 1) * G::UBLK;
        nDn = __ldcg(reinterpret_cast<const double2*>(blk + fo));
#pragma unroll
        for (int t = 1; t <= T; ++t) Xnn[t] = __ldcg(reinterpret_cast<const double2*>(blk + t * 64 + fo));
        zn = __ldcg(blk + (T + 1) * 64 + g);
    }
#pragma unroll 1
    for (int pnl = n_panels - 1; pnl >= 0; --pnl) {
        const double2 nD = nDn;
        double2 Xn[T + 1];
#pragma unroll
        for (int t = 1; t <= T; ++t) Xn[t] = Xnn[t];
        const double z = zn;
        if (pnl > 0) {
            const double* blk = U + (int64_t)(pnl - 1) * G::UBLK;
            nDn = __ldcg(reinterpret_cast<const double2*>(blk + fo));
#pragma unroll
            for (int t = 1; t <= T; ++t) Xnn[t] = __ldcg(reinterpret_cast<const double2*>(blk + t * 64 + fo));
            zn = __ldcg(blk + (T + 1) * 64 + g);
        }
        double ax = -nD.x * z, ay = -nD.y * z;       // Dinv[g][j] z[g], j = 2cp, 2cp+1
#pragma unroll
        for (int t = 1; t <= T; ++t) {
            ax = fma(Xn[t].x, xv[t - 1], ax);
            ay = fma(Xn[t].y, xv[t - 1], ay);
        }
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {    // sum over the row index g
            ax += __shfl_xor_sync(FULL, ax, off);
            ay += __shfl_xor_sync(FULL, ay, off);
        }
        // x_p[2cp], x_p[2cp+1] in every lane -> x_p[g] in every lane
        const double sx = __shfl_sync(FULL, ax, (lane & ~3) | (g >> 1)), sy = __shfl_sync(FULL, ay, (lane & ~3) | (g >> 1));
        const double xg = (g & 1) ? sy : sx;
        if (cp == 0) {
            double* dst = xs + G::XPAD + pnl * 8 + g;
            *dst = accumulate ? *dst + xg : xg;
        }
#pragma unroll
        for (int t = T - 1; t >= 1; --t) xv[t] = xv[t - 1];
        xv[0] = xg;
    }
}

// Resident warps per SM the register allocation is capped for (one warp per CTA), per T = ceil(nb / 8).  Measured per
// 1e5 systems of n = 128 (tools/time_k4.py): T = 1 at 20 / 24 / 32 warps 1.025 / 0.989 / 0.968 ms, T = 2 at 16 / 20 / 24 warps
// 1.406 / 1.337 / 1.375 ms, T = 3 at 16 / 20 / 24 warps 1.600 / 1.619 / 1.755 ms, T = 4 at 12 / 16 / 20 warps 2.041 / 1.922 / 2.300 ms.
#ifndef FRX_BLDL_OCC_T1
#define FRX_BLDL_OCC_T1 32
#endif
#ifndef FRX_BLDL_OCC_T2
#define FRX_BLDL_OCC_T2 20
#endif
#ifndef FRX_BLDL_OCC_T3
#define FRX_BLDL_OCC_T3 16
#endif
#ifndef FRX_BLDL_OCC_T4
#define FRX_BLDL_OCC_T4 16
#endif
constexpr int bldl_warps_per_sm(const int T) {
    return T == 1 ? FRX_BLDL_OCC_T1 : T == 2 ? FRX_BLDL_OCC_T2 : T == 3 ? FRX_BLDL_OCC_T3 : FRX_BLDL_OCC_T4;
}

template <int T>
__global__ void __launch_bounds__(32, bldl_warps_per_sm(T)) frx_k4_band_bldl(const BandParams p) {
    using G = BldlGeo<T>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int OFF = G::OFF;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, g = lane >> 2, cp = lane & 3, n = p.n;
    double* apad = sm;
    double* xs = apad + G::APAD;                         // xs[XPAD + i] = x_i, zero before 0 and beyond n
    double* rs = xs + G::xs_doubles(n);                  // residual of a refinement step, zero beyond n
    // the right-hand side, zero beyond n: read ONCE from global memory (which may be page-locked host memory:
    // frx_assemble_solve_host), one system ahead by asynchronous copies into the other of two buffers
    double* bs = rs + G::rs_doubles(n);
    double* bs_next = bs + G::rs_doubles(n);
    double* U = p.uscratch + (int64_t)blockIdx.x * p.ustride_cta;
    const int n_panels = (n + 7) >> 3;
    const int fo = g * 8 + 2 * cp;                       // this lane's place inside a stored fragment
    const int64_t n_sys = p.n_sys_dev ? (int64_t)*p.n_sys_dev : p.n_sys;
    for (int e = n + lane; e < G::rs_doubles(n); e += 32) bs[e] = bs_next[e] = 0.0;
    if ((int64_t)blockIdx.x < n_sys) {
        const double* rhs0 = p.rhs + (int64_t)p.order[blockIdx.x] * p.ldb;
        for (int e = lane; e < n; e += 32) cp_async8(bs + e, rhs0 + e);
    }
    // Work distribution: the first system of a warp is its block index, every further one comes from the launch's counter.
    // Systems differ in cost (an indefinite one takes ~1.7x a definite one), and with ~10 systems per warp a static
    // assignment leaves the slowest warp ~25 % behind the mean.
    int64_t si_next = blockIdx.x;
    for (;;) {
        const int64_t si = si_next;
        if (si >= n_sys) break;
        {
            int nx = 0;
            if (lane == 0) nx = atomicAdd(p.queue, 1);
            si_next = (int64_t)gridDim.x + __shfl_sync(FULL, nx, 0);
        }
        const int64_t s = p.order[si];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        double* xo = p.x + s * p.ldx;
        cp_async_wait_all();                                 // this system's right-hand side has landed in bs
        __syncwarp();
        if (si_next < n_sys) {
            const double* rhs1 = p.rhs + (int64_t)p.order[si_next] * p.ldb;
            for (int e = lane; e < n; e += 32) cp_async8(bs_next + e, rhs1 + e);
        }
        for (int e = lane; e < G::APAD; e += 32) apad[e] = 0.0;
        for (int e = lane; e < G::xs_doubles(n); e += 32) xs[e] = 0.0;
        __syncwarp();
        double amax = 0.0;
        for (int m = lane; m <= nb; m += 32) {               // the lower band, mirrored
            const double a = band_coefficient(-m, res, ih, rc);
            apad[OFF + m] = a;
            apad[OFF - m] = a;
            amax = fmax(amax, fabs(a));
        }
        __syncwarp();
        // the tile on tile-diagonal d of the (infinite) Toeplitz matrix, as this lane's fragment
        double2 F[T + 1];
#pragma unroll
        for (int d = 0; d <= T; ++d) F[d] = make_double2(apad[OFF + 8 * d + g - 2 * cp], apad[OFF + 8 * d + g - 2 * cp - 1]);
        // ... and what the padding beyond n makes of it (diagonal entries of the sign of a_0, so that a padded pivot passes
        // the sign test like any other): (R, C) = row and column of the fragment's first entry
        const double sgn = apad[OFF] < 0.0 ? -1.0 : 1.0;
        auto masked = [&](double2 v, const int R, const int C) -> double2 {
            if (R >= n || C >= n) v.x = R == C ? sgn : 0.0;
            if (R >= n || C + 1 >= n) v.y = R == C + 1 ? sgn : 0.0;
            return v;
        };
        double2 S[T][T];                                     // S[i][j], i >= j: trailing block behind the panel
        double bv[T];                                        // right-hand side of block row t, entry g (replicated over cp)
#pragma unroll
        for (int i = 0; i < T; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) S[i][j] = 8 * T > n ? masked(F[i - j], 8 * i + g, 8 * j + 2 * cp) : F[i - j];
            bv[i] = bs[8 * i + g];
        }
        double b_in = bs[8 * T + g];                         // the block row that enters with panel 0
        // high word of the smallest pivot times sgn(a_0) so far (warp-uniform); a NaN pivot makes every later entry NaN and
        // is caught on the last -Dinv below
        int piv_min = 0x7ff00000;
        bool nan_seen = false;
        const int sgn_bit = apad[OFF] < 0.0 ? (int)0x80000000 : 0;
#pragma unroll kBldlPanelUnroll
        for (int pnl = 0; pnl < n_panels; ++pnl) {
            const int k0 = pnl * 8;
            const bool edge = k0 + 8 * T + 8 > n;
            const double b_cur = b_in;
            b_in = bs[k0 + 8 * T + 8 + g];                                     // one panel ahead
            // ---- minus the inverse of the diagonal tile, in place in its fragment ------------------------------------------
            // Eight SWEEPS, sweep(c): p = M[c][c];  M[i][j] -= M[i][c] M[c][j] / p;  M[i][c] = M[c][i] = M[i][c] / p;
            // M[c][c] = -1 / p.  The tile stays symmetric through all eight; after them it is -Dinv.
            // -DFRX_BLDL_PREDICT=1 (measured, off): the NEXT pivot formed ahead of the update from the two entries it depends
            // on, p_{c+1} = M[c+1][c+1] - M[c+1][c]^2 / p_c (both fetched with the other shuffles of sweep c), so that its
            // reciprocal is under way while the rank-1 update of sweep c is applied -- the dependent chain pivot -> reciprocal
            // -> update -> shuffle -> pivot shrinks by 40 %, and the kernels get 7-12 % SLOWER (1.37 -> 1.47, 1.61 -> 1.81,
            // 1.92 -> 2.11 ms per 1e5 systems at nb <= 15 / 23 / 31): one more shuffle, multiply and FMA per sweep cost more
            // than the chain did.  The wide bins are bound by what they issue to the FP64 pipe, not by this latency.
            double2 M = S[0][0];
            double r_next = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int hc = c >> 1;
                const double own = (c & 1) ? M.y : M.x;
                double r;
                if (c == 0 || !kBldlPredict) {
                    const double pv = __shfl_sync(FULL, own, c * 4 + hc);      // M[c][c]
                    // smallest pivot seen, as the high word of p * sgn(a_0) (integer order = order of the doubles there;
                    // a pivot of the other sign has the sign bit set and is negative as an integer)
                    piv_min = min(piv_min, __double2hiint(pv) ^ sgn_bit);
                    r = pivot_reciprocal(pv);
                } else {
                    r = r_next;
                }
                if (kBldlPredict && c < 7) {
                    const int c1 = c + 1;
                    const double dn = __shfl_sync(FULL, (c1 & 1) ? M.y : M.x, c1 * 4 + (c1 >> 1));   // M[c+1][c+1]
                    const double bn = __shfl_sync(FULL, own, c1 * 4 + hc);                           // M[c+1][c]
                    const double pn = fma(-(bn * bn), r, dn);
                    piv_min = min(piv_min, __double2hiint(pn) ^ sgn_bit);
                    r_next = pivot_reciprocal(pn);
                }
                const bool rowc = g == c, atc = cp == hc;
                double dx = M.x, dy = M.y;
                if (T == 1) {
                    // narrow bands are bound by instruction issue: the rank-1 update as ONE DMMA whose operands are already
                    // in this lane (A = the fragment component that holds column c, k = 2cp + (c & 1);  B = -M[n][c] / p in
                    // the four lanes (n, c / 2) that own column c, zero elsewhere) -- one shuffle per sweep instead of four.
                    // Row and column c, which the DMMA leaves at ~0, are the lane's own old entries times 1 / p.
                    const double sx = M.x * r, sy = M.y * r;
                    const double so = (c & 1) ? sy : sx;                       // M[g][c] / p
                    dmma884(dx, dy, own, atc ? flip_sign(so) : 0.0);
                    if (rowc) { dx = sx; dy = sy; }
                    if (atc && !(c & 1)) dx = rowc ? flip_sign(r) : so;
                    if (atc && (c & 1)) dy = rowc ? flip_sign(r) : so;
                } else {
                    // wider bands are bound by the latency of this chain, and a DMMA in it is slower than shuffles + DFMAs
                    // (measured: 1.45 -> 1.85 ms per 1e5 systems at nb <= 15)
                    const double f = __shfl_sync(FULL, own, (lane & ~3) | hc);  // M[g][c]
                    const double prx = __shfl_sync(FULL, M.x, c * 4 + cp);     // M[c][2cp]
                    const double pry = __shfl_sync(FULL, M.y, c * 4 + cp);     // M[c][2cp+1]
                    const double px = prx * r, py = pry * r, fr = f * r;
                    dx = rowc ? px : fma(-f, px, dx);
                    dy = rowc ? py : fma(-f, py, dy);
                    if (atc && !(c & 1)) dx = rowc ? flip_sign(r) : fr;
                    if (atc && (c & 1)) dy = rowc ? flip_sign(r) : fr;
                }
                M.x = dx;
                M.y = dy;
            }
            const double2 nD = M;                                              // -Dinv (symmetric)
            nan_seen |= nD.x != nD.x;
            // ---- -L21(t) = A21(t) (-Dinv) -----------------------------------------------------------------------------
            double2 A21[T + 1], Xn[T + 1];
#pragma unroll
            for (int t = 1; t <= T; ++t) {
                if (t <= T - 1) A21[t] = S[t < T ? t : 0][0];
                else A21[t] = edge ? masked(F[T], k0 + 8 * T + g, k0 + 2 * cp) : F[T];
                Xn[t] = make_double2(0.0, 0.0);
                dmma884(Xn[t].x, Xn[t].y, A21[t].x, nD.x);
                dmma884(Xn[t].x, Xn[t].y, A21[t].y, nD.y);
            }
            // ---- this panel's block of the scratch ---------------------------------------------------------------------
            double* ublk = U + (int64_t)pnl * G::UBLK;
            *reinterpret_cast<double2*>(ublk + fo) = nD;
#pragma unroll
            for (int t = 1; t <= T; ++t) *reinterpret_cast<double2*>(ublk + t * 64 + fo) = Xn[t];
            const double z = bv[0];
            if (cp == 0) ublk[(T + 1) * 64 + g] = z;
            // ---- right-hand side: b(t) += (-L21(t)) z -------------------------------------------------------------------
            const double z0 = __shfl_sync(FULL, z, 8 * cp), z1 = __shfl_sync(FULL, z, 8 * cp + 4);    // z[2cp], z[2cp+1]
            double nbv[T];
#pragma unroll
            for (int t = 1; t <= T; ++t) {
                double part = fma(Xn[t].x, z0, Xn[t].y * z1);
                part += __shfl_xor_sync(FULL, part, 1);
                part += __shfl_xor_sync(FULL, part, 2);
                nbv[t - 1] = (t <= T - 1 ? bv[t < T ? t : 0] : b_cur) + part;
            }
            // ---- trailing update: tile (i+1, j+1) of this panel becomes tile (i, j) of the next ---------------------------
            double2 nS[T][T];
#pragma unroll
            for (int i = 0; i < T; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double2 c;
                    if (i + 1 <= T - 1) c = S[i + 1 < T ? i + 1 : 0][j + 1 < T ? j + 1 : 0];
                    else c = edge ? masked(F[i - j], k0 + 8 + 8 * i + g, k0 + 8 + 8 * j + 2 * cp) : F[i - j];
                    dmma884(c.x, c.y, Xn[i + 1].x, A21[j + 1].x);
                    dmma884(c.x, c.y, Xn[i + 1].y, A21[j + 1].y);
                    nS[i][j] = c;
                }
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) S[i][j] = nS[i][j];
                bv[i] = nbv[i];
            }
        }
        __syncwarp();                                        // the scratch stores of all lanes are visible to the warp
        bldl_backward<T>(lane, n_panels, U, xs, false);
        __syncwarp();
        // A definite matrix (every pivot of the sign of the first) needs nothing else: elimination without interchanges is
        // backward stable for it.  An indefinite one is ACCEPTED ONLY ON ITS RESIDUAL: up to kBldlRefineSteps steps of iterative refinement
        // with the same factors, each followed by the test  ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||)  (max norms; NaN fails).
        // every pivot p * sgn(a_0) > 1e-300 (0x01a56e1f = high word of 1e-300), none of them NaN
        const bool indef = piv_min <= 0x01a56e1f || __any_sync(FULL, nan_seen);
        bool ok = !indef;
        if (indef) {
            amax = warp_nanmax(amax);
            double r_prev = 1e300;
#pragma unroll 1
            for (int it = 0; it <= kBldlRefineSteps && !ok; ++it) {
                double rmax = 0.0, xmax = 0.0, bmax = 0.0;
                for (int i = lane; i < G::rs_doubles(n); i += 32) {
                    double acc = 0.0;
                    if (i < n) {
                        const double bi = bs[i];
                        acc = bi;
                        const double* xi = xs + G::XPAD + i;
                        for (int k = -nb; k <= nb; ++k) acc = fma(-apad[OFF + k], xi[k], acc);
                        rmax = nanmax(rmax, fabs(acc));
                        xmax = nanmax(xmax, fabs(xi[0]));
                        bmax = nanmax(bmax, fabs(bi));
                    }
                    rs[i] = acc;
                }
                rmax = warp_nanmax(rmax);
                xmax = warp_nanmax(xmax);
                bmax = warp_nanmax(bmax);
                ok = rmax <= 16.0 * 2.220446049250313e-16 * fma(amax, xmax, bmax);
                // give up when the budget is spent or the residual no longer halves (NaN included): the pivoted LU takes over
                if (ok || it == kBldlRefineSteps || !(rmax <= 0.5 * r_prev)) break;
                r_prev = rmax;
                __syncwarp();
                bldl_forward<T>(lane, n_panels, U, rs);
                __syncwarp();
                bldl_backward<T>(lane, n_panels, U, xs, true);
                __syncwarp();
            }
        }
        if (ok) {
            for (int i = lane; i < n; i += 32) xo[i] = xs[G::XPAD + i];
            if (lane == 0) {
                if (p.info) p.info[s] = 0;
                if (indef && p.refined_count) atomicAdd(p.refined_count, 1);
            }
        } else if (lane == 0) {
            p.fail_list[atomicAdd(p.fail_count, 1)] = (int32_t)s;
        }
        __syncwarp();
        discard_l2(U, G::ustride(n) * 8, lane, 32);
        __syncwarp();
        double* t = bs;
        bs = bs_next;
        bs_next = t;
    }
    cp_async_wait_all();
}

}  // namespace frx_k4
