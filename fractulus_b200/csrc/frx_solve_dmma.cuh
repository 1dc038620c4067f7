// frx_solve_dmma.cuh -- K4 `frx_k4_band_dmma`: blocked band LU with partial pivoting, ONE WARP PER SYSTEM, the
// trailing update on FP64 tensor cores (mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4).
//
// PARITY UNPINNED like everything in K4 (the reference has no solver; oracle = oracle/assemble.py + LAPACK).
//
// Why blocked: the column-at-a-time kernel (frx_k4_band_stream) spends two shared-memory loads and a store per DFMA
// and one dependent pivot search per column; ncu: FP64 pipe 12 %, stalled on the shared-memory round trips
// (profiles/r01_k4_band_stream_final_ncu_full.txt).  Here the elimination advances in panels of 8 columns:
//
//   A  panel      lane = row slot: the 8 panel entries of the lane's row (+ its right-hand side as a 9th column) sit
//                 in registers; per column one pivot search (3 redux), one broadcast of the pivot row's remaining
//                 panel entries (shuffles), <= 8 FMAs.  Rows never move (implicit pivoting).  The pivot row retires
//                 as a row of U (reciprocal pivot on the diagonal): its trailing window entries go into registers
//                 (lane = trailing column), its multipliers w.r.t. the earlier pivots of the panel into L11, and the
//                 matrix row that enters the band at this column (k + nb + 1) takes its slot at once, built from the
//                 2nb+1 band coefficients -- so nb + 1 <= 32 row slots are enough for every band this kernel takes.
//   B  U12        lane = trailing column: U12 = L11^-1 * (retired rows), 28 FMAs per column in registers, stored as the
//                 B operand (and to the U scratch the back substitution reads).
//   C  update     every 8x8 tile of the window: C += (-L21) * U12 = two DMMA m8n8k4; the multipliers phase A left in
//                 its registers go to shared memory once (`PB`) and are the A fragments as they lie, U12 the B
//                 fragments: one LDS.128 + 2 DMMA + one STS.128 per tile, loads one column tile ahead.
//
// Window: (nb + 1) row slots x (2nb + 8) columns, row slots fixed, column tiles of 8 in a ring (tile of window
// column jj of panel p: (p + jj/8) mod CT), plus a dense 8-column copy of the current panel (`PB`).  The column tile a
// panel frees is zeroed and becomes the right-most one.  Geometry <RT, CT> = row / column tiles:
//   nb <= 7: <1,3>   nb <= 15: <2,5>   nb <= 23: <3,7>   nb <= 31: <4,9>      (CT odd: row stride = 8 mod 16 doubles,
// so the C-fragment LDS.128/STS.128 are bank-conflict free).
// Back substitution: blocks of 8 rows; the sums over the columns right of the block are taken lane-per-column and
// transpose-reduced (9 shuffles for 8 sums), the 8x8 triangle is solved across lane groups; the next block's U rows
// are prefetched while the current one is solved.
// The index bookkeeping was prototyped in tools/proto_band_lu.py (NumPy, against numpy.linalg.solve).
#pragma once

namespace frx_k4 {

template <int RT_, int CT_>
struct BandGeo {
    static constexpr int RT = RT_, CT = CT_;
    static constexpr int S = RT * 8;                 // row slots (<= 32: one per lane)
    static constexpr int WC = CT * 8;                // window columns
    static constexpr int RS = WC;                    // row stride of W (doubles); CT odd -> RS = 8 (mod 16)
    static constexpr int TW = (CT - 1) * 8;          // trailing columns
    static constexpr int TPL = (TW + 31) / 32;       // trailing columns per lane
    static constexpr int WCL = (WC + 31) / 32;       // window columns per lane
    static constexpr int TS = TW + 2;                // stride of the B-operand staging (pairs); = 2 (mod 8)
    static constexpr int NB_MAX = (S - 1) < (WC - 8) / 2 ? (S - 1) : (WC - 8) / 2;
    static constexpr int APAD = WC + 32;             // band coefficients a[0 .. 2nb], zeros up to here
    static_assert(CT % 2 == 1, "CT must be odd");
    static_assert(S <= 32, "one row slot per lane");
    __host__ __device__ static inline int smem_doubles(int n) {
        return S * RS + S * 8 + 8 * TS + 64 + APAD + ((n + 1) & ~1);
    }
    __host__ __device__ static inline int smem_doubles2(int n) {      // frx_k4_band_dmma2: REC (80) instead of L11 (64), + PS
        return S * RS + S * 8 + 8 * TS + 80 + APAD + ((n + 1) & ~1) + 4;
    }
};

__device__ inline void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// The U rows of a finished system are dead: tell L2 not to write the scratch lines back (they would otherwise reach DRAM
// when they are evicted: 70-90 MB per launch against 11 MB of algorithmic bytes, profiles/r02_k4_band_dmma_v2_ncu_full.txt).
__device__ __forceinline__ void discard_l2(const double* base, int64_t bytes, int tid, int nthreads) {
    const char* b = reinterpret_cast<const char*>(base);
    for (int64_t off = (int64_t)tid * 128; off + 128 <= bytes; off += (int64_t)nthreads * 128)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(b + off) : "memory");
}

struct BandParams {
    int64_t n_sys;
    int n;
    const double* rcw;       // Riesz-Caputo weights, system s at rcw[wptr[s] .. wptr[s+1])
    const int64_t* wptr;
    const double* h;
    const double* rhs;
    int64_t ldb;
    double* x;
    int64_t ldx;
    int32_t* info;
    double* uscratch;        // [grid][ustride_cta]
    int64_t ustride_cta;     // doubles per CTA: n * (2 * nb_max + 1) rounded up to a multiple of 16 (128 bytes)
    const int32_t* order;    // the systems of this launch (one band-width bin)
    int nb_max;              // widest band (res + 1) among them
    const int32_t* n_sys_dev;   // if not NULL: the number of systems in `order` is read from here (a fail list of frx_k4_band_ldl)
    int32_t* fail_list;      // frx_k4_band_ldl: systems that turned out indefinite, appended at fail_list[atomicAdd(fail_count, 1)]
    int32_t* fail_count;
    int32_t* refined_count;  // frx_k4_band_bldl: indefinite systems accepted on their residual after iterative refinement (may be NULL)
    int32_t* queue;          // frx_k4_band_bldl: work counter of this launch (zero at launch), see the kernel
};

// Back substitution in blocks of 8 rows, one warp.  xs[k] = transformed rhs on entry, the solution on exit;
// U[k][0] = 1 / pivot; U[k][c], c >= 1, multiplies x_{k+c}.
template <class G>
__device__ __forceinline__ void band_back_substitution(const int lane, const int n, const int n_panels, const int us,
                                                       const double* __restrict__ U, double* xs) {
    constexpr int TW = G::TW, TPL = G::TPL;
    constexpr unsigned FULL = 0xffffffffu;
    const int gq = lane >> 2;
    auto load_block = [&](int kb, double (&uu)[TPL][8], double (&dd)[8]) {
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj) {
            const int t = lane + 32 * jj;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int rel = 8 - q + t;
                uu[jj][q] = (kb >= 0 && t < TW && kb + 8 + t < n && rel < us) ? __ldcg(U + (int64_t)(kb + q) * us + rel) : 0.0;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            dd[cc] = (kb >= 0 && cc >= gq && cc - gq < us && kb + cc < n) ? __ldcg(U + (int64_t)(kb + gq) * us + cc - gq) : 0.0;
    };
    double un[TPL][8], dn[8];
    load_block((n_panels - 1) * 8, un, dn);
    for (int kb = (n_panels - 1) * 8; kb >= 0; kb -= 8) {
        double uc[TPL][8], d[8];
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj)
#pragma unroll
            for (int q = 0; q < 8; ++q) uc[jj][q] = un[jj][q];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) d[cc] = dn[cc];
        load_block(kb - 8, un, dn);
        double sacc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sacc[q] = 0.0;
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj) {
            const int t = lane + 32 * jj, j = kb + 8 + t;
            const double xj = (t < TW && j < n) ? xs[j] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) sacc[q] = fma(uc[jj][q], xj, sacc[q]);
        }
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
        double rhsq = (kb + gq < n) ? xs[kb + gq] - v1 : 0.0;
        double rinvq = 0.0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) rinvq = cc == gq ? d[cc] : rinvq;
        __syncwarp();
#pragma unroll
        for (int qq = 7; qq >= 0; --qq) {
            const double xv = __shfl_sync(FULL, rhsq * rinvq, qq * 4);
            if (gq < qq) rhsq = fma(-d[qq], xv, rhsq);
            if (lane == qq * 4 && kb + qq < n) xs[kb + qq] = xv;
        }
        __syncwarp();
    }}

template <int RT, int CT>
__global__ void __launch_bounds__(32) frx_k4_band_dmma(const BandParams p) {
    using G = BandGeo<RT, CT>;
    constexpr int S = G::S, WC = G::WC, RS = G::RS, TW = G::TW, TPL = G::TPL, TS = G::TS, WCL = G::WCL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, g = lane >> 2, cp = lane & 3, n = p.n;
    double* W = sm;                         // W[slot * RS + tile * 8 + col]
    double* PB = W + S * RS;                // PB[slot * 8 + c]: current panel, then the multipliers (negated)
    double* Ub = PB + S * 8;                // Ub[(m * TS + t) * 2 + {0,1}] = U12[2m + {0,1}][t]
    double* L11 = Ub + 8 * TS;              // L11[q * 8 + c], c < q: multiplier (negated) of pivot row q w.r.t. pivot c
    double* apad = L11 + 64;                // apad[e] = a_{e - nb} for e <= 2nb, 0 beyond
    double* xs = apad + G::APAD;            // rhs, then transformed rhs, then the solution
    double* U = p.uscratch + (int64_t)blockIdx.x * p.ustride_cta;
    const int n_panels = (n + 7) >> 3;
    const int64_t n_sys = p.n_sys_dev ? (int64_t)*p.n_sys_dev : p.n_sys;
    for (int64_t si = blockIdx.x; si < n_sys; si += gridDim.x) {
        const int64_t s = p.order[si];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;
        const int us = 2 * nb + 1;                           // length (and stride) of a U row
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const double* rhs = p.rhs + s * p.ldb;
        double* xo = p.x + s * p.ldx;
        __syncwarp();
        for (int e = lane; e < G::APAD; e += 32) apad[e] = 0.0;
        for (int i = lane; i < n; i += 32) xs[i] = rhs[i];
        __syncwarp();
        for (int k = -nb_a + lane; k <= nb_a; k += 32) {     // band coefficients: terms and order of the dict-merging expansion
            double acc = 0.0;
            bool any = false;
            auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
            if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));
            if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));
            if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));
            if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));
            if (k >= -nb && k <= nb) apad[k + nb] = acc;
        }
        __syncwarp();
        // ---- initial window: rows 0 .. min(n, nb + 1) - 1 in the slots of the same number --------------------------
        const int rows0 = n < nb + 1 ? n : nb + 1;
        int label = lane < rows0 ? lane : -1;
        double bval = lane < rows0 ? xs[lane] : 0.0;
        for (int r0 = 0; r0 < rows0; ++r0) {
#pragma unroll
            for (int i = 0; i < WCL; ++i) {
                const int jj = lane + 32 * i;
                if (jj < WC) {
                    const double v = jj < n ? apad[jj - r0 + nb] : 0.0;     // d = jj - r0 >= -nb; 0 beyond +nb
                    W[r0 * RS + jj] = v;
                    if (i == 0 && lane < 8) PB[r0 * 8 + lane] = v;
                }
            }
        }
        int info = 0, pm = 0;                                // pm = panel index mod CT
        double areg[7];                                      // a_{-nb} .. a_{-nb+6}: the panel entries of a row entering the band
#pragma unroll
        for (int e = 0; e < 7; ++e) areg[e] = apad[e];
        __syncwarp();
        for (int pnl = 0; pnl < n_panels; ++pnl) {
            const int k0 = pnl * 8;
            // shared-memory offsets of this lane's trailing columns (trailing column t = window column 8 + t) in the ring
            int colidx[TPL];
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int t = lane + 32 * i;
                int tile = pm + 1 + (t >> 3);
                if (tile >= CT) tile -= CT;
                colidx[i] = tile * 8 + (t & 7);
            }
            // ================= phase A: panel factorisation, lane = row slot =================================
            double r[9], u[TPL][8];
            if (lane < S) {
                const double2* src = reinterpret_cast<const double2*>(PB + lane * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v = src[q];
                    r[2 * q] = v.x;
                    r[2 * q + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) r[q] = 0.0;
            }
            r[8] = bval;
#pragma unroll
            for (int i = 0; i < TPL; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) u[i][q] = 0.0;
            int ps_valid = 0;                                // bit c: column k0 + c exists
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (k0 + c < n) {                            // warp-uniform
                    const int k = k0 + c;
                    ps_valid |= 1 << c;
                    const bool act = label >= 0;
                    // |v| >= 0 orders like its bit pattern: two 32-bit reductions; ties -> lowest row label (LAPACK: first maximum)
                    const unsigned long long bits = act ? (unsigned long long)__double_as_longlong(fabs(r[c])) : 0ull;
                    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                    const unsigned mh = __reduce_max_sync(FULL, hi);
                    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
                    const int code = __reduce_min_sync(FULL, (act && hi == mh && lo == ml) ? ((label << 5) | lane) : 0x7fffffff);
                    const int P = code & 31, blog = code >> 5;
                    double prow[9];
#pragma unroll
                    for (int cc = c; cc < 9; ++cc) prow[cc] = __shfl_sync(FULL, r[cc], P);
                    const double pval = prow[c];
                    if (pval == 0.0 && info == 0) info = k + 1;
                    const double rinv = 1.0 / pval;
                    const bool is_pivot = lane == P;
                    if (act && !is_pivot) {
                        const double l = r[c] * rinv;
                        r[c] = l;
#pragma unroll
                        for (int cc = c + 1; cc < 9; ++cc) r[cc] = fma(-l, prow[cc], r[cc]);
                        if (label == k) label = blog;          // the interchange, on labels only
                    }
                    // the pivot row leaves the window: trailing entries to the registers of their columns' lanes,
                    // multipliers w.r.t. the earlier pivots to L11, panel entries to row k of U
                    // (every lane OWNS its trailing columns: it reads the retiring row's entry and writes the entering row's
                    // coefficient in its place itself, so no warp barrier is needed between the two, nor before the next pivot)
                    double* wrow = W + P * RS;
                    const int inew = k + nb + 1;
#pragma unroll
                    for (int i = 0; i < TPL; ++i) {
                        const int t = lane + 32 * i;
                        if (t < TW) {
                            u[i][c] = wrow[colidx[i]];
                            if (inew < n) wrow[colidx[i]] = (k0 + 8 + t < n) ? apad[t + 7 - c] : 0.0;   // window column 8 + t of row inew
                        } else {
                            u[i][c] = 0.0;
                        }
                    }
                    if (is_pivot) {
#pragma unroll
                        for (int cc = 0; cc < c; ++cc) L11[c * 8 + cc] = -r[cc];
                        double* urow = U + (int64_t)k * us;
                        urow[0] = rinv;                        // U keeps the reciprocal pivot on its diagonal
#pragma unroll
                        for (int cc = c + 1; cc < 8; ++cc)
                            if (cc - c < us && k0 + cc < n) urow[cc - c] = r[cc];
                        xs[k] = r[8];                          // transformed right-hand side of row k of U
                        // ... and the matrix row that enters the band at this column takes its slot
                        if (inew < n) {
#pragma unroll
                            for (int cc = 0; cc < 8; ++cc) r[cc] = (cc > c && k0 + cc < n) ? areg[cc > c ? cc - c - 1 : 0] : 0.0;
                            r[8] = xs[inew];
                            label = inew;
                        } else {
                            label = -1;
#pragma unroll
                            for (int cc = 0; cc < 9; ++cc) r[cc] = 0.0;
                        }
                    }
                }
            }
            __syncwarp();
            // multipliers (negated) to PB: the A operand of the trailing update
            if (lane < S) {
                const bool alive = label >= 0;
                double m[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) m[cc] = (alive && k0 + cc < n) ? -r[cc] : 0.0;
                double2* dst = reinterpret_cast<double2*>(PB + lane * 8);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) dst[qq] = make_double2(m[2 * qq], m[2 * qq + 1]);
            }
            bval = r[8];
            // ================= phase B: U12 = L11^-1 (retired rows), lane = trailing column ====================
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                if (ps_valid & (1 << q)) {
                    const double2* mrow = reinterpret_cast<const double2*>(L11 + q * 8);
                    double mq[8];
#pragma unroll
                    for (int qq = 0; qq < (q + 1) / 2; ++qq) {
                        const double2 v = mrow[qq];
                        mq[2 * qq] = v.x;
                        mq[2 * qq + 1] = v.y;
                    }
#pragma unroll
                    for (int i = 0; i < TPL; ++i)
#pragma unroll
                        for (int c = 0; c < q; ++c) u[i][q] = fma(mq[c], u[i][c], u[i][q]);
                }
            }
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int t = lane + 32 * i;
                if (t < TW) {
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        *reinterpret_cast<double2*>(Ub + (m * TS + t) * 2) = make_double2(u[i][2 * m], u[i][2 * m + 1]);
                    if (k0 + 8 + t < n) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int rel = 8 - q + t;
                            if ((ps_valid & (1 << q)) && rel < us) U[(int64_t)(k0 + q) * us + rel] = u[i][q];
                        }
                    }
                }
            }
            __syncwarp();
            // ================= phase C: trailing update on the tensor cores ====================================
            double2 am[RT];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) am[rt] = *reinterpret_cast<const double2*>(PB + (8 * rt + g) * 8 + 2 * cp);
            __syncwarp();                                   // PB may now take the next panel
            {
                double2 cf[RT], cfn[RT];
                int tile = pm + 1 == CT ? 0 : pm + 1;
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) cf[rt] = *reinterpret_cast<const double2*>(W + (8 * rt + g) * RS + tile * 8 + 2 * cp);
#pragma unroll
                for (int ct = 0; ct < CT - 1; ++ct) {
                    const int tile_next = tile + 1 == CT ? 0 : tile + 1;
                    if (ct + 1 < CT - 1) {
#pragma unroll
                        for (int rt = 0; rt < RT; ++rt)
                            cfn[rt] = *reinterpret_cast<const double2*>(W + (8 * rt + g) * RS + tile_next * 8 + 2 * cp);
                    }
                    const double2 bb = *reinterpret_cast<const double2*>(Ub + (cp * TS + 8 * ct + g) * 2);
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        dmma884(cf[rt].x, cf[rt].y, am[rt].x, bb.x);
                        dmma884(cf[rt].x, cf[rt].y, am[rt].y, bb.y);
                    }
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        if (ct == 0) *reinterpret_cast<double2*>(PB + (8 * rt + g) * 8 + 2 * cp) = cf[rt];   // the next panel
                        else *reinterpret_cast<double2*>(W + (8 * rt + g) * RS + tile * 8 + 2 * cp) = cf[rt];
                        cf[rt] = cfn[rt];
                    }
                    tile = tile_next;
                }
            }
#pragma unroll
            for (int rt = 0; rt < RT; ++rt)                 // the freed column tile: zero, it becomes the right-most one
                *reinterpret_cast<double2*>(W + (8 * rt + g) * RS + pm * 8 + 2 * cp) = make_double2(0.0, 0.0);
            pm = pm + 1 == CT ? 0 : pm + 1;
            __syncwarp();
        }
        band_back_substitution<G>(lane, n, n_panels, us, U, xs);
        for (int i = lane; i < n; i += 32) xo[i] = xs[i];
        if (lane == 0 && p.info) p.info[s] = info;
        __syncwarp();
        discard_l2(U, (int64_t)n * us * 8, lane, 32);
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Narrow bands: SEVERAL systems per warp.  A band with nb <= 7 (<= 15) needs 8 (16) row slots, so one warp holds 4 (2)
// systems -- lane groups of GS = 8 (16) lanes, each group its own system, window and U scratch.  The pivot chain of a
// column (three reductions, broadcast, reciprocal, update) is issued ONCE for all groups of the warp (reductions and
// shuffles confined to the group), the tensor-core phase walks through the groups with the whole warp.  These bins are
// instruction-issue bound with one system per warp (issue slots 54-59 %, profiles/r02_k4_band_dmma_final_ncu_full.txt).
// All systems of a launch share n, so the panel loop is uniform; band widths may differ from group to group.
// ----------------------------------------------------------------------------------------------------------------
template <class G, int GS>
__device__ __forceinline__ void band_back_substitution_group(const int gl, const int n, const int n_panels, const int us,
                                                             const double* __restrict__ U, double* xs) {
    constexpr int TW = G::TW, TPL = (TW + GS - 1) / GS;
    constexpr unsigned FULL = 0xffffffffu;
    const int gq = gl & 7;                           // row of the block this lane solves (GS = 16: lanes 8..15 duplicate 0..7)
    auto load_block = [&](int kb, double (&uu)[TPL][8], double (&dd)[8]) {
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj) {
            const int t = gl + GS * jj;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int rel = 8 - q + t;
                uu[jj][q] = (kb >= 0 && t < TW && kb + 8 + t < n && rel < us) ? __ldcg(U + (int64_t)(kb + q) * us + rel) : 0.0;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            dd[cc] = (kb >= 0 && cc >= gq && cc - gq < us && kb + cc < n) ? __ldcg(U + (int64_t)(kb + gq) * us + cc - gq) : 0.0;
    };
    double un[TPL][8], dn[8];
    load_block((n_panels - 1) * 8, un, dn);
    for (int kb = (n_panels - 1) * 8; kb >= 0; kb -= 8) {
        double uc[TPL][8], d[8];
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj)
#pragma unroll
            for (int q = 0; q < 8; ++q) uc[jj][q] = un[jj][q];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) d[cc] = dn[cc];
        load_block(kb - 8, un, dn);
        double sacc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sacc[q] = 0.0;
#pragma unroll
        for (int jj = 0; jj < TPL; ++jj) {
            const int t = gl + GS * jj, j = kb + 8 + t;
            const double xj = (t < TW && j < n) ? xs[j] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) sacc[q] = fma(uc[jj][q], xj, sacc[q]);
        }
        if (GS == 16) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sacc[q] += __shfl_xor_sync(FULL, sacc[q], 8);
        }
        // transpose-reduce over 8 lanes: lane gq ends with the sum of row kb + gq
        double v4[4], v2[2], v1;
        {
            const bool up = gl & 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double keep = up ? sacc[4 + i] : sacc[i], send = up ? sacc[i] : sacc[4 + i];
                v4[i] = keep + __shfl_xor_sync(FULL, send, 4);
            }
            const bool up2 = gl & 2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double keep = up2 ? v4[2 + i] : v4[i], send = up2 ? v4[i] : v4[2 + i];
                v2[i] = keep + __shfl_xor_sync(FULL, send, 2);
            }
            const bool up1 = gl & 1;
            const double keep = up1 ? v2[1] : v2[0], send = up1 ? v2[0] : v2[1];
            v1 = keep + __shfl_xor_sync(FULL, send, 1);
        }
        double rhsq = (kb + gq < n) ? xs[kb + gq] - v1 : 0.0;
        double rinvq = 0.0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) rinvq = cc == gq ? d[cc] : rinvq;
        __syncwarp();
#pragma unroll
        for (int qq = 7; qq >= 0; --qq) {
            const double xv = __shfl_sync(FULL, rhsq * rinvq, qq, GS);
            if (gq < qq) rhsq = fma(-d[qq], xv, rhsq);
            if (gl == qq && kb + qq < n) xs[kb + qq] = xv;
        }
        __syncwarp();
    }
}

template <int RT, int CT, int GS>
__global__ void __launch_bounds__(32) frx_k4_band_dmma_packed(const BandParams p) {
    using G = BandGeo<RT, CT>;
    constexpr int S = G::S, WC = G::WC, RS = G::RS, TW = G::TW, TS = G::TS;
    constexpr int NG = 32 / GS, TPL = (TW + GS - 1) / GS, WCL = (WC + GS - 1) / GS;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(S <= GS && (GS == 8 || GS == 16), "one row slot per lane of a group");
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, g = lane >> 2, cp = lane & 3, n = p.n;
    const int gl = lane & (GS - 1), gi = lane / GS;
    const unsigned gmask = ((1u << GS) - 1u) << (gi * GS);
    const int gsm = G::smem_doubles(n);                      // doubles of shared memory per system
    double* W = sm + gi * gsm;
    double* PB = W + S * RS;
    double* Ub = PB + S * 8;
    double* L11 = Ub + 8 * TS;
    double* apad = L11 + 64;
    double* xs = apad + G::APAD;
    double* U = p.uscratch + ((int64_t)blockIdx.x * NG + gi) * p.ustride_cta;
    const int n_panels = (n + 7) >> 3;
    const int64_t n_sys = p.n_sys_dev ? (int64_t)*p.n_sys_dev : p.n_sys;
    for (int64_t base = (int64_t)blockIdx.x * NG; base < n_sys; base += (int64_t)gridDim.x * NG) {
        // a group without a system of its own repeats the launch's last one (same values written twice: harmless)
        const int64_t sidx = base + gi < n_sys ? base + gi : n_sys - 1;
        const int64_t s = p.order[sidx];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;
        const int us = 2 * nb + 1;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const double* rhs = p.rhs + s * p.ldb;
        double* xo = p.x + s * p.ldx;
        __syncwarp();
        for (int e = gl; e < G::APAD; e += GS) apad[e] = 0.0;
        for (int i = gl; i < n; i += GS) xs[i] = rhs[i];
        __syncwarp();
        for (int k = -nb_a + gl; k <= nb_a; k += GS) {       // band coefficients: terms and order of the dict-merging expansion
            double acc = 0.0;
            bool any = false;
            auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
            if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));
            if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));
            if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));
            if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));
            if (k >= -nb && k <= nb) apad[k + nb] = acc;
        }
        __syncwarp();
        const int rows0 = n < nb + 1 ? n : nb + 1;
        int label = gl < rows0 ? gl : -1;
        double bval = gl < rows0 ? xs[gl] : 0.0;
#pragma unroll 1
        for (int r0 = 0; r0 < S; ++r0) {
            if (r0 < rows0) {
#pragma unroll
                for (int i = 0; i < WCL; ++i) {
                    const int jj = gl + GS * i;
                    if (jj < WC) {
                        const double v = jj < n ? apad[jj - r0 + nb] : 0.0;
                        W[r0 * RS + jj] = v;
                        if (jj < 8) PB[r0 * 8 + jj] = v;
                    }
                }
            }
        }
        int info = 0, pm = 0;
        double areg[7];
#pragma unroll
        for (int e = 0; e < 7; ++e) areg[e] = apad[e];
        __syncwarp();
        for (int pnl = 0; pnl < n_panels; ++pnl) {
            const int k0 = pnl * 8;
            int colidx[TPL];
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int t = gl + GS * i;
                int tile = pm + 1 + (t >> 3);
                if (tile >= CT) tile -= CT;
                colidx[i] = tile * 8 + (t & 7);
            }
            // ================= phase A: the pivot chain, issued once for all groups of the warp ==================
            double r[9], u[TPL][8];
            if (gl < S) {
                const double2* src = reinterpret_cast<const double2*>(PB + gl * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v = src[q];
                    r[2 * q] = v.x;
                    r[2 * q + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) r[q] = 0.0;
            }
            r[8] = bval;
#pragma unroll
            for (int i = 0; i < TPL; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) u[i][q] = 0.0;
            int ps_valid = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (k0 + c < n) {                            // warp-uniform: every system of the launch has n unknowns
                    const int k = k0 + c;
                    ps_valid |= 1 << c;
                    const bool act = label >= 0;
                    const unsigned long long bits = act ? (unsigned long long)__double_as_longlong(fabs(r[c])) : 0ull;
                    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                    const unsigned mh = __reduce_max_sync(gmask, hi);
                    const unsigned ml = __reduce_max_sync(gmask, hi == mh ? lo : 0u);
                    const int code = __reduce_min_sync(gmask, (act && hi == mh && lo == ml) ? ((label << 5) | gl) : 0x7fffffff);
                    const int P = code & 31, blog = code >> 5;       // P: lane within the group
                    double prow[9];
#pragma unroll
                    for (int cc = c; cc < 9; ++cc) prow[cc] = __shfl_sync(FULL, r[cc], P, GS);
                    const double pval = prow[c];
                    if (pval == 0.0 && info == 0) info = k + 1;
                    const double rinv = 1.0 / pval;
                    const bool is_pivot = gl == P;
                    if (act && !is_pivot) {
                        const double l = r[c] * rinv;
                        r[c] = l;
#pragma unroll
                        for (int cc = c + 1; cc < 9; ++cc) r[cc] = fma(-l, prow[cc], r[cc]);
                        if (label == k) label = blog;
                    }
                    double* wrow = W + P * RS;
                    const int inew = k + nb + 1;
#pragma unroll
                    for (int i = 0; i < TPL; ++i) {
                        const int t = gl + GS * i;
                        if (t < TW) {
                            u[i][c] = wrow[colidx[i]];
                            if (inew < n) wrow[colidx[i]] = (k0 + 8 + t < n) ? apad[t + 7 - c] : 0.0;
                        } else {
                            u[i][c] = 0.0;
                        }
                    }
                    if (is_pivot) {
#pragma unroll
                        for (int cc = 0; cc < c; ++cc) L11[c * 8 + cc] = -r[cc];
                        double* urow = U + (int64_t)k * us;
                        urow[0] = rinv;
#pragma unroll
                        for (int cc = c + 1; cc < 8; ++cc)
                            if (cc - c < us && k0 + cc < n) urow[cc - c] = r[cc];
                        xs[k] = r[8];
                        if (inew < n) {
#pragma unroll
                            for (int cc = 0; cc < 8; ++cc) r[cc] = (cc > c && k0 + cc < n) ? areg[cc > c ? cc - c - 1 : 0] : 0.0;
                            r[8] = xs[inew];
                            label = inew;
                        } else {
                            label = -1;
#pragma unroll
                            for (int cc = 0; cc < 9; ++cc) r[cc] = 0.0;
                        }
                    }
                }
            }
            __syncwarp();
            if (gl < S) {
                const bool alive = label >= 0;
                double m[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) m[cc] = (alive && k0 + cc < n) ? -r[cc] : 0.0;
                double2* dst = reinterpret_cast<double2*>(PB + gl * 8);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) dst[qq] = make_double2(m[2 * qq], m[2 * qq + 1]);
            }
            bval = r[8];
            // ================= phase B: U12 = L11^-1 (retired rows), lane of the group = trailing column =========
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                if (ps_valid & (1 << q)) {
                    const double2* mrow = reinterpret_cast<const double2*>(L11 + q * 8);
                    double mq[8];
#pragma unroll
                    for (int qq = 0; qq < (q + 1) / 2; ++qq) {
                        const double2 v = mrow[qq];
                        mq[2 * qq] = v.x;
                        mq[2 * qq + 1] = v.y;
                    }
#pragma unroll
                    for (int i = 0; i < TPL; ++i)
#pragma unroll
                        for (int c = 0; c < q; ++c) u[i][q] = fma(mq[c], u[i][c], u[i][q]);
                }
            }
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int t = gl + GS * i;
                if (t < TW) {
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        *reinterpret_cast<double2*>(Ub + (m * TS + t) * 2) = make_double2(u[i][2 * m], u[i][2 * m + 1]);
                    if (k0 + 8 + t < n) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int rel = 8 - q + t;
                            if ((ps_valid & (1 << q)) && rel < us) U[(int64_t)(k0 + q) * us + rel] = u[i][q];
                        }
                    }
                }
            }
            __syncwarp();
            // ================= phase C: trailing update, the whole warp on one group after the other ===========
#pragma unroll
            for (int g2 = 0; g2 < NG; ++g2) {
                double* W2 = sm + g2 * gsm;
                double* PB2 = W2 + S * RS;
                const double* Ub2 = PB2 + S * 8;
                double2 am[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) am[rt] = *reinterpret_cast<const double2*>(PB2 + (8 * rt + g) * 8 + 2 * cp);
                __syncwarp();                               // PB of this group may now take its next panel
                int tile = pm + 1 == CT ? 0 : pm + 1;
#pragma unroll
                for (int ct = 0; ct < CT - 1; ++ct) {
                    const double2 bb = *reinterpret_cast<const double2*>(Ub2 + (cp * TS + 8 * ct + g) * 2);
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        double2* cptr = reinterpret_cast<double2*>(W2 + (8 * rt + g) * RS + tile * 8 + 2 * cp);
                        double2 cf = *cptr;
                        dmma884(cf.x, cf.y, am[rt].x, bb.x);
                        dmma884(cf.x, cf.y, am[rt].y, bb.y);
                        if (ct == 0) *reinterpret_cast<double2*>(PB2 + (8 * rt + g) * 8 + 2 * cp) = cf;
                        else *cptr = cf;
                    }
                    tile = tile + 1 == CT ? 0 : tile + 1;
                }
#pragma unroll
                for (int rt = 0; rt < RT; ++rt)             // the freed column tile: zero, it becomes the right-most one
                    *reinterpret_cast<double2*>(W2 + (8 * rt + g) * RS + pm * 8 + 2 * cp) = make_double2(0.0, 0.0);
            }
            pm = pm + 1 == CT ? 0 : pm + 1;
            __syncwarp();
        }
        band_back_substitution_group<G, GS>(gl, n, n_panels, us, U, xs);
        for (int i = gl; i < n; i += GS) xo[i] = xs[i];
        if (gl == 0 && p.info) p.info[s] = info;
        __syncwarp();
        discard_l2(U, (int64_t)n * us * 8, gl, GS);
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Two warps per system (one CTA of 64 threads): the wide-band bins.  Shared memory, not registers, limits how many systems
// an SM holds (27 KB each at nb = 31 -> 8), and one warp running the dependent chain of a panel issues ~0.2 instructions
// per cycle (profiles/r02_k4_band_dmma_*): the SM idles.  Here warp 0 keeps the chain -- and nothing else: per column it
// publishes the retired row's registers (REC) and the pivot slot (PS) in shared memory -- and everything that is
// per-COLUMN work is done by all 64 threads after a block barrier: every thread owns ONE trailing column of the window
// (lifts the 8 retired rows' entries out, drops the entering rows' coefficients in, forms its column of U12, stores it to
// the U scratch), the 8x8 panel part of U goes out one element per thread, and the two warps split the column tiles of the
// tensor-core update.  Four block barriers per panel.
// ----------------------------------------------------------------------------------------------------------------
template <int RT, int CT>
__global__ void __launch_bounds__(64, 8) frx_k4_band_dmma2(const BandParams p) {
    using G = BandGeo<RT, CT>;
    constexpr int S = G::S, WC = G::WC, RS = G::RS, TW = G::TW, TS = G::TS, WCL = G::WCL;
    constexpr int H = (CT - 1) / 2;                 // column tiles per warp in the update
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(TW <= 64, "one trailing column per thread");
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, cp = lane & 3, n = p.n;
    double* W = sm;
    double* PB = W + S * RS;
    double* Ub = PB + S * 8;
    double* REC = Ub + 8 * TS;              // REC[c * 10 + i]: registers r[0..8] of the row that retired at panel step c
    double* apad = REC + 80;
    double* xs = apad + G::APAD;
    int* PS = reinterpret_cast<int*>(xs + ((n + 1) & ~1));      // PS[c]: pivot slot of step c, -1 beyond the matrix
    double* U = p.uscratch + (int64_t)blockIdx.x * p.ustride_cta;
    const int n_panels = (n + 7) >> 3;
    const int64_t n_sys = p.n_sys_dev ? (int64_t)*p.n_sys_dev : p.n_sys;
    for (int64_t si = blockIdx.x; si < n_sys; si += gridDim.x) {
        const int64_t s = p.order[si];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;
        const int us = 2 * nb + 1;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const double* rhs = p.rhs + s * p.ldb;
        double* xo = p.x + s * p.ldx;
        __syncthreads();
        for (int e = tid; e < G::APAD; e += 64) apad[e] = 0.0;
        for (int i = tid; i < n; i += 64) xs[i] = rhs[i];
        __syncthreads();
        for (int k = -nb_a + tid; k <= nb_a; k += 64) {      // band coefficients: terms and order of the dict-merging expansion
            double acc = 0.0;
            bool any = false;
            auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
            if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));
            if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));
            if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));
            if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));
            if (k >= -nb && k <= nb) apad[k + nb] = acc;
        }
        __syncthreads();
        const int rows0 = n < nb + 1 ? n : nb + 1;
        int label = (warp == 0 && lane < rows0) ? lane : -1;
        double bval = (warp == 0 && lane < rows0) ? xs[lane] : 0.0;
        for (int r0 = warp; r0 < rows0; r0 += 2) {
#pragma unroll
            for (int i = 0; i < WCL; ++i) {
                const int jj = lane + 32 * i;
                if (jj < WC) {
                    const double v = jj < n ? apad[jj - r0 + nb] : 0.0;
                    W[r0 * RS + jj] = v;
                    if (i == 0 && lane < 8) PB[r0 * 8 + lane] = v;
                }
            }
        }
        int info = 0, pm = 0;
        double areg[7];
#pragma unroll
        for (int e = 0; e < 7; ++e) areg[e] = apad[e];
        __syncthreads();
        for (int pnl = 0; pnl < n_panels; ++pnl) {
            const int k0 = pnl * 8;
            const bool interior = k0 + 8 + TW <= n && us >= 8;
            // ================= phase A (warp 0): the dependent chain of the panel, lane = row slot =================
            if (warp == 0) {
                double r[9];
                if (lane < S) {
                    const double2* src = reinterpret_cast<const double2*>(PB + lane * 8);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 v = src[q];
                        r[2 * q] = v.x;
                        r[2 * q + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) r[q] = 0.0;
                }
                r[8] = bval;
                int centered = -1;                           // panel step at which this lane's row entered the window
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (interior || k0 + c < n) {            // warp-uniform
                        const int k = k0 + c;
                        const bool act = label >= 0;
                        const unsigned long long bits = act ? (unsigned long long)__double_as_longlong(fabs(r[c])) : 0ull;
                        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                        const unsigned mh = __reduce_max_sync(FULL, hi);
                        const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
                        const int code = __reduce_min_sync(FULL, (act && hi == mh && lo == ml) ? ((label << 5) | lane) : 0x7fffffff);
                        const int P = code & 31, blog = code >> 5;
                        double prow[9];
#pragma unroll
                        for (int cc = c; cc < 9; ++cc) prow[cc] = __shfl_sync(FULL, r[cc], P);
                        const double pval = prow[c];
                        if (pval == 0.0 && info == 0) info = k + 1;
                        const double rinv = 1.0 / pval;
                        if (lane == P) {
                            // the retired row's registers as they are (its multipliers w.r.t. earlier pivots, reciprocal pivot, panel part of
                            // the U row, transformed rhs): the consumers know which entries belong to this row
                            r[c] = rinv;
                            double2* rec = reinterpret_cast<double2*>(REC + c * 10);
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) rec[qq] = make_double2(r[2 * qq], r[2 * qq + 1]);
                            REC[c * 10 + 8] = r[8];
                            xs[k] = r[8];
                            const int inew = k + nb + 1;
                            if (inew < n) {
#pragma unroll
                                for (int cc = c + 1; cc < 8; ++cc) r[cc] = (interior || k0 + cc < n) ? areg[cc - c - 1] : 0.0;
                                r[8] = xs[inew];
                                label = inew;
                                centered = c;
                            } else {
                                label = -1;
                            }
                        } else if (act) {
                            const double l = r[c] * rinv;
                            r[c] = l;
#pragma unroll
                            for (int cc = c + 1; cc < 9; ++cc) r[cc] = fma(-l, prow[cc], r[cc]);
                            if (label == k) label = blog;
                        }
                        if (lane == 0) PS[c] = P;
                    } else if (lane == 0) {
                        PS[c] = -1;
                    }
                }
                if (lane < S) {
                    const bool alive = label >= 0;
                    double m[8];
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) m[cc] = (alive && cc > centered && (interior || k0 + cc < n)) ? -r[cc] : 0.0;
                    double2* dst = reinterpret_cast<double2*>(PB + lane * 8);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) dst[qq] = make_double2(m[2 * qq], m[2 * qq + 1]);
                }
                bval = r[8];
            }
            __syncthreads();
            // ================= phase B (64 threads): one trailing column per thread ===============================
            {
                int ps[8];
                {
                    const int4 a4 = *reinterpret_cast<const int4*>(PS), b4 = *reinterpret_cast<const int4*>(PS + 4);
                    ps[0] = a4.x; ps[1] = a4.y; ps[2] = a4.z; ps[3] = a4.w;
                    ps[4] = b4.x; ps[5] = b4.y; ps[6] = b4.z; ps[7] = b4.w;
                }
                const int t = tid;
                if (t < TW) {
                    int tile = pm + 1 + (t >> 3);
                    if (tile >= CT) tile -= CT;
                    const int col = tile * 8 + (t & 7);
                    const bool col_in = interior || k0 + 8 + t < n;
                    double u[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        u[c] = 0.0;
                        if (ps[c] >= 0) {
                            double* w = W + ps[c] * RS + col;
                            u[c] = *w;
                            if (interior || k0 + c + nb + 1 < n) *w = col_in ? apad[t + 7 - c] : 0.0;   // window column 8 + t of row k + nb + 1
                        }
                    }
#pragma unroll
                    for (int q = 1; q < 8; ++q) {
                        if (ps[q] >= 0) {
                            int cent = -1;                   // step at which the row that retired at step q had entered this panel
#pragma unroll
                            for (int c1 = 0; c1 < q; ++c1) cent = ps[c1] == ps[q] ? c1 : cent;
                            const double2* mrow = reinterpret_cast<const double2*>(REC + q * 10);
#pragma unroll
                            for (int qq = 0; qq < (q + 1) / 2; ++qq) {
                                const double2 v = mrow[qq];
                                if (2 * qq < q && 2 * qq > cent) u[q] = fma(-v.x, u[2 * qq], u[q]);
                                if (2 * qq + 1 < q && 2 * qq + 1 > cent) u[q] = fma(-v.y, u[2 * qq + 1], u[q]);
                            }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        *reinterpret_cast<double2*>(Ub + (m * TS + t) * 2) = make_double2(u[2 * m], u[2 * m + 1]);
                    if (col_in) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int rel = 8 - q + t;
                            if (ps[q] >= 0 && rel < us) U[(int64_t)(k0 + q) * us + rel] = u[q];
                        }
                    }
                }
                {   // the 8 x 8 panel part of U (reciprocal pivots on the diagonal): one element per thread
                    const int q = tid >> 3, j = tid & 7;
                    if (j >= q && ps[q] >= 0 && j - q < us && (interior || k0 + j < n))
                        U[(int64_t)(k0 + q) * us + (j - q)] = REC[q * 10 + j];
                }
            }
            __syncthreads();
            // ================= phase C (both warps): trailing update on the tensor cores ========================
            double2 am[RT];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) am[rt] = *reinterpret_cast<const double2*>(PB + (8 * rt + g) * 8 + 2 * cp);
            __syncthreads();                                // PB may now take the next panel
            {
                double2 cf[RT], cfn[RT];
                int tile = pm + 1 + warp * H;
                if (tile >= CT) tile -= CT;
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) cf[rt] = *reinterpret_cast<const double2*>(W + (8 * rt + g) * RS + tile * 8 + 2 * cp);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const int ct = warp * H + i;
                    const int tile_next = tile + 1 == CT ? 0 : tile + 1;
                    if (i + 1 < H) {
#pragma unroll
                        for (int rt = 0; rt < RT; ++rt)
                            cfn[rt] = *reinterpret_cast<const double2*>(W + (8 * rt + g) * RS + tile_next * 8 + 2 * cp);
                    }
                    const double2 bb = *reinterpret_cast<const double2*>(Ub + (cp * TS + 8 * ct + g) * 2);
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        dmma884(cf[rt].x, cf[rt].y, am[rt].x, bb.x);
                        dmma884(cf[rt].x, cf[rt].y, am[rt].y, bb.y);
                    }
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        if (ct == 0) *reinterpret_cast<double2*>(PB + (8 * rt + g) * 8 + 2 * cp) = cf[rt];   // the next panel
                        else *reinterpret_cast<double2*>(W + (8 * rt + g) * RS + tile * 8 + 2 * cp) = cf[rt];
                        cf[rt] = cfn[rt];
                    }
                    tile = tile_next;
                }
            }
            if (warp == 1) {
#pragma unroll
                for (int rt = 0; rt < RT; ++rt)             // the freed column tile: zero, it becomes the right-most one
                    *reinterpret_cast<double2*>(W + (8 * rt + g) * RS + pm * 8 + 2 * cp) = make_double2(0.0, 0.0);
            }
            pm = pm + 1 == CT ? 0 : pm + 1;
            __syncthreads();
        }
        if (warp == 0) band_back_substitution<G>(lane, n, n_panels, us, U, xs);
        __syncthreads();
        for (int i = tid; i < n; i += 64) xo[i] = xs[i];
        if (tid == 0 && p.info) p.info[s] = info;
        discard_l2(U, (int64_t)n * us * 8, tid, 64);
    }
}

}  // namespace frx_k4
