// frx_solve.cu -- K4: assemble and solve batches of small fractional boundary-value systems in one kernel.
//
// PARITY UNPINNED: the reference has no assembler and no solver (upstream that is `fdm`'s job); the
// oracle is oracle/assemble.py (reference stencils composed through the fdm shim + numpy.linalg.solve).
//
// System s (n unknowns u_1..u_n at x_k = k*h, h = h[s]): the composed operator
//     Operator(central(h), Operator(create_riesz_caputo_stencil(rule, Settings(alpha, res*h, res)), Operator(central(h))))
// (d/dx o D^alpha_RC o d/dx; stencils: reference equation.py:13-25, central difference:
// reference test/integration/test_equation.py:64) expanded at every interior node.  Boundary rule DEFINED HERE
// (the reference has none): homogeneous Dirichlet data, u = 0 at and beyond both ends.  That makes
//     A[i][j] = a_{j-i},   a_k = ((T1 + T2) + T4) + T3,   |k| <= res + 1,
//     T1 = (-1/h)*(rc_k*(1/h)), T2 = (-1/h)*(rc_{k+1}*(-1/h)), T4 = (1/h)*(rc_{k-1}*(1/h)), T3 = (1/h)*(rc_k*(-1/h))
// with rc_o the Riesz-Caputo stencil weights -- the terms and their order are those of the dict-merging
// expansion, so the assembled matrix is bit-identical to the oracle's wherever the stencil weights are.
//
// Kernels: bands up to 31 wide and n <= 256 go, one warp per system, through the pivot-free block
// LDL^T of frx_solve_ldl.cuh (definite systems end there, indefinite ones only if refinement certifies their residual) and
// then, for what that rejects, through the partially pivoted blocked band LU of frx_solve_dmma.cuh; everything else through
// the CTA-per-system kernel of this file:
// The matrix never touches HBM: each CTA builds it in shared memory from the 2*res+1 stencil weights (produced
// for the whole batch by kernel K5), factors it by LU with partial pivoting, and solves.  A is banded
// (half-bandwidth nb = res+1; partial pivoting fills at most 2*nb above the diagonal), so it is stored in LAPACK
// band layout (3*nb+1 rows per column) whenever that is smaller than dense, and the elimination skips the
// structurally-zero part of the trailing matrix: n = 256, res = 32 fits in 205 KB of shared memory.
// Algorithmic bytes per system: 8*(2*res+1) weights + 8n rhs in, 8n solution out.
#include <cstring>
#include <chrono>
#include <math.h>
#include <stdlib.h>

#include "frx_internal.h"
#include "frx_rules.cuh"
#include "frx_solve_dmma.cuh"
#include "frx_solve_ldl.cuh"
#include "frx_solve_scalar.cuh"

namespace {

constexpr int kSolveThreads = 128;

struct SolveParams {
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
    double* dense;           // != NULL: write the assembled matrices (row-major n x n per system) and skip the solve
    int max_store;           // doubles of matrix storage reserved per CTA
    int a_len;               // doubles reserved for the band coefficients
};

__global__ void __launch_bounds__(kSolveThreads) frx_k4_assemble_solve(const SolveParams p) {
    extern __shared__ __align__(16) double sm[];
    const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* As = sm;                       // matrix storage
    double* b = sm + p.max_store;          // [n] right-hand side / solution
    double* a = b + n;                     // [2*nbmax+1] band coefficients a_k at a[k + nb]
    int* piv = reinterpret_cast<int*>(a + p.a_len);   // [n] pivot rows
    __shared__ int s_info;
    for (int64_t s = blockIdx.x; s < p.n_sys; s += gridDim.x) {
        // half-width of the Riesz-Caputo stencil: resolution, except 'rectangle' (resolution - 1, equation.py:70-71)
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        int nb = res + 1;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;   // rc[o], o = -res..res
        // ---- band coefficients -------------------------------------------------------------------------
        for (int k = -nb + tid; k <= nb; k += kSolveThreads) {
            double acc = 0.0;
            bool any = false;
            auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
            if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));          // T1
            if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));  // T2
            if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));    // T4
            if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));          // T3
            a[k + nb] = acc;
        }
        if (tid == 0) s_info = 0;
        __syncthreads();
        if (p.dense) {
            double* D = p.dense + s * (int64_t)n * n;
            for (int e = tid; e < n * n; e += kSolveThreads) {
                const int i = e / n, j = e - i * n, k = j - i;
                D[e] = (k >= -nb && k <= nb) ? a[k + nb] : 0.0;
            }
            __syncthreads();
            continue;
        }
        // ---- assemble into shared memory ------------------------------------------------------------------
        if (nb > n - 1) nb = n - 1;
        const bool band = 3 * nb + 1 < n;
        const int lda = band ? 3 * nb : n;        // element (i, j) lives at As[off + i + j*lda]
        const int off = band ? 2 * nb : 0;
        const int nb_a = res + 1;                 // index base of a[]
        for (int j = warp; j < n; j += kSolveThreads / 32) {
            const int lo = band ? max(0, j - 2 * nb) : 0, hi = band ? min(n - 1, j + nb) : n - 1;
            for (int i = lo + lane; i <= hi; i += 32) {
                const int k = j - i;
                As[off + i + j * lda] = (k >= -nb_a && k <= nb_a) ? a[k + nb_a] : 0.0;
            }
        }
        for (int i = tid; i < n; i += kSolveThreads) b[i] = p.rhs[s * p.ldb + i];
        __syncthreads();
        // ---- LU with partial pivoting on the augmented matrix [A | b], band-limited, ONE barrier per column ----
        // Warp 0 is the pivot warp: as soon as it has updated column k+1 it finds that column's pivot, swaps it to
        // the diagonal and scales the sub-column into multipliers in place (piv[k+1]).  Warps 1..3 update the other
        // columns of the window (lanes over the <= nb rows, the row interchange folded into the update) and b.
        auto AS = [&](int i, int j) -> double& { return As[off + i + j * lda]; };
        auto finalize_column = [&](int c) {       // warp 0, all lanes
            const int rm = min(n - 1, c + nb);
            const int i0 = c + lane, i1 = c + lane + 32;
            const double v0 = i0 <= rm ? AS(i0, c) : 0.0, v1 = i1 <= rm ? AS(i1, c) : 0.0;
            double best = i0 <= rm ? fabs(v0) : -1.0;
            int bi = i0;
            if (i1 <= rm && fabs(v1) > best) { best = fabs(v1); bi = i1; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, d);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            const int pv = bi;
            const double pval = AS(pv, c), top = AS(c, c);
            const double r = 1.0 / pval;
            __syncwarp();
            if (lane == 0) {
                AS(c, c) = pval;
                piv[c] = pv;
                if (pval == 0.0 && s_info == 0) s_info = c + 1;
            } else if (i0 <= rm) {
                AS(i0, c) = (i0 == pv ? top : v0) * r;
            }
            if (i1 <= rm) AS(i1, c) = (i1 == pv ? top : v1) * r;
            __syncwarp();
        };
        if (warp == 0) finalize_column(0);
        __syncthreads();
        constexpr int kWorkers = kSolveThreads / 32 - 1;
        for (int k = 0; k < n; ++k) {
            const int rmax = min(n - 1, k + nb);          // rows below are structurally zero in column k
            const int cmax = min(n - 1, k + 2 * nb);      // columns right of it are zero in rows k..rmax
            const int pv = piv[k];
            const int i0 = k + 1 + lane, i1 = k + 33 + lane;
            const double l0 = i0 <= rmax ? AS(i0, k) : 0.0, l1 = i1 <= rmax ? AS(i1, k) : 0.0;
            auto update_column = [&](int j) {
                const double ak = AS(k, j), u = AS(pv, j);             // u: the pivot row's entry after the interchange
                const double v0 = i0 <= rmax ? (i0 == pv ? ak : AS(i0, j)) : 0.0;
                const double v1 = i1 <= rmax ? (i1 == pv ? ak : AS(i1, j)) : 0.0;
                __syncwarp();
                if (lane == 0) AS(k, j) = u;
                if (i0 <= rmax) AS(i0, j) = fma(-l0, u, v0);
                if (i1 <= rmax) AS(i1, j) = fma(-l1, u, v1);
            };
            if (warp == 0) {
                if (k + 1 <= cmax) {
                    update_column(k + 1);
                    __syncwarp();
                    finalize_column(k + 1);
                }
            } else {
                for (int j = k + 1 + warp; j <= cmax; j += kWorkers) update_column(j);
                if (warp == 1 + (k % kWorkers)) {                     // right-hand side, rotated over the workers
                    const double bk = b[k], u = b[pv];
                    const double v0 = i0 <= rmax ? (i0 == pv ? bk : b[i0]) : 0.0;
                    const double v1 = i1 <= rmax ? (i1 == pv ? bk : b[i1]) : 0.0;
                    __syncwarp();
                    if (lane == 0) b[k] = u;
                    if (i0 <= rmax) b[i0] = fma(-l0, u, v0);
                    if (i1 <= rmax) b[i1] = fma(-l1, u, v1);
                }
            }
            __syncthreads();
        }
        // ---- back substitution on U (upper bandwidth 2*nb), one warp, column oriented ---------------------
        if (warp == 0) {
            for (int k = n - 1; k >= 0; --k) {
                const double xk = b[k] / As[off + k + k * lda];
                __syncwarp();
                if (lane == 0) b[k] = xk;
                const int lo = max(0, k - 2 * nb);
                for (int i = lo + lane; i < k; i += 32) b[i] = fma(-As[off + i + k * lda], xk, b[i]);
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += kSolveThreads) p.x[s * p.ldx + i] = b[i];
        if (tid == 0 && p.info) p.info[s] = s_info;
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------------------
// Streaming variant: ONE WARP PER SYSTEM, no block barriers.  Only the active window of the band elimination
// lives on chip: rows k..k+nb (lane r owns one row slot) x columns k..k+2nb (ring of C = 2nb+1 column slots) in
// shared memory, C odd -> the lanes' row-strided accesses are bank-conflict free.  Per column k:
//   pivot = lane with the largest |entry| in column k (ties: lowest row index, like LAPACK) -- rows are never
//   moved: the pivot lane's row simply retires as row k of U (implicit pivoting, row labels are swapped instead);
//   every other active lane subtracts l * (pivot row) from its row; the retired slot is refilled with matrix row
//   k+nb+1 straight from the 2nb+1 band coefficients; column slot k%C is zeroed and becomes column k+C.
// Retired U rows (2nb+1 entries) and the transformed right-hand side stream to a per-warp scratch in global memory
// (L2) and are read back, prefetched four rows ahead, by the back substitution.  Needs nb <= 31.
// ------------------------------------------------------------------------------------------------------------
// Shared memory per system is the occupancy limit (19 KB at the widest band -> 11 warps per SM), and the kernel is
// latency bound, so the host sorts the systems of a batch into bins by band width and launches one grid per bin with
// the window sized for that bin: narrow bands run with up to 32 one-warp CTAs per SM.
constexpr int kStreamWarps = 1;   // one-warp CTAs pack shared memory best
constexpr int kStreamMaxNb = 31;
constexpr int kStreamMaxN = 256;
constexpr int kStreamBins = 4;    // bins by nb = res + 1: <= 7, <= 15, <= 23, <= 31
constexpr int kDefaultTwoWarpBins = 0x0;  // FRX_K4_TWO_WARP bits 2,3: two warps per system in the wide-band bins (measured slower: DESIGN.md section 6)
constexpr int kDefaultPackedBins = 0x1;   // nb <= 7: four systems per warp (measured 1.38x; two per warp at nb <= 15 is slower: bit 1 off)
constexpr int kDefaultDmmaBins = 0xF;   // bins that run the blocked tensor-core kernel (frx_solve_dmma.cuh)
__host__ __device__ inline int stream_c(int nb_max) { return 2 * nb_max + 1; }
__host__ __device__ inline int stream_rs(int nb_max) { return ((stream_c(nb_max) + 7) & ~7) + 1; }   // ring padded to 8, +1 -> odd
// window + band coefficients (+1 pad) + rhs, rounded to an even number of doubles
__host__ __device__ inline int stream_warp_doubles(int nb_max) {
    return (32 * stream_rs(nb_max) + stream_c(nb_max) + 1 + kStreamMaxN + 1) & ~1;
}

struct StreamParams {
    int64_t n_sys;
    int n;
    const double* rcw;
    const int64_t* wptr;
    const double* h;
    const double* rhs;
    int64_t ldb;
    double* x;
    int64_t ldx;
    int32_t* info;
    double* uscratch;   // [total warps][n * (2 * nb_max + 1)]
    const int32_t* order;   // the systems of this launch (a bin), n_sys of them
    int nb_max;             // widest band (res + 1) among them: sizes the shared-memory window and the U rows
};

__global__ void __launch_bounds__(kStreamWarps * 32) frx_k4_band_stream(const StreamParams p) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = p.n;
    const int ustride = stream_c(p.nb_max);
    double* W = sm + warp * stream_warp_doubles(p.nb_max);   // W[row slot * RS + column slot]
    double* a = W + 32 * stream_rs(p.nb_max);            // a[k + nb]
    double* bs = a + ustride + 1;                        // right-hand side, overwritten in place by the transformed one
    const int64_t gw = (int64_t)blockIdx.x * kStreamWarps + warp, nw = (int64_t)gridDim.x * kStreamWarps;
    double* U = p.uscratch + gw * (int64_t)n * ustride;
    for (int64_t si = gw; si < p.n_sys; si += nw) {
        const int64_t s = p.order[si];
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const int nb_a = res + 1;
        const int nb = nb_a > n - 1 ? n - 1 : nb_a;      // rows below k+nb are zero in column k
        const int C = 2 * nb + 1;
        // the ring of C column slots is padded with always-zero slots to Cp (multiple of 8) so that the elimination
        // runs over whole batches of 8 slots without predicates; row stride Cp + 1 is odd -> conflict-free lanes
        const int Cp = (C + 7) & ~7, RS = Cp + 1;
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const double* rhs = p.rhs + s * p.ldb;
        double* xo = p.x + s * p.ldx;
        __syncwarp();
        for (int k = -nb_a + lane; k <= nb_a; k += 32) {   // band coefficients, same terms and order as the CTA kernel
            double acc = 0.0;
            bool any = false;
            auto add = [&](double t) { acc = any ? __dadd_rn(acc, t) : t; any = true; };
            if (k >= -res && k <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k], ih)));
            if (k + 1 >= -res && k + 1 <= res) add(__dmul_rn(-ih, __dmul_rn(rc[k + 1], -ih)));
            if (k - 1 >= -res && k - 1 <= res) add(__dmul_rn(ih, __dmul_rn(rc[k - 1], ih)));
            if (k >= -res && k <= res) add(__dmul_rn(ih, __dmul_rn(rc[k], -ih)));
            if (k >= -nb && k <= nb) a[k + nb] = acc;
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) bs[i] = rhs[i];
        __syncwarp();
        // initial window: rows 0..nb (lane = row), columns 0..2nb (column slot = column)
        int logical = (lane <= nb && lane < n) ? lane : -1;
        double bval = logical >= 0 ? bs[lane] : 0.0;
        if (logical >= 0)
            for (int j = 0; j < Cp; ++j) {
                const int d = j - lane;
                W[lane * RS + j] = (j < C && d >= -nb && d <= nb && j < n) ? a[d + nb] : 0.0;
            }
        int info = 0, kc = 0;                              // kc = k % C
        __syncwarp();
        for (int k = 0; k < n; ++k) {
            const bool active = logical >= 0;
            const double v = active ? W[lane * RS + kc] : 0.0;
            // pivot = active lane with the largest |v|, ties -> lowest row label (LAPACK's first maximum).  |v| >= 0, so
            // its bit pattern orders like an unsigned integer: two 32-bit warp reductions instead of a shuffle tree
            const unsigned long long bits = active ? (unsigned long long)__double_as_longlong(fabs(v)) : 0ull;
            const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            const bool cand = active && hi == mh && lo == ml;
            const int blog = __reduce_min_sync(0xffffffffu, cand ? logical : 0x7fffffff);
            const int P = __ffs(__ballot_sync(0xffffffffu, cand && logical == blog)) - 1;   // pivot lane; blog = its label
            const double pval = __shfl_sync(0xffffffffu, v, P), bt = __shfl_sync(0xffffffffu, bval, P);
            if (pval == 0.0 && info == 0) info = k + 1;
            const bool worker = active && lane != P;
            const double l = worker ? v * (1.0 / pval) : 0.0;
            if (worker && logical == k) logical = blog;    // the interchange, on labels only
            const int cmax = min(n - 1, k + 2 * nb), len = cmax - k;   // U row k: columns k..cmax
            const double* prow = W + P * RS;
            // retire the pivot row: U[k][0..len] and the transformed right-hand side
            for (int c = lane; c <= len; c += 32) {
                int cs = kc + c; if (cs >= C) cs -= C;
                U[(int64_t)k * ustride + c] = prow[cs];
            }
            if (lane == 0) bs[k] = bt;                     // slot k is free: rhs[k] entered the window long ago
            // eliminate
            if (worker) {
                // every slot of the padded ring, in batches of 8: slots beyond cmax and the pad slots hold zeros in both
                // rows, slot kc (column k itself) is overwritten below.  A warp issues in order, so the 16 loads of a
                // batch go out back to back, then the FMAs, then the stores; the pivot row is another lane's row, hence
                // __restrict__.
                double* __restrict__ myrow = W + lane * RS;
                const double* __restrict__ pr = prow;
                for (int cs = 0; cs < Cp; cs += 8) {
                    double pu[8], mv[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { pu[q] = pr[cs + q]; mv[q] = myrow[cs + q]; }
#pragma unroll
                    for (int q = 0; q < 8; ++q) myrow[cs + q] = fma(-l, pu[q], mv[q]);
                }
                bval = fma(-l, bt, bval);
                myrow[kc] = 0.0;                           // column slot kc becomes column k + C
            }
            __syncwarp();
            // refill the retired slot with matrix row k+nb+1 (columns k+1 .. k+2nb+1 = the whole new window)
            const int inew = k + nb + 1;
            if (inew < n) {
                double* nrow = W + P * RS;
                for (int c = lane; c < C; c += 32) {
                    const int j = k + 1 + c;
                    int cs = kc + 1 + c; if (cs >= C) cs -= C;
                    nrow[cs] = j < n ? a[j - inew + nb] : 0.0;
                }   // (the pad slots C..Cp-1 of this row slot are still zero: nothing ever writes a non-zero there)
                if (lane == P) { logical = inew; bval = bs[inew]; }
            } else if (lane == P) {
                logical = -1;
            }
            kc = kc + 1 == C ? 0 : kc + 1;
            __syncwarp();
        }
        // back substitution: x_k = (bt_k - sum_{c=1..len} U[k][c] x_{k+c}) / U[k][0]; U rows prefetched 4 ahead
        double* xs = W;                                    // the window is free now: xs[n] (n <= 32*C is checked on the host)
        constexpr int PF = 4;
        double u0[PF], u1[PF], u2[PF];
        auto load_row = [&](int k, double& r0, double& r1, double& r2) {
            const int len = min(n - 1, k + 2 * nb) - k;
            const double* ur = U + (int64_t)k * ustride;
            r0 = (k >= 0 && lane <= len) ? __ldcg(ur + lane) : 0.0;
            r1 = (k >= 0 && lane + 32 <= len) ? __ldcg(ur + lane + 32) : 0.0;
            r2 = 0.0;
        };
#pragma unroll
        for (int q = 0; q < PF; ++q) load_row(n - 1 - q, u0[q], u1[q], u2[q]);
        for (int kb = n - 1; kb >= 0; kb -= PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int k = kb - q;
                if (k < 0) break;
                const int len = min(n - 1, k + 2 * nb) - k;
                // lane holds U[k][lane], U[k][lane+32]; entry c multiplies x_{k+c}, c >= 1
                double sum = 0.0;
                if (lane >= 1 && lane <= len) sum = u0[q] * xs[k + lane];
                if (lane + 32 <= len) sum = fma(u1[q], xs[k + lane + 32], sum);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
                const double diag = __shfl_sync(0xffffffffu, u0[q], 0);
                const double xk = (bs[k] - sum) / diag;
                if (lane == 0) { xs[k] = xk; }
                load_row(k - PF, u0[q], u1[q], u2[q]);
                __syncwarp();
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) xo[i] = xs[i];
        if (lane == 0 && p.info) p.info[s] = info;
        __syncwarp();
    }
}

}  // namespace

// launch geometry of one band-width bin: grid (persistent one-warp CTAs), dynamic shared memory, U scratch
struct BinLaunch {
    int64_t grid;
    size_t dyn, scratch;
    int kind;   // 0: frx_k4_band_stream, 1..4: frx_k4_band_dmma<1,3> / <2,5> / <3,7> / <4,9>, 5..6: frx_k4_band_dmma2<3,7> / <4,9>, 7..8: frx_k4_band_dmma_packed<1,3,8> / <2,5,16>
};

// U scratch of one CTA in doubles: a multiple of 128 bytes so that every CTA's slab is line-aligned (discard.global.L2)
static inline int64_t band_ustride_cta(int n, int nb_bin) { return ((int64_t)n * (2 * nb_bin + 1) + 15) / 16 * 16; }

template <int RT, int CT>
static int band_dmma_geometry(frx_ctx* ctx, int n, int64_t count, int nb_bin, BinLaunch* out) {
    using G = frx_k4::BandGeo<RT, CT>;
    out->dyn = sizeof(double) * (size_t)G::smem_doubles(n);
    FRX_CUDA(cudaFuncSetAttribute(frx_k4::frx_k4_band_dmma<RT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->dyn));
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4::frx_k4_band_dmma<RT, CT>, 32, out->dyn));
    if (per_sm < 1) per_sm = 1;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    if (out->grid > count) out->grid = count;
    out->scratch = sizeof(double) * (size_t)out->grid * band_ustride_cta(n, nb_bin);
    return FRX_OK;
}

template <int RT, int CT, int GS>
static int band_packed_geometry(frx_ctx* ctx, int n, int64_t count, int nb_bin, BinLaunch* out) {
    using G = frx_k4::BandGeo<RT, CT>;
    constexpr int NG = 32 / GS;
    out->dyn = sizeof(double) * (size_t)NG * G::smem_doubles(n);
    FRX_CUDA(cudaFuncSetAttribute(frx_k4::frx_k4_band_dmma_packed<RT, CT, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)out->dyn));
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4::frx_k4_band_dmma_packed<RT, CT, GS>, 32, out->dyn));
    if (per_sm < 1) per_sm = 1;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    const int64_t warps_needed = (count + NG - 1) / NG;
    if (out->grid > warps_needed) out->grid = warps_needed;
    out->scratch = sizeof(double) * (size_t)out->grid * NG * band_ustride_cta(n, nb_bin);
    return FRX_OK;
}

template <int RT, int CT, int GS>
static int band_packed_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_dmma_packed<RT, CT, GS><<<(unsigned)L.grid, 32, L.dyn, ctx->stream>>>(q));
    return FRX_OK;
}

template <int RT, int CT>
static int band_dmma2_geometry(frx_ctx* ctx, int n, int64_t count, int nb_bin, BinLaunch* out) {
    using G = frx_k4::BandGeo<RT, CT>;
    out->dyn = sizeof(double) * (size_t)G::smem_doubles2(n);
    FRX_CUDA(cudaFuncSetAttribute(frx_k4::frx_k4_band_dmma2<RT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->dyn));
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4::frx_k4_band_dmma2<RT, CT>, 64, out->dyn));
    if (per_sm < 1) per_sm = 1;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    if (out->grid > count) out->grid = count;
    out->scratch = sizeof(double) * (size_t)out->grid * band_ustride_cta(n, nb_bin);
    return FRX_OK;
}

template <int RT, int CT>
static int band_dmma2_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_dmma2<RT, CT><<<(unsigned)L.grid, 64, L.dyn, ctx->stream>>>(q));
    return FRX_OK;
}


template <int T>
static int band_ldl_geometry(frx_ctx* ctx, int n, int64_t count, int nb_bin, BinLaunch* out) {
    using G = frx_k4::LdlGeo<T>;
    out->dyn = sizeof(double) * (size_t)G::smem_doubles(n);
    FRX_CUDA(cudaFuncSetAttribute(frx_k4::frx_k4_band_ldl<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->dyn));
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4::frx_k4_band_ldl<T>, 32, out->dyn));
    if (per_sm < 1) per_sm = 1;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    if (out->grid > count) out->grid = count;
    out->scratch = sizeof(double) * (size_t)out->grid * (size_t)G::ustride(n);
    return FRX_OK;
}

template <int T>
static int band_bldl_geometry(frx_ctx* ctx, int n, int64_t count, int nb_bin, BinLaunch* out) {
    using G = frx_k4::BldlGeo<T>;
    (void)nb_bin;
    out->dyn = sizeof(double) * (size_t)G::smem_doubles(n);
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4::frx_k4_band_bldl<T>, 32, out->dyn));
    if (per_sm < 1) per_sm = 1;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    if (out->grid > count) out->grid = count;
    out->scratch = sizeof(double) * (size_t)out->grid * (size_t)G::ustride(n);
    return FRX_OK;
}

// one thread per system (nb <= 7): a few single-warp CTAs per SM, each with one scratch slab that stays in L2
static int band_scalar_geometry(frx_ctx* ctx, int n, int64_t count, BinLaunch* out) {
    using G = frx_k4::ScalarGeo;
    out->dyn = 0;
    // Single-warp CTAs per SM (FRX_K4_SCALAR_WARPS).  Each keeps one slab of n * 2 KB.  Few warps keep the slabs in L2 but
    // leave every dependent chain exposed; measured per 1e5 systems with 1 / 2 / 4 / 8 / 12 warps per SM
    // (gpurun_out/k4_scalar_sweep.txt): n = 64: 466 / 246 / 158 / 144 / 145 us, n = 128: 875 / 472 / 364 / 311 / 304 us,
    // n = 256: 1791 / 1122 / 757 / 633 / 613 us -- from 8 warps on the kernel runs at the HBM rate of its slab traffic
    // (n = 128: 1.6 GB at 5.2 TB/s).
    static const int per_sm_env = getenv("FRX_K4_SCALAR_WARPS") ? atoi(getenv("FRX_K4_SCALAR_WARPS")) : 0;
    const int per_sm = per_sm_env > 0 ? per_sm_env : 8;
    out->grid = (int64_t)ctx->sm_count * per_sm;
    if (out->grid > (count + 31) / 32) out->grid = (count + 31) / 32;
    out->scratch = sizeof(double) * (size_t)out->grid * (size_t)G::ustride(n);
    return FRX_OK;
}

static int band_scalar_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_scalar<<<(unsigned)L.grid, 32, 0, ctx->stream>>>(q));
    return FRX_OK;
}

template <int T>
static int band_bldl_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_bldl<T><<<(unsigned)L.grid, 32, L.dyn, ctx->stream>>>(q));
    return FRX_OK;
}

template <int T>
static int band_ldl_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_ldl<T><<<(unsigned)L.grid, 32, L.dyn, ctx->stream>>>(q));
    return FRX_OK;
}

template <int RT, int CT>
static int band_dmma_launch(frx_ctx* ctx, const frx_k4::BandParams& q, const BinLaunch& L) {
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4::frx_k4_band_dmma<RT, CT><<<(unsigned)L.grid, 32, L.dyn, ctx->stream>>>(q));
    return FRX_OK;
}

static int solve_common(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                        const double* h_res, const double* d_rhs, int64_t ldb, double* d_x, int64_t ldx, int32_t* d_info,
                        double* d_dense, bool host_io = false) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(n_sys >= 0 && n_sys < ((int64_t)1 << 31) && n >= 1 && n <= 4096, FRX_ERR_INVALID_ARG, "bad n_sys / n");
    if (n_sys == 0) return FRX_OK;
    FRX_REQUIRE(h_alpha && h_h && h_res, FRX_ERR_INVALID_ARG, "NULL parameter array");
    FRX_REQUIRE(d_dense || (d_rhs && d_x && ldb >= n && ldx >= n), FRX_ERR_INVALID_ARG, "NULL buffer or bad leading dimension");
    FRX_ON_DEVICE(ctx);
    // FRX_K4_HOST_PROFILE=1: host microseconds of this call by phase, on stderr
    static const bool host_profile = getenv("FRX_K4_HOST_PROFILE") && atoi(getenv("FRX_K4_HOST_PROFILE"));
    auto t_now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = host_profile ? t_now() : 0.0;
    double t_stage = 0.0, t_pass = 0.0, t_k5 = 0.0, t_geo = 0.0;
    if (!d_dense) {
        ctx->k4_n_sys = n_sys;
        ctx->k4_ldl = 0;
    }
    // One host pass (domain checks, stencil row pointers, band-width bins) into the context's pinned staging block,
    // laid out like its device image:  [alpha | lf | res | h : 4 n_sys doubles][wptr : n_sys + 1][order : n_sys int32]
    const size_t o_par = 0;
    const size_t o_wptr = frx_align_up(sizeof(double) * 4 * (size_t)n_sys, 256);
    const size_t o_order = o_wptr + frx_align_up(sizeof(int64_t) * (size_t)(n_sys + 1), 256);
    const size_t stage_bytes = o_order + frx_align_up(sizeof(int32_t) * (size_t)n_sys, 256);
    int slot = 0;
    void* host_image = nullptr;
    int rc = frx_k4_stage_begin(ctx, stage_bytes, &slot, &host_image);
    if (rc) return rc;
    if (host_profile) t_stage = t_now();
    char* hs = (char*)host_image;
    double* s_par = (double*)(hs + o_par);
    int64_t* wptr = (int64_t*)(hs + o_wptr);
    int32_t* h_order = (int32_t*)(hs + o_order);
    auto bin_of = [](int nb_a) { return nb_a <= 7 ? 0 : nb_a <= 15 ? 1 : nb_a <= 23 ? 2 : 3; };
    int64_t bin_count[kStreamBins] = {0, 0, 0, 0}, bin_begin[kStreamBins + 1];
    int bin_nb[kStreamBins] = {0, 0, 0, 0};
    wptr[0] = 0;
    int max_store = 0, max_res = 0;
    // (the common case inline: the full check -- and its error text -- only for a setting the quick test does not pass;
    // 0.6-0.7 ms per 1e5 systems on one core of the GPU box.  Measured and dropped: the pass in chunks on 4-8 threads, 1.5 ms;
    // non-temporal stores for the image, 0.82 ms)
    const double alpha_max = rule == FRX_RULE_RECTANGLE ? 1.0 : 2.0, res_min = rule == FRX_RULE_CAPUTO || rule == FRX_RULE_TRAPEZOIDAL ? 1.0 : 2.0;
    const bool quick_rule = rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_SIMPSON;
    for (int64_t s = 0; s < n_sys; ++s) {
        int32_t first, count;
        const double al = h_alpha[s], rs = h_res[s];
        if (quick_rule && al >= 0.0 && al < alpha_max && rs >= res_min && rs <= 1.0e9 && rs == (double)(long long)rs) {
            long long f, c;
            frx_left_layout(rule, (long long)rs, &f, &c);
            const long long last = f + c - 1, lo = f < -last ? f : -last;       // as frx_stencil_layout, Riesz-Caputo side
            first = (int32_t)lo;
            count = (int32_t)(-2 * lo + 1);
        } else {
            rc = frx_stencil_layout(rule, FRX_SIDE_RIESZ_CAPUTO, al, rs, &first, &count);
            if (rc) return rc;
        }
        if (!(h_h[s] > 0.0)) { frx_set_error("grid step must be positive"); return FRX_ERR_DOMAIN; }
        if (rule == FRX_RULE_SIMPSON && ((int64_t)h_res[s] & 1)) {
            frx_set_error("'simpson' with odd resolution has a node right of the point; use an even resolution here");
            return FRX_ERR_DOMAIN;
        }
        wptr[s + 1] = wptr[s] + count;
        s_par[s] = h_alpha[s];
        s_par[n_sys + s] = h_res[s] * h_h[s];
        s_par[2 * n_sys + s] = h_res[s];
        s_par[3 * n_sys + s] = h_h[s];
        const int res = (count - 1) / 2, nb_a = res + 1, nb = nb_a > n - 1 ? n - 1 : nb_a;   // res = stencil half-width
        const int store = d_dense ? 0 : ((3 * nb + 1 < n) ? n * (3 * nb + 1) : n * n);      // band columns are 3nb+1 long
        if (store > max_store) max_store = store;
        if (res > max_res) max_res = res;
        const int b = bin_of(nb_a);
        ++bin_count[b];
        if (nb_a > bin_nb[b]) bin_nb[b] = nb_a;
    }
    bin_begin[0] = 0;
    for (int b = 0; b < kStreamBins; ++b) bin_begin[b + 1] = bin_begin[b] + bin_count[b];
    {
        int64_t fill[kStreamBins] = {bin_begin[0], bin_begin[1], bin_begin[2], bin_begin[3]};
        for (int64_t s = 0; s < n_sys; ++s) h_order[fill[bin_of((int)((wptr[s + 1] - wptr[s] - 1) / 2) + 1)]++] = (int32_t)s;
    }
    if (host_profile) t_pass = t_now();
    const int64_t total_w = wptr[n_sys];
    const int a_len = (2 * (max_res + 1) + 1 + 1) & ~1;
    const size_t smem = sizeof(double) * ((size_t)((max_store + 1) & ~1) + n + a_len) + sizeof(int) * (size_t)(n + 4);
    // workspace: [rc weights][K5 scalars][fail lists]
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += frx_align_up(bytes, 256); return o; };
    const size_t o_w = carve(sizeof(double) * (size_t)total_w);
    extern size_t frx_k5_scratch_bytes(int64_t);
    const size_t o_k5 = carve(frx_k5_scratch_bytes(n_sys));
    const size_t o_fail = carve(sizeof(int32_t) * (size_t)n_sys);       // fail lists of the pivot-free attempt, bin by bin
    const size_t o_retry = carve(sizeof(int32_t) * (size_t)(bin_count[0] + 1));   // what frx_k4_band_scalar hands to the block LDL^T
    rc = frx_ws_reserve(ctx, off);
    if (rc) return rc;
    char* ws = (char*)ctx->ws;
    // the image travels on the context's copy stream into the device block of this call's parity; this stream waits for it
    void* device_image = nullptr;
    rc = frx_k4_stage_upload(ctx, slot, stage_bytes, &device_image);
    if (rc) return rc;
    struct StageDone {      // whatever path the call takes from here: the block's readers end with what this call queues
        frx_ctx* c; int k;
        ~StageDone() { frx_k4_stage_done(c, k); }
    } stage_done{ctx, slot};
    char* img = (char*)device_image;
    double* d_par = (double*)(img + o_par);
    // K5 for the whole batch (row pointer already on the device)
    extern int frx_launch_stencil_weights(frx_ctx*, int, int, int64_t, const double*, const double*, const double*,
                                          const int64_t*, int64_t, void*, int32_t*, double*);
    rc = frx_launch_stencil_weights(ctx, rule, FRX_SIDE_RIESZ_CAPUTO, n_sys, d_par, d_par + n_sys, d_par + 2 * n_sys,
                                    (const int64_t*)(img + o_wptr), total_w, ws + o_k5, nullptr, (double*)(ws + o_w));
    if (rc) return rc;
    if (host_profile) t_k5 = t_now();
    // FRX_K4_VARIANT: 1 = CTA-per-system kernel for everything, 2 = column-at-a-time warp kernel for every bin;
    // FRX_K4_DMMA_BINS: bit b set = bin b runs the blocked tensor-core kernel (default: see kDefaultDmmaBins)
    static const int variant = getenv("FRX_K4_VARIANT") ? atoi(getenv("FRX_K4_VARIANT")) : 0;
    static const int dmma_bins = variant == 2 ? 0 : (getenv("FRX_K4_DMMA_BINS") ? atoi(getenv("FRX_K4_DMMA_BINS")) : kDefaultDmmaBins);
    // FRX_K4_TWO_WARP: bit b set = bin b (2 or 3) runs the two-warps-per-system form of the blocked kernel
    // FRX_K4_PACKED: bit b set = bin b (0 or 1) runs several systems per warp (frx_k4_band_dmma_packed)
    static const int packed_bins = getenv("FRX_K4_PACKED") ? atoi(getenv("FRX_K4_PACKED")) : kDefaultPackedBins;
    static const int two_warp_bins = getenv("FRX_K4_TWO_WARP") ? atoi(getenv("FRX_K4_TWO_WARP")) : kDefaultTwoWarpBins;
    // FRX_K4_LDL: 1 (default) = every system is first tried without pivoting (frx_k4_band_bldl: definite systems end there),
    // the pivoted kernels work through what that attempt rejects; 2 = the same with the first form of the attempt
    // (frx_k4_band_ldl, scalar pivots, tiles in shared memory); 0 = pivoted kernels for everything
    static const int use_ldl_env = getenv("FRX_K4_LDL") ? atoi(getenv("FRX_K4_LDL")) : 1;
    const bool use_ldl = use_ldl_env && variant == 0 && dmma_bins == 0xF && two_warp_bins == 0;
    // FRX_K4_SCALAR: 1 (default) = the narrowest bin (nb <= 7) is first tried with one THREAD per system
    // (frx_k4_band_scalar: definite systems end there), the block LDL^T works through what that rejects; 0 = off
    static const int use_scalar_env = getenv("FRX_K4_SCALAR") ? atoi(getenv("FRX_K4_SCALAR")) : 1;
    // (not when the kernels read and write page-locked HOST buffers in place, unless FRX_K4_SCALAR=2: a group of 32 systems
    // reads all its right-hand sides before it writes any solution, and with ~1 group per warp the two directions of the
    // PCIe link are used one after the other: measured 4.38 against 4.15 ms end to end per 1e5 systems of n = 128)
    const bool use_scalar = use_ldl && use_ldl_env == 1 && use_scalar_env && bin_count[0] > 0 && (!host_io || use_scalar_env == 2);
    if (!d_dense && variant != 1 && max_res + 1 <= kStreamMaxNb && n <= kStreamMaxN) {
        frx_k4::BandParams q;
        q.n = n;
        q.rcw = (const double*)(ws + o_w);
        q.wptr = (const int64_t*)(img + o_wptr);
        q.h = d_par + 3 * n_sys;
        q.rhs = d_rhs;
        q.ldb = ldb;
        q.x = d_x;
        q.ldx = ldx;
        q.info = d_info;
        q.n_sys_dev = nullptr;
        q.fail_list = nullptr;
        q.fail_count = nullptr;
        q.refined_count = nullptr;
        q.queue = nullptr;
        BinLaunch L[kStreamBins], Lf[kStreamBins], Ls;
        size_t scratch = 0;
        Ls.grid = 0;
        if (use_scalar) {
            rc = band_scalar_geometry(ctx, n, bin_count[0], &Ls);
            if (rc) return rc;
            scratch = Ls.scratch;
        }
        for (int b = 0; b < kStreamBins; ++b) {
            L[b].grid = 0;
            Lf[b].grid = 0;
            if (!bin_count[b]) continue;
            if (use_ldl) {
                if (use_ldl_env == 2)
                    rc = b == 0 ? band_ldl_geometry<1>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                       : b == 1 ? band_ldl_geometry<2>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                       : b == 2 ? band_ldl_geometry<3>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                                : band_ldl_geometry<4>(ctx, n, bin_count[b], bin_nb[b], &Lf[b]);
                else
                    rc = b == 0 ? band_bldl_geometry<1>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                       : b == 1 ? band_bldl_geometry<2>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                       : b == 2 ? band_bldl_geometry<3>(ctx, n, bin_count[b], bin_nb[b], &Lf[b])
                                : band_bldl_geometry<4>(ctx, n, bin_count[b], bin_nb[b], &Lf[b]);
                if (rc) return rc;
                if (Lf[b].scratch > scratch) scratch = Lf[b].scratch;
            }
            if ((packed_bins & (1 << b)) && b <= 1 && (dmma_bins & (1 << b))) {
                L[b].kind = 7 + b;     // 7, 8: several systems per warp, bins <1,3> (4 systems) and <2,5> (2 systems)
                rc = b == 0 ? band_packed_geometry<1, 3, 8>(ctx, n, bin_count[b], bin_nb[b], &L[b])
                            : band_packed_geometry<2, 5, 16>(ctx, n, bin_count[b], bin_nb[b], &L[b]);
                if (rc) return rc;
            } else if ((two_warp_bins & (1 << b)) && b >= 2 && (dmma_bins & (1 << b))) {
                L[b].kind = 3 + b;     // 5, 6: the two-warp kernel for the bins <3,7> and <4,9>
                rc = b == 2 ? band_dmma2_geometry<3, 7>(ctx, n, bin_count[b], bin_nb[b], &L[b])
                            : band_dmma2_geometry<4, 9>(ctx, n, bin_count[b], bin_nb[b], &L[b]);
                if (rc) return rc;
            } else if (dmma_bins & (1 << b)) {
                L[b].kind = 1 + b;
                rc = b == 0 ? band_dmma_geometry<1, 3>(ctx, n, bin_count[b], bin_nb[b], &L[b])
                   : b == 1 ? band_dmma_geometry<2, 5>(ctx, n, bin_count[b], bin_nb[b], &L[b])
                   : b == 2 ? band_dmma_geometry<3, 7>(ctx, n, bin_count[b], bin_nb[b], &L[b])
                            : band_dmma_geometry<4, 9>(ctx, n, bin_count[b], bin_nb[b], &L[b]);
                if (rc) return rc;
            } else {
                L[b].kind = 0;
                L[b].dyn = sizeof(double) * (size_t)kStreamWarps * stream_warp_doubles(bin_nb[b]);
                FRX_CUDA(cudaFuncSetAttribute(frx_k4_band_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L[b].dyn));
                int per_sm = 1;
                FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4_band_stream, kStreamWarps * 32, L[b].dyn));
                if (per_sm < 1) per_sm = 1;
                L[b].grid = (int64_t)ctx->sm_count * per_sm;
                if (L[b].grid * kStreamWarps > bin_count[b]) L[b].grid = (bin_count[b] + kStreamWarps - 1) / kStreamWarps;
                L[b].scratch = sizeof(double) * (size_t)L[b].grid * kStreamWarps * n * stream_c(bin_nb[b]);
            }
            if (L[b].scratch > scratch) scratch = L[b].scratch;
        }
        // FRX_K4_CONCURRENT: 1 (default) = the chains of the four band-width bins (thread-per-system kernel -> block LDL^T ->
        // pivoted LU on the fail list) run on four streams of the context, forked from and joined to its own stream: the
        // bins share nothing but read-only inputs, and the tail of one launch (a few resident CTAs finishing their last
        // system, or a pivoted kernel with a handful of systems) is filled with the CTAs of the next; 0 = one stream
        static const int conc_env = getenv("FRX_K4_CONCURRENT") ? atoi(getenv("FRX_K4_CONCURRENT")) : 1;
        const bool concurrent = use_ldl && conc_env != 0;
        size_t bin_scratch[kStreamBins] = {0, 0, 0, 0}, bin_off[kStreamBins] = {0, 0, 0, 0}, io_total = scratch;
        if (concurrent) {                      // every bin its own region of the scratch
            io_total = 0;
            for (int b = 0; b < kStreamBins; ++b) {
                if (!bin_count[b]) continue;
                bin_scratch[b] = Lf[b].scratch > L[b].scratch ? Lf[b].scratch : L[b].scratch;
                if (b == 0 && use_scalar && Ls.scratch > bin_scratch[b]) bin_scratch[b] = Ls.scratch;
                bin_off[b] = io_total;
                io_total += frx_align_up(bin_scratch[b], 256);
            }
        }
        if (host_profile) t_geo = t_now();
        rc = frx_io_reserve(ctx, io_total);    // factor scratch of the resident systems (each warp reuses its slab)
        if (rc) return rc;
        q.uscratch = (double*)ctx->io;
        ctx->k4_ldl = use_ldl ? 1 : 0;
        // the pivot-free attempt of bin b (thread per system for the narrowest bin, then the block LDL^T)
        auto launch_attempt = [&](const int b) -> int {
            int rc2 = FRX_OK;
            q.n_sys = bin_count[b];
            q.n_sys_dev = nullptr;
            q.order = (const int32_t*)(img + o_order) + bin_begin[b];
            q.nb_max = bin_nb[b];
            if (use_ldl_env == 2)
                q.ustride_cta = b == 0 ? frx_k4::LdlGeo<1>::ustride(n) : b == 1 ? frx_k4::LdlGeo<2>::ustride(n)
                              : b == 2 ? frx_k4::LdlGeo<3>::ustride(n) : frx_k4::LdlGeo<4>::ustride(n);
            else
                q.ustride_cta = b == 0 ? frx_k4::BldlGeo<1>::ustride(n) : b == 1 ? frx_k4::BldlGeo<2>::ustride(n)
                              : b == 2 ? frx_k4::BldlGeo<3>::ustride(n) : frx_k4::BldlGeo<4>::ustride(n);
            q.fail_list = (int32_t*)(ws + o_fail) + bin_begin[b];
            q.fail_count = (int32_t*)((char*)ctx->scalars + 256) + b;
            q.refined_count = (int32_t*)((char*)ctx->scalars + 256) + kStreamBins + b;
            q.queue = (int32_t*)((char*)ctx->scalars + 256) + 2 * kStreamBins + b;
            if (b == 0 && use_scalar) {
                // one thread per system first; the block LDL^T then takes the retry list (counted on the device)
                frx_k4::BandParams qs = q;
                qs.ustride_cta = frx_k4::ScalarGeo::ustride(n);
                qs.fail_list = (int32_t*)(ws + o_retry);
                qs.fail_count = (int32_t*)((char*)ctx->scalars + 256) + 3 * kStreamBins;
                rc2 = band_scalar_launch(ctx, qs, Ls);
                if (rc2) return rc2;
                q.order = qs.fail_list;
                q.n_sys_dev = qs.fail_count;
            }
            if (use_ldl_env == 2)
                rc2 = b == 0 ? band_ldl_launch<1>(ctx, q, Lf[b]) : b == 1 ? band_ldl_launch<2>(ctx, q, Lf[b])
                    : b == 2 ? band_ldl_launch<3>(ctx, q, Lf[b]) : band_ldl_launch<4>(ctx, q, Lf[b]);
            else
                rc2 = b == 0 ? band_bldl_launch<1>(ctx, q, Lf[b]) : b == 1 ? band_bldl_launch<2>(ctx, q, Lf[b])
                    : b == 2 ? band_bldl_launch<3>(ctx, q, Lf[b]) : band_bldl_launch<4>(ctx, q, Lf[b]);
            q.fail_list = nullptr;
            q.fail_count = nullptr;
            q.refined_count = nullptr;
            q.n_sys_dev = nullptr;
            return rc2;
        };
        // the pivoted kernels of bin b: every system of the bin, or (behind the attempt) its fail list, counted on the device
        auto launch_pivoted = [&](const int b) -> int {
            int rc2 = FRX_OK;
            q.n_sys = bin_count[b];
            q.n_sys_dev = nullptr;
            q.order = (const int32_t*)(img + o_order) + bin_begin[b];
            if (use_ldl) {
                q.order = (const int32_t*)(ws + o_fail) + bin_begin[b];
                q.n_sys_dev = (const int32_t*)((char*)ctx->scalars + 256) + b;
            }
            q.nb_max = bin_nb[b];
            q.ustride_cta = band_ustride_cta(n, bin_nb[b]);
            if (L[b].kind == 0) {
                StreamParams sp;
                sp.n_sys = q.n_sys; sp.n = n; sp.rcw = q.rcw; sp.wptr = q.wptr; sp.h = q.h; sp.rhs = q.rhs; sp.ldb = ldb;
                sp.x = d_x; sp.ldx = ldx; sp.info = d_info; sp.uscratch = q.uscratch; sp.order = q.order; sp.nb_max = q.nb_max;
                // (the attribute is the maximum the function may use: raise it to this launch's size)
                FRX_CUDA(cudaFuncSetAttribute(frx_k4_band_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L[b].dyn));
                FRX_LAUNCH(ctx, FRX_K_SOLVE,
                           frx_k4_band_stream<<<(unsigned)L[b].grid, kStreamWarps * 32, L[b].dyn, ctx->stream>>>(sp));
            } else if (L[b].kind >= 7) {
                rc2 = L[b].kind == 7 ? band_packed_launch<1, 3, 8>(ctx, q, L[b]) : band_packed_launch<2, 5, 16>(ctx, q, L[b]);
            } else if (L[b].kind >= 5) {
                rc2 = L[b].kind == 5 ? band_dmma2_launch<3, 7>(ctx, q, L[b]) : band_dmma2_launch<4, 9>(ctx, q, L[b]);
            } else {
                rc2 = L[b].kind == 1 ? band_dmma_launch<1, 3>(ctx, q, L[b])
                    : L[b].kind == 2 ? band_dmma_launch<2, 5>(ctx, q, L[b])
                    : L[b].kind == 3 ? band_dmma_launch<3, 7>(ctx, q, L[b])
                                     : band_dmma_launch<4, 9>(ctx, q, L[b]);
            }
            q.n_sys_dev = nullptr;
            return rc2;
        };
        if (use_ldl)
            FRX_CUDA(cudaMemsetAsync((char*)ctx->scalars + 256, 0, sizeof(int32_t) * (3 * kStreamBins + 1), ctx->stream));
        if (concurrent) {
            rc = frx_bin_streams(ctx);
            if (rc) return rc;
            // with per-kernel timing on, the fork-to-join region is ONE timed interval of the context's stream
            rc = frx_timing_begin(ctx, FRX_K_SOLVE);
            if (rc) return rc;
            struct Restore {       // the context's stream and timing flag come back on every exit path
                frx_ctx* c; cudaStream_t s; int t;
                ~Restore() { c->stream = s; c->timing = t; }
            } restore{ctx, ctx->stream, ctx->timing};
            cudaStream_t main_stream = ctx->stream;
            FRX_CUDA(cudaEventRecord(ctx->bin_fork, main_stream));
            ctx->timing = 0;
            // FRX_K4_BIN_ORDER: the order in which the chains are queued.  Default "0321": the narrowest bin first -- its
            // thread-per-system kernel runs at the HBM rate and takes few warps per SM, so it shares the SMs with the
            // compute-bound widest bin instead of running alone at the end -- then widest first.  K4 kernels per 1e5 systems,
            // "3210" -> "0321": 0.769 -> 0.734 ms (n = 64), 1.503 -> 1.460 (n = 128), 3.492 -> 3.468 (n = 256).
            static const char* order_env = getenv("FRX_K4_BIN_ORDER");
            static const int default_order[kStreamBins] = {0, 3, 2, 1};
            for (int k = 0; k < kStreamBins; ++k) {
                int b = default_order[k];
                if (order_env && strlen(order_env) == kStreamBins && order_env[k] >= '0' && order_env[k] < '0' + kStreamBins) b = order_env[k] - '0';
                if (!bin_count[b]) continue;
                FRX_CUDA(cudaStreamWaitEvent(ctx->bin_stream[b], ctx->bin_fork, 0));
                ctx->stream = ctx->bin_stream[b];
                q.uscratch = (double*)((char*)ctx->io + bin_off[b]);
                rc = launch_attempt(b);
                if (rc) return rc;
                rc = launch_pivoted(b);
                if (rc) return rc;
                FRX_CUDA(cudaEventRecord(ctx->bin_join[b], ctx->bin_stream[b]));
                FRX_CUDA(cudaStreamWaitEvent(main_stream, ctx->bin_join[b], 0));
            }
            ctx->stream = main_stream;
            ctx->timing = restore.t;
            if (host_profile)
                fprintf(stderr, "frx_assemble_solve host us: wait for the host block %.0f, pass over %lld settings %.0f, upload + K5 launch %.0f, "
                                "launch geometry %.0f, bin chains %.0f\n", t_stage - t_begin, (long long)n_sys, t_pass - t_stage,
                        t_k5 - t_pass, t_geo - t_k5, t_now() - t_geo);
            return frx_timing_end(ctx);
        }
        if (use_ldl) {
            for (int b = kStreamBins - 1; b >= 0; --b) {
                if (!bin_count[b]) continue;
                rc = launch_attempt(b);
                if (rc) return rc;
            }
        }
        for (int b = kStreamBins - 1; b >= 0; --b) {    // widest bands first: the longest-running CTAs start first
            if (!bin_count[b]) continue;
            rc = launch_pivoted(b);
            if (rc) return rc;
        }
        return FRX_OK;
    }
    int smem_max = 0;
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    if (!d_dense && smem > (size_t)smem_max) {
        frx_set_error("system too large for shared memory: needs %zu bytes, device offers %d (n = %d, max resolution %d)",
                      smem, smem_max, n, max_res);
        return FRX_ERR_INVALID_ARG;
    }
    SolveParams p;
    p.n_sys = n_sys;
    p.n = n;
    p.rcw = (const double*)(ws + o_w);
    p.wptr = (const int64_t*)(img + o_wptr);
    p.h = d_par + 3 * n_sys;
    p.rhs = d_rhs;
    p.ldb = ldb;
    p.x = d_x;
    p.ldx = ldx;
    p.info = d_info;
    p.dense = d_dense;
    p.max_store = (max_store + 1) & ~1;
    p.a_len = a_len;
    const size_t dyn = smem;
    FRX_CUDA(cudaFuncSetAttribute(frx_k4_assemble_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int per_sm = 1;
    FRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frx_k4_assemble_solve, kSolveThreads, dyn));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > n_sys) grid = n_sys;
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_k4_assemble_solve<<<(unsigned)grid, kSolveThreads, dyn, ctx->stream>>>(p));
    return FRX_OK;
}

extern "C" int frx_assemble_solve(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                                  const double* h_res, const double* d_rhs, int64_t ldb, double* d_x, int64_t ldx,
                                  int32_t* d_info) {
    return solve_common(ctx, rule, n_sys, n, h_alpha, h_h, h_res, d_rhs, ldb, d_x, ldx, d_info, nullptr);
}

// Host-buffer form: right-hand sides and solutions in PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister, e.g. a
// pinned NumPy array or the shared segment of a multi-GPU sweep), used IN PLACE: every kernel reads a system's right-hand
// side once (coalesced, into shared memory) and writes its solution once, so the two transfers ride on the solves instead
// of framing them.  Pageable buffers are refused (the Python layer stages those through chunked copies).
extern "C" int frx_assemble_solve_host(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                                       const double* h_res, const double* h_rhs, int64_t ldb, double* h_x, int64_t ldx) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    if (n_sys == 0) return FRX_OK;
    FRX_REQUIRE(h_rhs && h_x, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    cudaPointerAttributes at_b, at_x;
    const bool ok_b = cudaPointerGetAttributes(&at_b, h_rhs) == cudaSuccess && at_b.type == cudaMemoryTypeHost && at_b.devicePointer;
    const bool ok_x = cudaPointerGetAttributes(&at_x, h_x) == cudaSuccess && at_x.type == cudaMemoryTypeHost && at_x.devicePointer;
    cudaGetLastError();
    if (!ok_b || !ok_x) {
        frx_set_error("frx_assemble_solve_host needs page-locked (pinned or registered) host buffers");
        return FRX_ERR_INVALID_ARG;
    }
    return solve_common(ctx, rule, n_sys, n, h_alpha, h_h, h_res, (const double*)at_b.devicePointer, ldb,
                        (double*)at_x.devicePointer, ldx, nullptr, nullptr, /*host_io=*/true);
}

extern "C" int frx_assemble_dense(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                                  const double* h_res, double* d_A) {
    FRX_REQUIRE(d_A != nullptr, FRX_ERR_INVALID_ARG, "d_A is NULL");
    return solve_common(ctx, rule, n_sys, n, h_alpha, h_h, h_res, nullptr, 0, nullptr, 0, nullptr, d_A);
}
