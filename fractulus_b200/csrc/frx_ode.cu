// frx_ode.cu -- EXTENSION (BASELINE cfg 4 "fractional ODE ... solve"): invert the growing-history
// operator.  Given f_1..f_N and u_0, find u_1..u_N with
//     mult[i] * ( w0[i]*u_0 + sum_{j=1..i} c[i-j]*u_j ) = f_i ,        i = 1..N
// i.e. the values whose fractional derivative (rows = reference stencils of Settings(alpha, i*h, i), call pattern of
// reference test/integration/test_equation.py:51-66) is f: one step of an implicit fractional ODE / Volterra solve.
// The reference has no such solver (it has no solver at all): parity is defined against forward substitution with the
// reference's own weights (oracle/frx_oracle.c orc_history_solve).
//
// With b_i = f_i/mult[i] - w0[i]*u_0 this is a lower-triangular Toeplitz system T(c) x = b.  The inverse of a
// lower-triangular Toeplitz matrix is lower-triangular Toeplitz: T(c)^-1 = T(v) with the power series v = 1/c(z).
//
// Default algorithm (tensor cores): v is computed to full length by Newton doubling on the series,
//     v_{2m} = v_m - v_m * (c * v_m - 1)   mod z^{2m},
// where each product of series truncated to its first rows IS a lower-triangular Toeplitz mat-vec, i.e. the history
// kernel (frx_toeplitz_raw, FP64 DMMA, ~32 TFLOP/s) with a caller-supplied sequence:
//     e = T(c) [v_m; 0]  (2m rows; e = delta + z^m r),    v[m..2m) = -T(v_m) r,
// starting from the first 1024 coefficients (kernel S1, sequential recurrence), and the solve itself is one more
// product x = T(v) b.  For the M-matrix-like sequences of the quadrature rules (c_0 > 0 > c_d, v_d > 0) neither product
// cancels.  ~1.33 N^2 FMA in total instead of N^2/2, but all of it on the tensor-core kernel and with O(log N) launches
// instead of 2N/1024 dependent ones.
//
// FRX_ODE_VARIANT=1 selects the first version: blocked forward substitution with block size NB = 1024, the NB x NB
// diagonal block inverted once (S1) and applied as a small Toeplitz mat-vec (S2), all later rows updated after every
// block (S3); N/NB sequential stages of two launches each in plain DFMA kernels.
#include <math.h>
#include <stdlib.h>

#include "frx_internal.h"

int frx_launch_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                             double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                             double* d_cm1);
int frx_toeplitz_raw(frx_ctx* ctx, int64_t n_seq, const double* d_c, int64_t ldc, int64_t N, int64_t n_fields, int paired,
                     const double* d_g, int64_t ldg, double scale, double* d_y, int64_t ldy, int64_t n_cols);

namespace {

constexpr int kNB = 1024;

// S0: b[i'] = f[i'+1]/mult[i'+1] - w0[i'+1]*u0, i' = 0..N-1 (shifted rows)
__global__ void frx_ode_rhs(int64_t N, const double* __restrict__ f, int64_t ldf, const double* __restrict__ u0,
                            const double* __restrict__ w0, const double* __restrict__ mult, int64_t ldw,
                            double* __restrict__ b, int64_t ldb, int64_t n_fill) {
    const int64_t a = blockIdx.y;
    for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < n_fill; ip += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (ip < N) v = f[a * ldf + ip + 1] / mult[a * ldw + ip + 1] - w0[a * ldw + ip + 1] * u0[a];
        b[a * ldb + ip] = v;
    }
}

// S1: v = first NB coefficients of 1 / c(z), i.e. c * v = delta.  Thread t accumulates delta_t - sum_s c_{t-s} v_s.
// Blocks of 32 coefficients: warp 0 finds the first 32 by the sequential recurrence (shuffles, no barriers); after
// that every block is (a) a rank-32 update of all later accumulators and (b) a 32 x 32 diagonal solve, which is a
// small Toeplitz mat-vec with the first 32 coefficients of v themselves (the leading block's inverse) -- 32 independent
// shuffles instead of 32 dependent steps.  One barrier per block.
__global__ void __launch_bounds__(kNB) frx_ode_inverse_series(const double* __restrict__ c, int64_t ldc,
                                                              double* __restrict__ v, int64_t ldv, int nb) {
    __shared__ double sc[kNB], sv32[32], sblk[2][32];
    const int64_t a = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    sc[t] = t < nb ? c[a * ldc + t] : 0.0;
    __syncthreads();
    const double c0 = sc[0], rc0 = 1.0 / c0;
    double acc = t == 0 ? 1.0 : 0.0;
    double mine = 0.0;
    if (warp == 0) {
        for (int s = 0; s < 32; ++s) {
            double q = acc * rc0;
            q = fma(fma(-c0, q, acc), rc0, q);     // acc / c0 to within an ulp without a divide on the critical path
            const double vs = __shfl_sync(0xffffffffu, q, s);
            if (lane == s) mine = vs;
            if (lane > s) acc = fma(-sc[lane - s], vs, acc);
        }
        sv32[lane] = mine;
        sblk[0][lane] = mine;
    }
    __syncthreads();
    for (int B = 0; B < kNB / 32 - 1; ++B) {
        const double* vb = sblk[B & 1];
        if (t >= 32 * (B + 1)) {   // rank-32 update with the coefficients of block B
            const double* cw = sc + (t - 32 * B);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int s = 0; s < 32; s += 2) {
                s0 = fma(cw[-s], vb[s], s0);
                s1 = fma(cw[-s - 1], vb[s + 1], s1);
            }
            acc -= s0 + s1;
        }
        if (warp == B + 1) {       // diagonal solve of block B+1: x = T(v[0..31]) acc
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int s = 0; s < 32; s += 2) {
                const double a0 = __shfl_sync(0xffffffffu, acc, s), a1 = __shfl_sync(0xffffffffu, acc, s + 1);
                if (lane >= s) x0 = fma(sv32[lane - s], a0, x0);
                if (lane >= s + 1) x1 = fma(sv32[lane - s - 1], a1, x1);
            }
            mine = x0 + x1;
            sblk[(B + 1) & 1][lane] = mine;
        }
        __syncthreads();
    }
    if (t < ldv) v[a * ldv + t] = t < nb ? mine : 0.0;
}

// S2: x_I = T(v) b_I for block I (rows I*NB .. I*NB+NB-1), one CTA per alpha
__global__ void __launch_bounds__(kNB) frx_ode_diag_apply(int64_t N, int64_t I, const double* __restrict__ v,
                                                          const double* __restrict__ b, int64_t ldb,
                                                          double* __restrict__ x, int64_t ldx) {
    __shared__ double sv[kNB], sb[kNB];
    const int64_t a = blockIdx.x, r0 = I * kNB;
    const int t = threadIdx.x;
    sv[t] = v[a * kNB + t];
    sb[t] = r0 + t < N ? b[a * ldb + r0 + t] : 0.0;
    __syncthreads();
    // threads t and NB-1-t share the work of two rows so that every thread does ~NB/2 FMAs
    double acc = 0.0;
    for (int s = 0; s <= t; ++s) acc = fma(sv[t - s], sb[s], acc);
    if (r0 + t < N) x[a * ldx + r0 + t + 1] = acc;      // x_{i'} = u_{i'+1}
}

// S3: rows i' >= (I+1)*NB: b[i'] -= sum_{s<NB} c[i' - I*NB - s] * x[I*NB + s].  One CTA = 256 threads x 4 consecutive rows:
// per 4 block entries s a thread needs 7 table values (a sliding window of distances) and 4 x values for 16 FMAs.
constexpr int kS3Rows = 4;
constexpr int kS3Tile = 256 * kS3Rows;
__global__ void __launch_bounds__(256) frx_ode_trailing_update(int64_t N, int64_t I, const double* __restrict__ c, int64_t ldc,
                                                               const double* __restrict__ x, int64_t ldx,
                                                               double* __restrict__ b, int64_t ldb) {
    __shared__ double sx[kNB];
    __shared__ double sc[kNB + kS3Tile];
    const int64_t a = blockIdx.y, c0row = (I + 1) * kNB + (int64_t)blockIdx.x * kS3Tile;   // first row of this CTA
    const int t = threadIdx.x;
    const int64_t xb = I * kNB;
    for (int s = t; s < kNB; s += 256) sx[s] = xb + s < N ? x[a * ldx + xb + s + 1] : 0.0;
    // distances needed: row - (xb + s), s = 0..NB-1, rows c0row..c0row+Tile-1 -> [c0row - xb - NB + 1, c0row - xb + Tile - 1]
    const int64_t dbase = c0row - xb - kNB + 1;
    for (int q = t; q < kNB + kS3Tile - 1; q += 256) {
        const int64_t d = dbase + q;
        sc[q] = (d >= 0 && d < N) ? c[a * ldc + d] : 0.0;
    }
    __syncthreads();
    // thread t owns rows c0row + 4t .. +3; c[row - xb - s] = sc[(row - c0row) + NB - 1 - s]
    const int r0 = kS3Rows * t;
    double acc[kS3Rows] = {0.0, 0.0, 0.0, 0.0};
    for (int s = 0; s < kNB; s += 4) {
        // window w[k] = sc[r0 + NB - 1 - s - 3 + k], k = 0..6: row r, entry s+e uses w[3 + r - e]
        double w[7], xv[4];
        const double* wp = sc + r0 + kNB - 4 - s;
#pragma unroll
        for (int k = 0; k < 7; ++k) w[k] = wp[k];
#pragma unroll
        for (int e = 0; e < 4; ++e) xv[e] = sx[s + e];
#pragma unroll
        for (int r = 0; r < kS3Rows; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[r] = fma(w[3 + r - e], xv[e], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kS3Rows; ++r) {
        const int64_t ip = c0row + r0 + r;
        if (ip < N) b[a * ldb + ip] -= acc[r];
    }
}

}  // namespace

extern "C" int frx_history_solve(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                                 const double* d_f, int64_t ldf, const double* d_u0, double* d_u, int64_t ldu) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    // rules whose weight of the evaluation point itself, c[0], is non-zero ('rectangle': c[0] = 0, 'simpson': two sequences)
    FRX_REQUIRE(rule == FRX_RULE_CAPUTO || rule == FRX_RULE_TRAPEZOIDAL || rule == FRX_RULE_GRUNWALD_LETNIKOV, FRX_ERR_DOMAIN,
                "frx_history_solve offers the 'caputo', 'trapezoidal' and 'grunwald_letnikov' operators");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30) && n_alpha >= 0 && n_alpha <= 65535, FRX_ERR_INVALID_ARG, "bad N / n_alpha");
    FRX_REQUIRE(ldf >= N + 1 && ldu >= N + 1, FRX_ERR_INVALID_ARG, "leading dimension too small (N+1 nodes)");
    if (n_alpha == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_f && d_u0 && d_u, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    static const int variant = getenv("FRX_ODE_VARIANT") ? atoi(getenv("FRX_ODE_VARIANT")) : 0;
    const int64_t ldc = frx_table_ldc(N), ldw = N + 1, nblk = (N + kNB - 1) / kNB, ldb = nblk * kNB;
    // sequences of the Newton variant carry one pad element in front (the history kernel's node 0) and are a whole
    // number of 1024-blocks long
    const int64_t ldv = ldb + 8;
    // the table launcher and the history kernel use the workspace: everything that must survive them lives in the io buffer
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += frx_align_up(bytes, 256); return o; };
    const size_t o_c = carve(sizeof(double) * (size_t)n_alpha * ldc), o_w0 = carve(sizeof(double) * (size_t)n_alpha * ldw),
                 o_m = carve(sizeof(double) * (size_t)n_alpha * ldw), o_b = carve(sizeof(double) * (size_t)n_alpha * ldv),
                 o_v = carve(sizeof(double) * (size_t)n_alpha * ldv), o_e = carve(sizeof(double) * (size_t)n_alpha * ldv);
    int rc = frx_io_reserve(ctx, off);
    if (rc) return rc;
    char* io = (char*)ctx->io;
    double *c = (double*)(io + o_c), *w0 = (double*)(io + o_w0), *mult = (double*)(io + o_m), *b = (double*)(io + o_b),
           *v = (double*)(io + o_v), *e = (double*)(io + o_e);
    rc = frx_launch_weights_table(ctx, rule, n_alpha, d_alpha, h, N, c, nullptr, ldc, w0, mult, ldw, nullptr);
    if (rc) return rc;
    const int nb = (int)(N < kNB ? N : kNB);
    FRX_CUDA(cudaMemcpy2DAsync(d_u, sizeof(double) * ldu, d_u0, sizeof(double), sizeof(double), (size_t)n_alpha,
                               cudaMemcpyDeviceToDevice, ctx->stream));   // u[a][0] = u0[a]
    if (variant == 1) {
        FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_rhs<<<dim3((unsigned)((ldb + 255) / 256), (unsigned)n_alpha), 256, 0, ctx->stream>>>(
                                         N, d_f, ldf, d_u0, w0, mult, ldw, b, ldb, ldb));
        FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_inverse_series<<<(unsigned)n_alpha, kNB, 0, ctx->stream>>>(c, ldc, v, kNB, nb));
        for (int64_t I = 0; I < nblk; ++I) {
            FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_diag_apply<<<(unsigned)n_alpha, kNB, 0, ctx->stream>>>(N, I, v, b, ldb, d_u, ldu));
            const int64_t rows_left = N - (I + 1) * kNB;
            if (rows_left > 0) {
                dim3 grid((unsigned)((rows_left + kS3Tile - 1) / kS3Tile), (unsigned)n_alpha);
                FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_trailing_update<<<grid, 256, 0, ctx->stream>>>(N, I, c, ldc, d_u, ldu, b, ldb));
            }
        }
        return FRX_OK;
    }
    // Newton variant.  Layout of b, v, e: element 0 is the pad ("node 0"), entry k of the sequence sits at [1 + k].
    FRX_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)n_alpha * ldv, ctx->stream));
    FRX_CUDA(cudaMemsetAsync(e, 0, sizeof(double) * (size_t)n_alpha * ldv, ctx->stream));
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_rhs<<<dim3((unsigned)((ldb + 255) / 256), (unsigned)n_alpha), 256, 0, ctx->stream>>>(
                                     N, d_f, ldf, d_u0, w0, mult, ldw, b + 1, ldv, ldb));
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_inverse_series<<<(unsigned)n_alpha, kNB, 0, ctx->stream>>>(c, ldc, v + 1, ldv, nb));
    for (int64_t m = nb; m < N;) {
        const int64_t m2 = 2 * m < N ? 2 * m : N;
        // e[1 + i] = (c * v_m)[i], m <= i < m2   (v is zero beyond m: a rectangular block of the product)
        rc = frx_toeplitz_raw(ctx, n_alpha, c, ldc, m2, 1, 1, v, ldv, 1.0, e, ldv, getenv("FRX_ODE_NORECT") ? 0 : m);
        if (rc) return rc;
        // v[m + i] = -(v_m * r)[i], r[i] = e[1 + m + i], i < m2 - m: the kernel writes rows 1.. of the output at v + m
        rc = frx_toeplitz_raw(ctx, n_alpha, v + 1, ldv, m2 - m, 1, 1, e + m, ldv, -1.0, v + m, ldv, 0);
        if (rc) return rc;
        m = m2;
    }
    // u[1 + i] = (v * b)[i]
    return frx_toeplitz_raw(ctx, n_alpha, v + 1, ldv, N, 1, 1, b, ldv, 1.0, d_u, ldu, 0);
}
