// frx_ode.cu -- EXTENSION (BASELINE cfg 4 "fractional ODE ... solve"): invert the growing-history
// operator.  Given f_1..f_N and u_0, find u_1..u_N with
//     mult[i] * ( w0[i]*u_0 + sum_{j=1..i} c[i-j]*u_j ) = f_i ,        i = 1..N
// i.e. the values whose fractional derivative (rows = reference stencils of Settings(alpha, i*h, i), call pattern of
// reference test/integration/test_equation.py:51-66) is f: one step of an implicit fractional ODE / Volterra solve.
// The reference has no such solver (it has no solver at all): parity is defined against forward substitution with the
// reference's own weights (oracle/frx_oracle.c orc_history_solve).
//
// With b_i = f_i/mult[i] - w0[i]*u_0 this is a lower-triangular Toeplitz system T x = b, T[i][j] = c[i-j].  Blocked
// forward substitution with block size NB = 1024:
//   * the inverse of a lower-triangular Toeplitz matrix is lower-triangular Toeplitz, so the NB x NB diagonal block is
//     inverted once per alpha as a power series v (c * v = delta, kernel S1) and every diagonal solve is a small
//     Toeplitz mat-vec x_I = T(v) b_I (kernel S2) instead of NB sequential steps;
//   * after block I is known, all later rows are updated at once, b_i -= sum_s c[i - (I*NB+s)] x_{I*NB+s} (kernel S3,
//     one thread per row, c window and x_I staged in shared memory).
// N/NB sequential stages of two launches each; total work N^2/2 FMA like the forward operator.  First version: plain
// DFMA kernels, not tuned.
#include <math.h>

#include "frx_internal.h"

int frx_launch_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                             double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                             double* d_cm1);

namespace {

constexpr int kNB = 1024;

// S0: b[i'] = f[i'+1]/mult[i'+1] - w0[i'+1]*u0, i' = 0..N-1 (shifted rows)
__global__ void frx_ode_rhs(int64_t N, const double* __restrict__ f, int64_t ldf, const double* __restrict__ u0,
                            const double* __restrict__ w0, const double* __restrict__ mult, int64_t ldw,
                            double* __restrict__ b, int64_t ldb) {
    const int64_t a = blockIdx.y;
    for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < ldb; ip += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (ip < N) v = f[a * ldf + ip + 1] / mult[a * ldw + ip + 1] - w0[a * ldw + ip + 1] * u0[a];
        b[a * ldb + ip] = v;
    }
}

// S1: v = first NB coefficients of 1 / c(z): v_0 = 1/c_0, v_t = -(sum_{s=1..t} c_s v_{t-s}) / c_0 (right-looking)
__global__ void __launch_bounds__(kNB) frx_ode_inverse_series(const double* __restrict__ c, int64_t ldc,
                                                              double* __restrict__ v, int nb) {
    __shared__ double sv;
    const int64_t a = blockIdx.x;
    const int t = threadIdx.x;
    const double* ca = c + a * ldc;
    const double c0 = ca[0];
    if (t >= nb) v[a * kNB + t] = 0.0;
    double acc = t == 0 ? 1.0 : 0.0;           // delta_t - sum_{s<t..} c_{t-s} v_s accumulates here
    for (int s = 0; s < nb; ++s) {
        if (t == s) sv = acc / c0;
        __syncthreads();
        const double vs = sv;
        if (t == s) v[a * kNB + s] = vs;
        if (t > s && t < nb) acc = fma(-ca[t - s], vs, acc);
        __syncthreads();
    }
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

// S3: rows i' >= (I+1)*NB: b[i'] -= sum_{s<NB} c[i' - I*NB - s] * x[I*NB + s]
__global__ void __launch_bounds__(256) frx_ode_trailing_update(int64_t N, int64_t I, const double* __restrict__ c, int64_t ldc,
                                                               const double* __restrict__ x, int64_t ldx,
                                                               double* __restrict__ b, int64_t ldb) {
    __shared__ double sx[kNB];
    __shared__ double sc[kNB + 256];
    const int64_t a = blockIdx.y, c0row = (I + 1) * kNB + (int64_t)blockIdx.x * 256;   // first row of this CTA
    const int t = threadIdx.x;
    const int64_t xb = I * kNB;
    for (int s = t; s < kNB; s += 256) sx[s] = xb + s < N ? x[a * ldx + xb + s + 1] : 0.0;
    // distances needed: row - (xb + s), s = 0..NB-1, rows c0row..c0row+255 -> [c0row - xb - NB + 1, c0row - xb + 255]
    const int64_t dbase = c0row - xb - kNB + 1;
    for (int q = t; q < kNB + 255; q += 256) {
        const int64_t d = dbase + q;
        sc[q] = (d >= 0 && d < N) ? c[a * ldc + d] : 0.0;
    }
    __syncthreads();
    const int64_t ip = c0row + t;
    if (ip >= N) return;
    double acc = 0.0;
    // c[ip - xb - s] = sc[(ip - c0row) + (NB - 1 - s)]
    const double* cw = sc + t + kNB - 1;
#pragma unroll 8
    for (int s = 0; s < kNB; ++s) acc = fma(cw[-s], sx[s], acc);
    b[a * ldb + ip] -= acc;
}

}  // namespace

extern "C" int frx_history_solve(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                                 const double* d_f, int64_t ldf, const double* d_u0, double* d_u, int64_t ldu) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule == FRX_RULE_CAPUTO || rule == FRX_RULE_TRAPEZOIDAL, FRX_ERR_DOMAIN,
                "frx_history_solve offers the 'caputo' and 'trapezoidal' operators");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30) && n_alpha >= 0 && n_alpha <= 65535, FRX_ERR_INVALID_ARG, "bad N / n_alpha");
    FRX_REQUIRE(ldf >= N + 1 && ldu >= N + 1, FRX_ERR_INVALID_ARG, "leading dimension too small (N+1 nodes)");
    if (n_alpha == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_f && d_u0 && d_u, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_CUDA(cudaSetDevice(ctx->device));
    const int64_t ldc = frx_table_ldc(N), ldw = N + 1, nblk = (N + kNB - 1) / kNB, ldb = nblk * kNB;
    // the table launcher uses the start of the workspace for its scalars: tables live in the io buffer
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += frx_align_up(bytes, 256); return o; };
    const size_t o_c = carve(sizeof(double) * (size_t)n_alpha * ldc), o_w0 = carve(sizeof(double) * (size_t)n_alpha * ldw),
                 o_m = carve(sizeof(double) * (size_t)n_alpha * ldw), o_b = carve(sizeof(double) * (size_t)n_alpha * ldb),
                 o_v = carve(sizeof(double) * (size_t)n_alpha * kNB);
    int rc = frx_io_reserve(ctx, off);
    if (rc) return rc;
    char* io = (char*)ctx->io;
    double *c = (double*)(io + o_c), *w0 = (double*)(io + o_w0), *mult = (double*)(io + o_m), *b = (double*)(io + o_b),
           *v = (double*)(io + o_v);
    rc = frx_launch_weights_table(ctx, rule, n_alpha, d_alpha, h, N, c, nullptr, ldc, w0, mult, ldw, nullptr);
    if (rc) return rc;
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_rhs<<<dim3((unsigned)((ldb + 255) / 256), (unsigned)n_alpha), 256, 0, ctx->stream>>>(
                                     N, d_f, ldf, d_u0, w0, mult, ldw, b, ldb));
    const int nb = (int)(N < kNB ? N : kNB);
    FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_inverse_series<<<(unsigned)n_alpha, kNB, 0, ctx->stream>>>(c, ldc, v, nb));
    FRX_CUDA(cudaMemcpy2DAsync(d_u, sizeof(double) * ldu, d_u0, sizeof(double), sizeof(double), (size_t)n_alpha,
                               cudaMemcpyDeviceToDevice, ctx->stream));   // u[a][0] = u0[a]
    for (int64_t I = 0; I < nblk; ++I) {
        FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_diag_apply<<<(unsigned)n_alpha, kNB, 0, ctx->stream>>>(N, I, v, b, ldb, d_u, ldu));
        const int64_t rows_left = N - (I + 1) * kNB;
        if (rows_left > 0) {
            dim3 grid((unsigned)((rows_left + 255) / 256), (unsigned)n_alpha);
            FRX_LAUNCH(ctx, FRX_K_SOLVE, frx_ode_trailing_update<<<grid, 256, 0, ctx->stream>>>(N, I, c, ldc, d_u, ldu, b, ldb));
        }
    }
    return FRX_OK;
}
