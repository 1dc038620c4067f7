// frx_history.cu -- K2: the "history sum" of the growing-history fractional operator,
//     S_i = sum_{d=0}^{i-1} c[d] * g_{i-d},   i = 0..N         (lower-triangular Toeplitz mat-vec)
//     y_i = mult[i] * (S_i + w0[i]*g_0 [+ cm1*g_{i+1}]),
// the O(N^2) loop that the reference performs in Python by rebuilding the stencil of
// Settings(alpha, i*h, i) for every node i and summing weight*value over its nodes
// (reference fractulus/test/integration/test_equation.py:48-57; weights equation.py:42-176).
//
// FP64 DFMA-pipe bound (N = 1e5: 5e9 FMA over 1.6 MB of input), so the design goal is one DFMA per
// (i, d) pair with as few other instructions as possible:
//   * work unit  = kTI rows x kTD distances.  Units are laid out in a 1-D sequence (alpha, field, row
//     tile, pass, distance chunk) and dealt in equal contiguous spans to a grid of
//     sm_count * kCtasPerSm persistent CTAs (stream-K), which balances the triangle exactly.
//   * inside a unit each thread owns kR consecutive rows (kR accumulators) and slides a register window
//     of 2*kR field values: per kR distances it issues kR*kR DFMAs for kR/2 LDS.128 of field values and
//     kR/2 broadcast LDS.128 of weights.
//   * weights c[] and the field window are staged global -> shared with 16-byte cp.async into a double
//     buffer; the field window is stored with 16 bytes of padding per 128-byte line so that the
//     per-thread 64-byte chunks are bank-conflict free.
//   * the triangular edge and the first column are handled by zero padding (a padded copy of the field
//     with g_j = 0 for j <= 0 is made by the prep kernel), so the inner loop has no predicates.
//   * a row tile finished inside one CTA is scaled and written directly; a tile split between CTAs
//     goes through per-CTA partial slots that a small fix-up kernel adds in CTA order -> bitwise
//     deterministic results, no atomics.
#include <math.h>

#include "frx_internal.h"

namespace {

constexpr int kR = 8;              // rows per thread
constexpr int kNT = 128;           // threads per CTA
constexpr int kTI = kR * kNT;      // rows per tile (1024)
constexpr int kTD = 256;           // distances per unit
constexpr int kCtasPerSm = 4;
constexpr int kPadF = kTI + kTD;   // zero padding in front of every padded field
constexpr int kChunkBytes = 8 * kR;
constexpr int kWinElems = kTI + kTD;
constexpr int kWinBytes = kWinElems * 8 + (kWinElems * 8 / 128) * 16;  // +16 B per 128 B line
constexpr int kStageBytes = kTD * 8 + kWinBytes;

static_assert(kTI % kTD == 0, "tile rows must be a multiple of the distance chunk");
static_assert((kTD / kR) % 2 == 0, "inner loop is unrolled by two blocks");

struct HistParams {
    int rule, n_pass;
    int64_t N, n_alpha, n_fields, n_tiles, units_per_prob, total_units;
    const double *cA, *cB;
    int64_t ldc;
    const double *w0, *mult;
    int64_t ldw;
    const double* cm1;
    const double* gp;  // padded fields [n_fields][n_pass][ldgp]
    int64_t ldgp;
    const double* g;   // original fields
    int64_t ldg;
    double* y;
    int64_t ldy;
    double* partial;   // [2 * grid][kTI]
};

__host__ __device__ inline int64_t tile_chunks(int64_t I, int64_t n_tiles, int64_t N) {
    // distances needed by the rows of tile I: d <= i_last - 1
    return I + 1 < n_tiles ? (I + 1) * (kTI / kTD) : (N + kTD - 1) / kTD;
}
__host__ __device__ inline int64_t tile_prefix(int64_t I, int n_pass) {  // units before tile I (all tiles < I are full)
    return (int64_t)n_pass * (kTI / kTD) * (I * (I + 1) / 2);
}

struct Cursor {
    int64_t prob, I, v, chunks, units;  // v in [0, units), units = n_pass * chunks
};

__device__ inline Cursor cursor_at(const HistParams& p, int64_t u) {
    Cursor c;
    c.prob = u / p.units_per_prob;
    int64_t w = u - c.prob * p.units_per_prob;
    const double K = (double)p.n_pass * (kTI / kTD);
    int64_t I = (int64_t)((sqrt(1.0 + 8.0 * (double)w / K) - 1.0) * 0.5);
    if (I > p.n_tiles - 1) I = p.n_tiles - 1;
    while (I > 0 && tile_prefix(I, p.n_pass) > w) --I;
    while (I + 1 < p.n_tiles && tile_prefix(I + 1, p.n_pass) <= w) ++I;
    c.I = I;
    c.v = w - tile_prefix(I, p.n_pass);
    c.chunks = tile_chunks(I, p.n_tiles, p.N);
    c.units = c.chunks * p.n_pass;
    return c;
}

__device__ inline void cursor_advance(const HistParams& p, Cursor& c) {
    if (++c.v == c.units) {
        c.v = 0;
        if (++c.I == p.n_tiles) {
            c.I = 0;
            ++c.prob;
        }
        c.chunks = tile_chunks(c.I, p.n_tiles, p.N);
        c.units = c.chunks * p.n_pass;
    }
}

__device__ inline void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ inline void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// stage the weights chunk and the field window of one unit into shared memory
__device__ inline void stage_unit(const HistParams& p, const Cursor& c, char* stage) {
    const int64_t pass = c.v / c.chunks;
    const int64_t d0 = (c.v - pass * c.chunks) * kTD;
    const int64_t a = c.prob / p.n_fields, f = c.prob - a * p.n_fields;
    const double* ctab = (pass ? p.cB : p.cA) + a * p.ldc + d0;
    const double* gsrc = p.gp + (f * p.n_pass + pass) * p.ldgp + kPadF + c.I * kTI - d0 - kTD;
    const int t = threadIdx.x;
    for (int q = t; q < kTD / 2; q += kNT) cp_async16(stage + 16 * q, ctab + 2 * q);
    char* win = stage + kTD * 8;
#pragma unroll
    for (int q = t; q < kWinElems / 2; q += kNT) {
        const int byte = 16 * q;
        cp_async16(win + byte + ((byte >> 7) << 4), gsrc + 2 * q);
    }
}

__device__ inline void load_chunk(double (&dst)[kR], const char* win, int chunk) {
    const int byte = kChunkBytes * chunk;
    const double2* src = reinterpret_cast<const double2*>(win + byte + ((byte >> 7) << 4));
#pragma unroll
    for (int q = 0; q < kR / 2; ++q) {
        double2 v = src[q];
        dst[2 * q] = v.x;
        dst[2 * q + 1] = v.y;
    }
}

// acc[r] += sum_s c[s] * window(r - s), window(k) = hi[k] for k >= 0, lo[k + kR] for k < 0
__device__ inline void fma_block(double (&acc)[kR], const double* cs, const double (&hi)[kR], const double (&lo)[kR]) {
    double c[kR];
#pragma unroll
    for (int q = 0; q < kR / 2; ++q) {
        double2 v = reinterpret_cast<const double2*>(cs)[q];
        c[2 * q] = v.x;
        c[2 * q + 1] = v.y;
    }
#pragma unroll
    for (int s = 0; s < kR; ++s) {
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const int k = r - s;
            acc[r] = fma(c[s], k >= 0 ? hi[k] : lo[k + kR], acc[r]);
        }
    }
}

__device__ inline void compute_unit(double (&acc)[kR], const char* stage) {
    const double* cs = reinterpret_cast<const double*>(stage);
    const char* win = stage + kTD * 8;
    const int c0 = threadIdx.x + kTD / kR;
    double hi[kR], lo[kR];
    load_chunk(hi, win, c0);
#pragma unroll 2
    for (int b = 0; b < kTD / kR; b += 2) {
        load_chunk(lo, win, c0 - b - 1);
        fma_block(acc, cs + b * kR, hi, lo);
        load_chunk(hi, win, c0 - b - 2 >= 0 ? c0 - b - 2 : 0);
        fma_block(acc, cs + (b + 1) * kR, lo, hi);
    }
}

// scale, add the first column / right-of-point terms and store the kR rows of this thread
__device__ inline void finish_rows(const HistParams& p, int64_t prob, int64_t I, const double (&acc)[kR]) {
    const int64_t a = prob / p.n_fields, f = prob - a * p.n_fields;
    const double* gf = p.g + f * p.ldg;
    const double g0 = gf[0];
    const double cm1 = p.rule == FRX_RULE_SIMPSON ? p.cm1[a] : 0.0;
    const int64_t ib = I * kTI + (int64_t)threadIdx.x * kR;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
        const int64_t i = ib + r;
        if (i > p.N) break;
        double v;
        if (i == 0) {
            v = 0.0;
        } else if (i == 1 && (p.rule == FRX_RULE_RECTANGLE || p.rule == FRX_RULE_SIMPSON)) {
            v = nan("");  // the reference cannot build this stencil
        } else {
            double s = fma(p.w0[a * p.ldw + i], g0, acc[r]);
            if (p.rule == FRX_RULE_SIMPSON && (i & 1)) s = fma(cm1, gf[i + 1], s);
            v = p.mult[a * p.ldw + i] * s;
        }
        p.y[prob * p.ldy + i] = v;
    }
}

__global__ void __launch_bounds__(kNT, kCtasPerSm) frx_k2_history_main(const HistParams p) {
    __shared__ __align__(128) char smem[2 * kStageBytes];
    const int64_t G = gridDim.x, k = blockIdx.x;
    const int64_t ubeg = k * p.total_units / G, uend = (k + 1) * p.total_units / G;
    if (ubeg >= uend) return;
    Cursor cur = cursor_at(p, ubeg), nxt = cur;
    double acc[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) acc[r] = 0.0;
    bool from_start = (cur.v == 0);
    stage_unit(p, cur, smem);
    cp_async_commit();
    for (int64_t u = ubeg; u < uend; ++u) {
        char* stage = smem + ((u - ubeg) & 1) * kStageBytes;
        if (u + 1 < uend) {
            cursor_advance(p, nxt);
            stage_unit(p, nxt, smem + ((u + 1 - ubeg) & 1) * kStageBytes);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        compute_unit(acc, stage);
        __syncthreads();
        const bool tile_done = (cur.v + 1 == cur.units), span_done = (u + 1 == uend);
        if (tile_done || span_done) {
            if (tile_done && from_start) {
                finish_rows(p, cur.prob, cur.I, acc);
            } else {
                // continuation of a tile begun by an earlier CTA -> slot 2k; tile begun here but not
                // finished -> slot 2k+1
                double* dst = p.partial + ((2 * k + (from_start ? 1 : 0)) * kTI + (int64_t)threadIdx.x * kR);
#pragma unroll
                for (int r = 0; r < kR; r += 2) *reinterpret_cast<double2*>(dst + r) = make_double2(acc[r], acc[r + 1]);
            }
#pragma unroll
            for (int r = 0; r < kR; ++r) acc[r] = 0.0;
            from_start = true;
        }
        cur = nxt;
    }
}

// One CTA per span boundary b = 1..G-1.  If the boundary cuts a row tile and is the first boundary
// inside that tile, this CTA adds the partial slots of every CTA that worked on the tile, in CTA
// order, and finishes the rows.
__global__ void __launch_bounds__(kNT) frx_k2_history_fixup(const HistParams p, int64_t G) {
    const int64_t b = (int64_t)blockIdx.x + 1;
    const int64_t ub = b * p.total_units / G;
    if (ub >= p.total_units || ub == (b - 1) * p.total_units / G) return;  // empty neighbour span
    const Cursor c = cursor_at(p, ub);
    if (c.v == 0) return;                       // boundary coincides with a tile start
    const int64_t tile_begin = ub - c.v, tile_end = tile_begin + c.units;
    // first boundary inside the tile <=> the previous non-empty span starts at or before the tile start
    int64_t prev = b - 1;
    while (prev > 0 && prev * p.total_units / G == (prev + 1) * p.total_units / G) --prev;
    if (prev * p.total_units / G > tile_begin) return;
    double acc[kR];
    const double* src = p.partial + (2 * prev + 1) * kTI + (int64_t)threadIdx.x * kR;
#pragma unroll
    for (int r = 0; r < kR; ++r) acc[r] = src[r];
    for (int64_t k = b; k < G; ++k) {
        const int64_t kb = k * p.total_units / G, ke = (k + 1) * p.total_units / G;
        if (kb >= tile_end) break;
        if (kb == ke) continue;
        src = p.partial + (2 * k) * kTI + (int64_t)threadIdx.x * kR;
#pragma unroll
        for (int r = 0; r < kR; ++r) acc[r] += src[r];
    }
    finish_rows(p, c.prob, c.I, acc);
}

// padded copies of the fields: gp[(f*n_pass + pass)*ldgp + kPadF + j] = g_j for 1 <= j <= N (and, for
// 'simpson', j of the pass's parity: pass 0 even nodes, pass 1 odd nodes), 0 elsewhere
__global__ void frx_k2_pad_fields(const double* __restrict__ g, int64_t ldg, int64_t N, int n_pass, int64_t ldgp,
                                  double* __restrict__ gp) {
    const int64_t fp = blockIdx.y, f = fp / n_pass, pass = fp - f * n_pass;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < ldgp; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = x - kPadF;
        double v = 0.0;
        if (j >= 1 && j <= N && (n_pass == 1 || (j & 1) == pass)) v = g[f * ldg + j];
        gp[fp * ldgp + x] = v;
    }
}

}  // namespace

int frx_launch_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                             double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                             double* d_cm1);

extern "C" int64_t frx_table_ldc(int64_t N) { return (N + 1 + kTD - 1) / kTD * kTD; }

extern "C" int frx_history(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                           int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, double* d_y, int64_t ldy) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_SIMPSON, FRX_ERR_UNKNOWN_RULE, "unknown approximation key");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30), FRX_ERR_INVALID_ARG, "N out of range");
    FRX_REQUIRE(n_alpha >= 0 && n_alpha <= 65535 && n_fields >= 0 && n_fields <= 32767, FRX_ERR_INVALID_ARG,
                "n_alpha (<= 65535) / n_fields (<= 32767) out of range");
    const int64_t need_g = N + 1 + (rule == FRX_RULE_SIMPSON ? 1 : 0);
    FRX_REQUIRE(n_g >= need_g && ldg >= n_g && ldy >= N + 1, FRX_ERR_INVALID_ARG,
                "field too short (needs N+1 nodes, N+2 for 'simpson') or leading dimension too small");
    if (n_alpha == 0 || n_fields == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_g && d_y, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_CUDA(cudaSetDevice(ctx->device));

    HistParams p;
    p.rule = rule;
    p.n_pass = rule == FRX_RULE_SIMPSON ? 2 : 1;
    p.N = N;
    p.n_alpha = n_alpha;
    p.n_fields = n_fields;
    p.n_tiles = (N + 1 + kTI - 1) / kTI;
    p.units_per_prob = tile_prefix(p.n_tiles - 1, p.n_pass) + p.n_pass * tile_chunks(p.n_tiles - 1, p.n_tiles, N);
    p.total_units = p.units_per_prob * n_alpha * n_fields;
    p.ldc = frx_table_ldc(N);
    p.ldw = (N + 1 + 1) / 2 * 2;
    p.ldgp = kPadF + p.n_tiles * kTI;
    int64_t G = (int64_t)ctx->sm_count * kCtasPerSm;
    if (G > p.total_units) G = p.total_units;

    // workspace carve-up (all offsets multiples of 256 bytes)
    size_t off = 0;
    auto carve = [&](size_t n_doubles) {
        size_t o = off;
        off += frx_align_up(n_doubles * sizeof(double), 256);
        return o;
    };
    const size_t o_cA = carve((size_t)n_alpha * p.ldc);
    const size_t o_cB = carve(p.n_pass == 2 ? (size_t)n_alpha * p.ldc : 0);
    const size_t o_w0 = carve((size_t)n_alpha * p.ldw);
    const size_t o_mult = carve((size_t)n_alpha * p.ldw);
    const size_t o_cm1 = carve((size_t)n_alpha);
    const size_t o_gp = carve((size_t)n_fields * p.n_pass * p.ldgp);
    const size_t o_part = carve((size_t)2 * G * kTI);
    int rc = frx_ws_reserve(ctx, off);
    if (rc) return rc;
    char* ws = (char*)ctx->ws;
    double* cA = (double*)(ws + o_cA);
    double* cB = p.n_pass == 2 ? (double*)(ws + o_cB) : nullptr;
    double* w0 = (double*)(ws + o_w0);
    double* mult = (double*)(ws + o_mult);
    double* cm1 = (double*)(ws + o_cm1);
    double* gp = (double*)(ws + o_gp);
    p.cA = cA;
    p.cB = cB;
    p.w0 = w0;
    p.mult = mult;
    p.cm1 = cm1;
    p.gp = gp;
    p.g = d_g;
    p.ldg = ldg;
    p.y = d_y;
    p.ldy = ldy;
    p.partial = (double*)(ws + o_part);

    rc = frx_launch_weights_table(ctx, rule, n_alpha, d_alpha, h, N, cA, cB, p.ldc, w0, mult, p.ldw, cm1);
    if (rc) return rc;
    {
        dim3 grid((unsigned)((p.ldgp + 255) / 256), (unsigned)(n_fields * p.n_pass));
        FRX_LAUNCH(ctx, FRX_K_PAD_FIELDS,
                   frx_k2_pad_fields<<<grid, 256, 0, ctx->stream>>>(d_g, ldg, N, p.n_pass, p.ldgp, gp));
    }
    FRX_LAUNCH(ctx, FRX_K_HISTORY_MAIN, frx_k2_history_main<<<(unsigned)G, kNT, 0, ctx->stream>>>(p));
    if (G > 1) {
        FRX_LAUNCH(ctx, FRX_K_HISTORY_FIXUP, frx_k2_history_fixup<<<(unsigned)(G - 1), kNT, 0, ctx->stream>>>(p, G));
    }
    return FRX_OK;
}

extern "C" int frx_history_host(frx_ctx* ctx, int rule, int64_t n_alpha, const double* h_alpha, double h, int64_t N,
                                int64_t n_fields, const double* h_g, int64_t ldg, int64_t n_g, double* h_y, int64_t ldy) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(h_alpha && h_g && h_y, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_REQUIRE(n_alpha >= 0 && n_fields >= 0 && N >= 1, FRX_ERR_INVALID_ARG, "bad sizes");
    for (int64_t a = 0; a < n_alpha; ++a) {
        int rc = frx_check_rule_alpha(rule, h_alpha[a]);
        if (rc) return rc;
    }
    if (n_alpha == 0 || n_fields == 0) return FRX_OK;
    FRX_REQUIRE(ldg >= n_g && ldy >= N + 1, FRX_ERR_INVALID_ARG, "leading dimension too small");
    FRX_CUDA(cudaSetDevice(ctx->device));
    // caller-facing device buffers live in their own allocation (the workspace is re-carved by frx_history)
    const size_t b_alpha = frx_align_up(sizeof(double) * n_alpha, 256);
    const size_t b_g = frx_align_up(sizeof(double) * (size_t)n_fields * ldg, 256);
    const size_t b_y = sizeof(double) * (size_t)n_alpha * n_fields * ldy;
    char* io = nullptr;
    FRX_CUDA(cudaMallocAsync((void**)&io, b_alpha + b_g + b_y, ctx->stream));
    double* d_alpha = (double*)io;
    double* d_g = (double*)(io + b_alpha);
    double* d_y = (double*)(io + b_alpha + b_g);
    int rc = FRX_OK;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_alpha, h_alpha, sizeof(double) * n_alpha, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_g, h_g, sizeof(double) * (size_t)n_fields * ldg, cudaMemcpyHostToDevice, ctx->stream)) !=
            cudaSuccess) {
        frx_set_error("host->device copy failed: %s", cudaGetErrorString(e));
        rc = FRX_ERR_CUDA;
    }
    if (!rc) rc = frx_history(ctx, rule, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, n_g, d_y, ldy);
    if (!rc && (e = cudaMemcpyAsync(h_y, d_y, b_y, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) {
        frx_set_error("device->host copy failed: %s", cudaGetErrorString(e));
        rc = FRX_ERR_CUDA;
    }
    cudaFreeAsync(io, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e != cudaSuccess) {
        frx_set_error("stream synchronize failed: %s", cudaGetErrorString(e));
        rc = FRX_ERR_CUDA;
    }
    return rc;
}
