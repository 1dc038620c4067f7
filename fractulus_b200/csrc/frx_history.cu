// frx_history.cu -- K2 / K3: the "history sum" of the growing-history fractional operator,
//     S_i = sum_{j=1}^{i} c[i-j] * g_j,                       i = 1..N   (lower-triangular Toeplitz mat-vec)
//     y_i = mult[i] * (S_i + w0[i]*g_0 [+ cm1*g_{i+1}]),      y_0 = 0,
// the O(N^2) loop that the reference performs in Python by rebuilding the stencil of Settings(alpha, i*h, i) for every
// node i and summing weight*value over its nodes (reference fractulus/test/integration/test_equation.py:48-57; weights
// equation.py:42-176).  FP64-compute bound (N = 1e5: 5e9 FMA over 1.6 MB of input).
//
// Common machinery (both kernel variants)
//   * rows and columns are tiled in shifted coordinates i' = i-1, j' = j-1: exactly N x N, lower triangular.
//   * work unit = TI rows x TD columns (tensor variant) / distances (DFMA variant).  Units are laid out in a 1-D sequence
//     (alpha, field, [row parity], row tile, pass, chunk) and dealt in contiguous spans, cut at sub-unit granularity, to
//     a grid of exactly sm_count * CTAS_PER_SM persistent CTAs (stream-K); dynamic shared memory is sized so that
//     exactly CTAS_PER_SM CTAs fit on an SM (all CTAs resident, evenly spread).
//   * tables and field chunks are staged global -> shared into a double buffer: by two cp.async.bulk copies per unit (TMA
//     engine, completion on an mbarrier) in the default geometry, by 16-byte cp.async in the others.
//   * the triangular edge is handled by zero padding (tables have zeros for negative distances, the prepare kernel makes
//     a zero-padded copy of every field), so the inner loops have no predicates.
//   * a row tile finished inside one CTA is scaled and written directly; a tile split between CTAs goes through per-CTA
//     partial slots, and the LAST CTA to arrive (atomic ticket per tile) adds the slots in CTA order -> bitwise
//     deterministic results, no floating-point atomics, no second kernel.
//   * per call: K0 (per-alpha scalars) -> prepare (K1 weight tables + field padding + tickets) -> history kernel, chained
//     by programmatic dependent launch.
//   * epilogue options: right-sided / Riesz-Caputo combination (frx_history_sided), stores into every NVLink peer's copy
//     of a gathered result array (frx_history_gather).
//   * 'simpson': two interleaved weight sequences -> two half-size problems (row parity) of two passes each over
//     decimated tables and fields (see frx_table_block), same N^2/2 FMA as the other rules.
// Tensor variant MGeo (default MGeo<4,1024,2,true>: units of 1024 rows x 1024 columns, 2 CTAs of 4 warps per SM, units staged
//   by two cp.async.bulk copies (TMA engine) against an mbarrier): block-Toeplitz
//   formulation on mma.m8n8k4.f64 (DMMA), see MGeo below.
// DFMA variant Geo (FRX_K2_VARIANT=0/1, kept for comparison): each thread owns R consecutive rows and slides a register
//   window of 2R field values: per R distances R*R DFMAs for R/2 LDS.128 of field values (chunk stride 8R+16 bytes ->
//   conflict-free) and R/2 broadcast LDS.128 of weights.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "frx_internal.h"
#include "frx_table.cuh"

// defined in frx_weights.cu
int frx_launch_alpha_scalars(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, void* d_scalars);
size_t frx_alpha_scalars_bytes(void);
int frx_launch_gl_table(frx_ctx* ctx, int64_t n_alpha, const double* d_alpha, double h, int64_t N, const void* d_scalars,
                        int c_pad, int c_perm, double* d_cA, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                        double* d_cm1);

namespace {

constexpr int kNT = 128;      // threads per CTA
constexpr int kLdcQuantum = 256;
// partial slots loaded together in the split-tile fix-up: as many as the register budget of the geometry allows (with
// 4 CTAs per SM, i.e. 128 registers, more than 2 costs the main loop its schedule)
#ifndef FRX_FIXUP_BATCH
#define FRX_FIXUP_BATCH 8
#endif
template <class G>
struct FixupBatch {
    static constexpr int value = G::CTAS_PER_SM <= 2 ? FRX_FIXUP_BATCH : 2;
};

template <int R_, int TD_, int CTAS_>
struct Geo {
    static constexpr int R = R_;                 // rows per thread
    static constexpr int TD = TD_;               // distances per unit
    static constexpr int CTAS_PER_SM = CTAS_;
    static constexpr int TI = R * kNT;           // rows per tile
    static constexpr int PADF = TI + TD;         // zero padding in front of every padded field
    static constexpr int CHUNK_BYTES = 8 * R;
    static constexpr int CHUNK_STRIDE = CHUNK_BYTES + 16;
    static constexpr int WIN_ELEMS = TI + TD;
    static constexpr int WIN_CHUNKS = WIN_ELEMS / R;
    static constexpr int WIN_BYTES = WIN_CHUNKS * CHUNK_STRIDE;
    static constexpr int STAGE_BYTES = (TD * 8 + WIN_BYTES + 127) / 128 * 128;
    static constexpr int SUBS = 1;               // a unit is not divisible between CTAs
    static constexpr int C_PAD = 0;              // c tables: natural order, no front padding
    static constexpr int C_PERM = 0;
    static constexpr bool TENSOR = false;
    static constexpr bool BULK = false;
    static constexpr int WBUF_BYTES = 0;         // no epilogue-operand prefetch
    static_assert(TI % TD == 0, "tile rows must be a multiple of the distance chunk");
    static_assert(STAGE_BYTES >= TI * 8, "the epilogue stages one tile of rows in a stage buffer");
    static_assert((TD / R) % 2 == 0, "inner loop is unrolled by two blocks");
    static_assert(kNT == FRX_TAB_THREADS && kLdcQuantum % TD == 0, "table padding quantum must cover a chunk");
};


// Geometry of the DMMA (FP64 tensor core) variant.  Rows are grouped in blocks of 8 (i = 8p + r) and columns
// in blocks of 8 (j = 8m + r - s): for a fixed column block m
//     Y'[p][r] += sum_{s=0..7} c[8(p-m) + s] * g[8m + r - s]          (exactly the terms d = i - j, no waste)
// is a (P x 8) x (8 x 8) product, i.e. two mma.m8n8k4 per 8 block-rows p: M = p, K = s, N = r.  A warp owns
// T interleaved 8-row-block tiles (tile t holds p = P0 + t + T*k, k = 0..7), so the A fragment of tile t at
// column block m is the A fragment of tile t-1 at block m-1: ONE new LDS.128 per column block feeds 2T DMMAs.
template <int T_, int TD_, int CTAS_, bool BULK_ = false>
struct MGeo {
    // BULK: units are staged with cp.async.bulk (the TMA engine's 1-D copies: one 64-byte copy per window row, one copy
    // for the field chunk) completing on an mbarrier, instead of 16-byte cp.async per thread
    static constexpr bool BULK = BULK_;
    static constexpr int T = T_;                 // interleaved tiles per warp
    static constexpr int R = 2 * T;              // accumulator doubles per thread
    static constexpr int TD = TD_;               // columns per unit
    static constexpr int MC = TD / 8;            // column blocks per unit
    static constexpr int SUBS = MC / T_;         // groups of T column blocks: the granularity at which spans are cut
    static constexpr int CTAS_PER_SM = CTAS_;
    static constexpr int WARPS = kNT / 32;
    static constexpr int TI = WARPS * 64 * T;    // rows per tile
    static constexpr int PADF = TI + TD;         // zero padding in front of padded fields and c tables
    static constexpr int NU = (TI + TD) / 8;     // 8-distance rows of the c window
    // cp.async staging: 64 B of data + 16 B of padding per window row (LDS.128 of 8 lanes x (k, k+1) conflict free for
    // T = 4).  BULK staging: dense 64-byte rows, the conflict-free placement comes from the table's own row swizzle
    // (frx_table_slot, c_perm = 2): the window is ONE bulk copy.
    static constexpr int ROW_STRIDE = BULK_ ? 64 : 80;
    static constexpr int CWIN_BYTES = NU * ROW_STRIDE;
    static constexpr int GCH_ELEMS = TD + 8;     // 8-element halo in front (B fragments reach back 7)
    static constexpr int STAGE_BYTES = (CWIN_BYTES + GCH_ELEMS * 8 + 64 + 127) / 128 * 128;   // + 64: see compute_unit_mma
    static constexpr int C_PAD = PADF;
    static constexpr int C_PERM = BULK_ ? 2 : 1;
    static constexpr bool TENSOR = true;
    // epilogue operands (w0 and mult of the tile's rows) are prefetched into shared memory together with the unit that
    // ends a tile or a span, one buffer per stage, when the shared-memory budget allows (<= 2 CTAs per SM)
    static constexpr int WBUF_BYTES = CTAS_ <= 2 ? 2 * TI * 8 : 0;
    static_assert(TI % TD == 0 && TD % 8 == 0 && MC % T == 0, "bad tile geometry");
    static_assert(STAGE_BYTES >= TI * 8, "the epilogue stages one tile of rows in a stage buffer");
    static_assert(T == 4, "ROW_STRIDE = 80 is conflict free for T = 4 only");
};

struct HistParams {
    int rule, n_pass;
    // n_sub sub-problems per (alpha, field): 1, or 2 for 'simpson' (rows of even / odd shifted index i' = 2r + q, each a
    // half-size Toeplitz problem in r with two passes over the even / odd columns); NR = rows per sub-problem
    int n_sub;
    int64_t N, NR, n_alpha, n_fields, n_tiles, units_per_prob, total_units;
    const double* tabs;   // Toeplitz tables [n_alpha][n_sub * n_pass][ldc]
    int64_t ldc;
    const double *w0, *mult;
    int64_t ldw;
    const double* cm1;
    const double* gp;  // padded fields [n_fields][n_pass][ldgp]
    int64_t ldgp;
    const double* g;   // original fields (the kernels after the prepare kernel do not read them: they may be host memory)
    int64_t ldg;
    const double* g0;  // node 0 (node N of a mirrored field) of every field, copied by the prepare kernel
    double* y;
    int64_t ldy;
    int side;          // FRX_SIDE_LEFT; FRX_SIDE_RIGHT: field read reversed, y[N-i] = -value; FRX_SIDE_RIESZ_CAPUTO: second
                       // pass over a y that holds the left result: y[N-i] = krc * (y[N-i] + sgn * (-value))
    const frx_alpha_scalars* sc;
    // fused gather (left side only): every result row is stored into n_peers buffers -- this GPU's and its NVLink
    // peers' copies of the gathered array -- at row y_row0 + prob instead of into y
    int n_peers;
    int64_t y_row0;
    double* peer_y[FRX_MAX_PEERS];
    double* partial;   // [2 * grid][TI]
    int* tickets;      // [n_alpha * n_fields * n_tiles], zero on entry, zero again on exit
    // plain Toeplitz product with caller-supplied sequences (frx_toeplitz_raw): y_i = raw_scale * S_i, row 0 untouched,
    // no first-column / multiplier terms.  paired: table a goes with fields [a*n_fields, (a+1)*n_fields) only
    int raw, paired;
    double raw_scale;
    // rectangular restriction (raw products whose field is zero beyond column TD*rect_chunks): only row tiles
    // I >= rect_I0 are computed, each over the first rect_chunks column chunks; requires (rect_I0+1)*TI/TD >= rect_chunks
    // (no chunk beyond the tile's diagonal).  rect_chunks = 0: off
    int64_t rect_I0, rect_chunks;
};

// Rows and columns are tiled in SHIFTED coordinates i' = i - 1, j' = j - 1 (row 0 is identically 0 and column 0 is
// the first column, handled in the epilogue), so the operator is exactly N x N lower triangular incl. the diagonal:
// S_{i'} = sum_{j' <= i'} c[i' - j'] g_{j'+1}.  Tile I holds rows i' in [I*TI, (I+1)*TI) and needs the chunks
// (of columns j' for the tensor variant, of distances for the DFMA variant) up to i'_last.
template <class G>
__host__ __device__ inline int64_t tile_chunks(int64_t I, int64_t n_tiles, int64_t N) {
    return I + 1 < n_tiles ? (I + 1) * (G::TI / G::TD) : (N + G::TD - 1) / G::TD;
}
template <class G>
__host__ __device__ inline int64_t tile_prefix(int64_t I, int n_pass) {  // units before tile I (tiles < I are full)
    return (int64_t)n_pass * (G::TI / G::TD) * (I * (I + 1) / 2);
}

struct Cursor {
    int64_t prob, I, v, chunks, units;  // v in [0, units), units = n_pass * chunks
    // decoded once (cursor_at) and then kept up to date incrementally, so the per-unit code has no 64-bit divisions:
    int64_t a, f, chunk;                // prob = ((a * n_fields + f) * n_sub + q),  v = pass * chunks + chunk
    int q, pass;
};

template <class G>
__device__ inline Cursor cursor_at(const HistParams& p, int64_t u) {
    Cursor c;
    c.prob = u / p.units_per_prob;
    int64_t w = u - c.prob * p.units_per_prob;
    if (p.rect_chunks) {
        const int64_t per_tile = p.rect_chunks * p.n_pass, t = w / per_tile;
        c.I = p.rect_I0 + t;
        c.v = w - t * per_tile;
        c.chunks = p.rect_chunks;
    } else {
        const double K = (double)p.n_pass * (G::TI / G::TD);
        int64_t I = (int64_t)((sqrt(1.0 + 8.0 * (double)w / K) - 1.0) * 0.5);
        if (I > p.n_tiles - 1) I = p.n_tiles - 1;
        while (I > 0 && tile_prefix<G>(I, p.n_pass) > w) --I;
        while (I + 1 < p.n_tiles && tile_prefix<G>(I + 1, p.n_pass) <= w) ++I;
        c.I = I;
        c.v = w - tile_prefix<G>(I, p.n_pass);
        c.chunks = tile_chunks<G>(I, p.n_tiles, p.NR);
    }
    c.units = c.chunks * p.n_pass;
    const int64_t pf = c.prob / p.n_sub;
    c.q = (int)(c.prob - pf * p.n_sub);
    c.a = pf / p.n_fields;
    c.f = pf - c.a * p.n_fields;
    c.pass = (int)(c.v / c.chunks);
    c.chunk = c.v - c.pass * c.chunks;
    return c;
}

template <class G>
__device__ inline void cursor_advance(const HistParams& p, Cursor& c) {
    ++c.v;
    if (++c.chunk == c.chunks) {
        c.chunk = 0;
        ++c.pass;
    }
    if (c.v == c.units) {
        c.v = 0;
        c.pass = 0;
        if (++c.I == p.n_tiles) {
            c.I = p.rect_I0;
            ++c.prob;
            if (++c.q == p.n_sub) {
                c.q = 0;
                if (++c.f == p.n_fields) {
                    c.f = 0;
                    ++c.a;
                }
            }
        }
        c.chunks = p.rect_chunks ? p.rect_chunks : tile_chunks<G>(c.I, p.n_tiles, p.NR);
        c.units = c.chunks * p.n_pass;
    }
}

__device__ inline void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ inline void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
// w0[i], mult[i] of the rows i = I*TI + local + 1 of tile c.I -> wbuf[local], wbuf[TI + local]
template <class G>
__device__ inline void stage_epilogue_operands(const HistParams& p, const Cursor& c, double* wbuf) {
    const double* w0 = p.w0 + c.a * p.ldw + c.I * G::TI + 1;
    const double* mu = p.mult + c.a * p.ldw + c.I * G::TI + 1;
    const int64_t left = p.N - c.I * G::TI;   // rows of this tile that exist
    for (int local = threadIdx.x; local < G::TI && local < left; local += kNT) {
        cp_async8(wbuf + local, w0 + local);
        cp_async8(wbuf + G::TI + local, mu + local);
    }
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ inline void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ inline unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ inline void bulk_copy(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// BULK staging of one unit: the same bytes to the same places as stage_unit_mma
template <class G>
__device__ inline void stage_unit_bulk(const HistParams& p, const Cursor& c, char* stage, unsigned long long* bar) {
    const int64_t pass = c.pass, j0 = c.chunk * G::TD, q = c.q, a = c.a, f = c.f;
    const int64_t d_base = c.I * G::TI - j0 - G::TD;
    const double* csrc = p.tabs + ((a * p.n_sub + q) * p.n_pass + pass) * p.ldc + G::C_PAD + d_base;
    const int64_t fg = p.paired ? a * p.n_fields + f : f;
    const double* gsrc = p.gp + (fg * p.n_pass + pass) * p.ldgp + G::PADF + j0 - 8;
    const int t = threadIdx.x;
    // the buffer was last touched through the generic proxy (operand loads, epilogue scratch)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (t == 0) {   // two copies per unit, issued by one thread
        mbar_expect_tx(bar, (unsigned)(G::CWIN_BYTES + G::GCH_ELEMS * 8));
        bulk_copy(stage, csrc, (unsigned)G::CWIN_BYTES, bar);
        bulk_copy(stage + G::CWIN_BYTES, gsrc, (unsigned)(G::GCH_ELEMS * 8), bar);
    }
}

// tensor variant: unit = TI rows x TD columns [j0, j0+TD).  Stage the c window (all distances i - j of the unit,
// as NU rows of 8, in the table's permuted order) and the column chunk of the padded field (+8 halo).
template <class G>
__device__ inline void stage_unit_mma(const HistParams& p, const Cursor& c, char* stage) {
    const int64_t pass = c.pass, j0 = c.chunk * G::TD, q = c.q, a = c.a, f = c.f;
    // row u of the window holds distances 8e..8e+7, e = e_base + u, e_base = p0 - m0 - MC
    const int64_t d_base = c.I * G::TI - j0 - G::TD;
    const double* csrc = p.tabs + ((a * p.n_sub + q) * p.n_pass + pass) * p.ldc + G::C_PAD + d_base;
    const int64_t fg = p.paired ? a * p.n_fields + f : f;
    const double* gsrc = p.gp + (fg * p.n_pass + pass) * p.ldgp + G::PADF + j0 - 8;
    const int t = threadIdx.x;
#pragma unroll
    for (int q = t; q < G::NU * 4; q += kNT) cp_async16(stage + (q >> 2) * G::ROW_STRIDE + (q & 3) * 16, csrc + 2 * q);
    char* gch = stage + G::CWIN_BYTES;
    for (int q = t; q < G::GCH_ELEMS / 2; q += kNT) cp_async16(gch + 16 * q, gsrc + 2 * q);
}

__device__ inline void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// tensor variant: acc[2t], acc[2t+1] = D fragment of tile t (rows 8(P0w + t + T k) + 2 s' + {0,1}).
// The operands of column block mi+1 are loaded before the DMMAs of block mi are issued (register double buffer), so no
// load sits between a warp and its next DMMA -- matters when few warps share the tensor pipe (tails of the launch).
template <class G>
__device__ inline void compute_unit_mma(double (&acc)[G::R], const char* stage, int s0, int s1) {
    constexpr int T = G::T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = lane >> 2, sp = lane & 3;
    // B fragments: g[8m + k - s' - 4kh] = gch[8 + 8mi + k - s' - 4kh]
    const double* gch = reinterpret_cast<const double*>(stage + G::CWIN_BYTES) + (4 + k - sp);
    double2 af[T];
    if constexpr (!G::BULK) {
        // A fragments: row (8T*warp + t + MC - mi + T*k) of the c window, 16 bytes at 16*s' = {c[..+s'], c[..+s'+4]}
        const char* arow = stage + (8 * T * warp + G::MC + T * k) * G::ROW_STRIDE + 16 * sp;
#pragma unroll
        for (int t = 1; t < T; ++t) af[t] = *reinterpret_cast<const double2*>(arow + (t - s0 * T) * G::ROW_STRIDE);
        // operands of the first block
        double2 a_nxt = *reinterpret_cast<const double2*>(arow - (s0 * T) * G::ROW_STRIDE);
        double b1_nxt = gch[8 * s0 * T], b0_nxt = gch[8 * s0 * T + 4];
#pragma unroll 1
        for (int mg = s0 * T; mg < s1 * T; mg += T) {
#pragma unroll
            for (int j = 0; j < T; ++j) {
                const int mi = mg + j;
                // tile t uses slot (t - mi) mod T; the new fragment (tile 0) replaces the one tile T-1 used last time
                af[(T - j) % T] = a_nxt;
                const double b1 = b1_nxt, b0 = b0_nxt;
                // next block (one past the unit's last block reads valid but unused shared memory: the window row
                // below and the 64 bytes of padding behind the field chunk)
                a_nxt = *reinterpret_cast<const double2*>(arow - (mi + 1) * G::ROW_STRIDE);
                b1_nxt = gch[8 * (mi + 1)];
                b0_nxt = gch[8 * (mi + 1) + 4];
#pragma unroll
                for (int t = 0; t < T; ++t) dmma884(acc[2 * t], acc[2 * t + 1], af[(t - j + T) % T].x, b0);
#pragma unroll
                for (int t = 0; t < T; ++t) dmma884(acc[2 * t], acc[2 * t + 1], af[(t - j + T) % T].y, b1);
            }
        }
    } else {
        // dense window (64-byte rows): row u sits at u ^ ((u >> 2) & 1) (frx_table_slot).  The four rows U-4 .. U-1 that
        // one loop iteration loads (U = ubase - mg, a multiple of 4) share that bit, which alternates from one
        // iteration to the next: byte offsets {0,64,128,192} from row U-4, XORed with x = 64 * bit, i.e. constant
        // offsets from the two pointers base + x and base - x.
        static_assert(T == 4, "row swizzle of the dense window is laid out for T = 4");
        const int ubase = 8 * T * warp + G::MC + T * k;
        const char* abase = stage + 16 * sp;
        const int U0 = ubase - s0 * T;                         // multiple of 4
        {
            const int x0 = ((U0 >> 2) & 1) * 64;               // rows U0 .. U0+3
            const char* q = abase + U0 * 64;
#pragma unroll
            for (int t = 1; t < T; ++t) af[t] = *reinterpret_cast<const double2*>(q + ((t * 64) ^ x0));
        }
        double2 a_nxt = *reinterpret_cast<const double2*>(abase + (U0 ^ ((U0 >> 2) & 1)) * 64);
        double b1_nxt = gch[8 * s0 * T], b0_nxt = gch[8 * s0 * T + 4];
        const char* base = abase + (U0 - 4) * 64;              // row U-4 of the current iteration
        int x = (((U0 - 4) >> 2) & 1) * 64;
#pragma unroll 1
        for (int mg = s0 * T; mg < s1 * T; mg += T) {
            const char* pp = base + x;                         // offsets with bit 6 clear: (c ^ x) = c + x
            const char* pm = base - x;                         // offsets with bit 6 set:   (c ^ x) = c - x
#pragma unroll
            for (int j = 0; j < T; ++j) {
                const int mi = mg + j;
                af[(T - j) % T] = a_nxt;
                const double b1 = b1_nxt, b0 = b0_nxt;
                constexpr int kRowOff[4] = {192, 128, 64, 0};  // row U-1-j relative to row U-4
                a_nxt = *reinterpret_cast<const double2*>(((kRowOff[j] & 64) ? pm : pp) + kRowOff[j]);
                b1_nxt = gch[8 * (mi + 1)];
                b0_nxt = gch[8 * (mi + 1) + 4];
#pragma unroll
                for (int t = 0; t < T; ++t) dmma884(acc[2 * t], acc[2 * t + 1], af[(t - j + T) % T].x, b0);
#pragma unroll
                for (int t = 0; t < T; ++t) dmma884(acc[2 * t], acc[2 * t + 1], af[(t - j + T) % T].y, b1);
            }
            base -= 4 * 64;
            x ^= 64;
        }
    }
}

// row of accumulator pair `t` of this thread, relative to the tile's first row (tensor variant)
template <class G>
__device__ inline int mma_row(int t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return 8 * (8 * G::T * warp + t + G::T * (lane >> 2)) + 2 * (lane & 3);
}

// stage the weights chunk and the field window of one unit into shared memory
template <class G>
__device__ inline void stage_unit(const HistParams& p, const Cursor& c, char* stage) {
    if constexpr (G::TENSOR) {
        stage_unit_mma<G>(p, c, stage);
        return;
    } else {
    const int64_t pass = c.pass, d0 = c.chunk * G::TD, q = c.q, a = c.a, f = c.f;
    const double* ctab = p.tabs + ((a * p.n_sub + q) * p.n_pass + pass) * p.ldc + d0;
    const int64_t fg = p.paired ? a * p.n_fields + f : f;
    const double* gsrc = p.gp + (fg * p.n_pass + pass) * p.ldgp + G::PADF + c.I * G::TI - d0 - G::TD;
    const int t = threadIdx.x;
    for (int q = t; q < G::TD / 2; q += kNT) cp_async16(stage + 16 * q, ctab + 2 * q);
    char* win = stage + G::TD * 8;
    constexpr int QPC = G::R / 2;  // 16-byte quads per chunk
#pragma unroll
    for (int q = t; q < G::WIN_ELEMS / 2; q += kNT)
        cp_async16(win + (q / QPC) * G::CHUNK_STRIDE + (q % QPC) * 16, gsrc + 2 * q);
    }
}

template <class G>
__device__ inline void load_chunk(double (&dst)[G::R], const char* win, int chunk) {
    const double2* src = reinterpret_cast<const double2*>(win + chunk * G::CHUNK_STRIDE);
#pragma unroll
    for (int q = 0; q < G::R / 2; ++q) {
        double2 v = src[q];
        dst[2 * q] = v.x;
        dst[2 * q + 1] = v.y;
    }
}

// acc[r] += sum_s c[s] * window(r - s), window(k) = hi[k] for k >= 0, lo[k + R] for k < 0
template <class G>
__device__ inline void fma_block(double (&acc)[G::R], const double* cs, const double (&hi)[G::R],
                                 const double (&lo)[G::R]) {
    constexpr int R = G::R;
#pragma unroll
    for (int s2 = 0; s2 < R; s2 += 2) {
        const double2 c2 = *reinterpret_cast<const double2*>(cs + s2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int s = s2 + h;
            const double c = h ? c2.y : c2.x;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k = r - s;
                acc[r] = fma(c, k >= 0 ? hi[k] : lo[k + R], acc[r]);
            }
        }
    }
}

template <class G>
__device__ inline void compute_unit(double (&acc)[G::R], const char* stage, int s0, int s1) {
    if constexpr (G::TENSOR) {
        compute_unit_mma<G>(acc, stage, s0, s1);
        return;
    } else {
    constexpr int R = G::R;
    const double* cs = reinterpret_cast<const double*>(stage);
    const char* win = stage + G::TD * 8;
    const int c0 = threadIdx.x + G::TD / R;
    double hi[R], lo[R];
    load_chunk<G>(hi, win, c0);
#pragma unroll 1
    for (int b = 0; b < G::TD / R; b += 2) {
        load_chunk<G>(lo, win, c0 - b - 1);
        fma_block<G>(acc, cs + b * R, hi, lo);
        load_chunk<G>(hi, win, c0 - b - 2);  // (last iteration: chunk `threadIdx.x`, loaded but unused)
        fma_block<G>(acc, cs + (b + 1) * R, lo, hi);
    }
    }
}

// position (relative to the tile's first row) of accumulator pair q = 0..R/2-1 of this thread
template <class G>
__device__ inline int acc_pair_row(int q) {
    if constexpr (G::TENSOR) return mma_row<G>(q);
    else return threadIdx.x * G::R + 2 * q;
}

// scale, add the first column / right-of-point terms and store the R rows of this thread
template <class G>
__device__ inline void finish_rows(const HistParams& p, int64_t prob, int64_t I, const double (&acc)[G::R], double* srow,
                                   const double* wsm) {
    // prob counts sub-problems: ((alpha * n_fields + field) * n_sub + q)
    const int64_t pf = prob / p.n_sub, q = prob - pf * p.n_sub, a = pf / p.n_fields, f = pf - a * p.n_fields;
    prob = pf;   // from here on: the (alpha, field) pair that owns the output row
    const int64_t fg = p.paired ? a * p.n_fields + f : f;
    const bool rev = p.side != FRX_SIDE_LEFT;
    const double g0 = p.raw ? 0.0 : p.g0[fg];
    // 'simpson', odd row i: node i+1 = shifted column j' = i (odd: pass 1, slot (i-1)/2 of the padded copy)
    const double* g_odd = p.gp + (fg * p.n_pass + 1) * p.ldgp + G::PADF;
    const double cm1 = p.rule == FRX_RULE_SIMPSON ? p.cm1[a] : 0.0;
    auto store_row = [&](int64_t i, double v) {   // v: value of row i of the (possibly mirrored) left operator
        if (!rev) {
            if (p.n_peers == 0) {
                p.y[prob * p.ldy + i] = v;
            } else {
                const int64_t at = (p.y_row0 + prob) * p.ldy + i;
                for (int q = 0; q < p.n_peers; ++q) p.peer_y[q][at] = v;   // P2P stores over NVLink for q != self
            }
        } else {
            // row i of the mirrored problem is node N-i; create_right_stencil negates the weights (equation.py:38-39)
            double* dst = p.y + prob * p.ldy + (p.N - i);
            if (p.side == FRX_SIDE_RIGHT) {
                *dst = -v;
            } else {
                const bool boundary = (i == 0 || i == p.N);   // one of the two one-sided stencils does not exist there
                *dst = boundary ? nan("") : p.sc[a].krc * (*dst + p.sc[a].sgn * (-v));
            }
        }
    };
    // the accumulators go through shared memory (the stage buffer that was just consumed) so that consecutive threads
    // finish consecutive rows: coalesced loads of w0 / mult and full-line stores of y, which matters when y is a
    // peer's or the host's memory
#pragma unroll
    for (int r = 0; r < G::R; ++r) srow[acc_pair_row<G>(r >> 1) + (r & 1)] = acc[r];
    __syncthreads();
    for (int local = threadIdx.x; local < G::TI; local += kNT) {
        const int64_t ip = (I * G::TI + local) * p.n_sub + q;   // shifted row i' = i - 1
        const int64_t i = ip + 1;
        if (i > p.N) break;
        const double sum = srow[local];
        if (p.raw) {
            store_row(i, p.raw_scale * sum);
            continue;
        }
        if (ip == 0) store_row(0, 0.0);               // row 0 has no stencil: y_0 = 0 by convention
        double v;
        if ((i == 1 && (p.rule == FRX_RULE_RECTANGLE || p.rule == FRX_RULE_SIMPSON)) ||
            (rev && p.rule == FRX_RULE_SIMPSON && i == p.N && (i & 1))) {
            v = nan("");  // the reference cannot build this stencil (resolution 1; mirrored 'simpson' node left of node 0)
        } else {
            // wsm: this tile's w0 / mult rows, prefetched into shared memory with the unit (n_sub == 1 only)
            double s = fma(wsm ? wsm[local] : p.w0[a * p.ldw + i], g0, sum);
            if (p.rule == FRX_RULE_SIMPSON && (i & 1)) s = fma(cm1, g_odd[(i - 1) >> 1], s);
            v = (wsm ? wsm[G::TI + local] : p.mult[a * p.ldw + i]) * s;
        }
        store_row(i, v);
    }
    __syncthreads();   // the buffer is the target of the next prefetch
}

// CTA that owns unit u when `total` units are dealt to `G` CTAs as [k*total/G, (k+1)*total/G)
__device__ inline int64_t span_owner(int64_t u, int64_t total, int64_t G) { return ((u + 1) * G - 1) / total; }

#ifdef FRX_K2_TRACE
// tuning aid (not in production builds): per-CTA timestamps of the phases of the history kernel
__device__ unsigned long long frx_k2_trace[1024 * 16];
__device__ inline unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FRX_TRACE(slot) do { if (threadIdx.x == 0 && blockIdx.x < 1024) frx_k2_trace[blockIdx.x * 16 + (slot)] = trace_now(); } while (0)
// upper-bound experiments (WRONG results, timing only): bit 0 = stage only the first unit of a span, bit 1 = drop the
// barrier after compute
__device__ int frx_k2_debug_mode;
#else
#define FRX_TRACE(slot) do { } while (0)
#endif

template <class G>
__global__ void __launch_bounds__(kNT, G::CTAS_PER_SM) frx_k2_history_main(const HistParams p) {
    extern __shared__ __align__(128) char smem[];
    __shared__ int s_last;
    __shared__ __align__(8) unsigned long long full_bar[2];   // BULK: "unit landed" barriers of the two stage buffers
    constexpr int R = G::R;
    // spans are cut at sub-step granularity (groups of T column blocks for the tensor variant): every CTA gets
    // total_groups / grid groups to within one, so a span may start and end in the middle of a unit
    const int64_t NG = gridDim.x, k = blockIdx.x;
    const int64_t total_groups = p.total_units * G::SUBS;
    const int64_t gbeg = k * total_groups / NG, gend = (k + 1) * total_groups / NG;
    if (gbeg >= gend) return;
    FRX_TRACE(0);
#ifdef FRX_K2_TRACE
    long long trace_compute = 0, trace_groups = 0;
    const long long trace_c0 = clock64();
#endif
    const int64_t ubeg = gbeg / G::SUBS, uend = (gend + G::SUBS - 1) / G::SUBS;
    Cursor cur = cursor_at<G>(p, ubeg), nxt = cur;
    frx_grid_dependency_wait();   // tables / padded fields / tickets come from the prepare kernel
    FRX_TRACE(1);
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    bool from_start = (cur.v == 0 && gbeg == ubeg * G::SUBS);
    const bool wpref = G::WBUF_BYTES > 0 && !p.raw && p.n_sub == 1;
    double* const wbufs = reinterpret_cast<double*>(smem + 2 * G::STAGE_BYTES);
    if constexpr (G::BULK) {
        if (threadIdx.x == 0) {
            mbar_init(&full_bar[0], 1);
            mbar_init(&full_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        stage_unit_bulk<G>(p, cur, smem, &full_bar[0]);
    } else {
        stage_unit<G>(p, cur, smem);
    }
    if (wpref && (cur.v + 1 == cur.units || ubeg + 1 == uend)) stage_epilogue_operands<G>(p, cur, wbufs);
    cp_async_commit();
    for (int64_t u = ubeg; u < uend; ++u) {
        char* stage = smem + ((u - ubeg) & 1) * G::STAGE_BYTES;
        if (u + 1 < uend) {
            cursor_advance<G>(p, nxt);
            if constexpr (G::BULK) {
                stage_unit_bulk<G>(p, nxt, smem + ((u + 1 - ubeg) & 1) * G::STAGE_BYTES, &full_bar[(u + 1 - ubeg) & 1]);
            } else {
#ifdef FRX_K2_TRACE
                if (!(frx_k2_debug_mode & 1))
#endif
                stage_unit<G>(p, nxt, smem + ((u + 1 - ubeg) & 1) * G::STAGE_BYTES);
            }
            if (wpref && (nxt.v + 1 == nxt.units || u + 2 == uend))
                stage_epilogue_operands<G>(p, nxt, wbufs + ((u + 1 - ubeg) & 1) * (G::WBUF_BYTES / 8));
        }
        cp_async_commit();
        cp_async_wait<1>();
        if constexpr (G::BULK) {
            // every thread waits for the unit itself (the async writes are visible to a thread that saw the phase
            // complete); buffer (u - ubeg) & 1 is used for the ((u - ubeg) >> 1)-th time
            mbar_wait(&full_bar[(u - ubeg) & 1], (unsigned)(((u - ubeg) >> 1) & 1));
        } else {
            __syncthreads();
        }
#ifdef FRX_K2_TRACE
        if (u == ubeg) FRX_TRACE(2);
        if (u + 1 == uend) FRX_TRACE(3);
#endif
        const int s0 = u == ubeg ? (int)(gbeg - ubeg * G::SUBS) : 0;
        const int s1 = u + 1 == uend ? (int)(gend - u * G::SUBS) : G::SUBS;
#ifdef FRX_K2_TRACE
        const long long tc0 = clock64();
#endif
        compute_unit<G>(acc, stage, s0, s1);
#ifdef FRX_K2_TRACE
        trace_compute += clock64() - tc0;
        trace_groups += s1 - s0;
#endif
#ifdef FRX_K2_TRACE
        if (!(frx_k2_debug_mode & 2))
#endif
        __syncthreads();
#ifdef FRX_K2_TRACE
        if (u + 1 == uend) FRX_TRACE(4);
#endif
        const bool tile_done = (cur.v + 1 == cur.units && s1 == G::SUBS), span_done = (u + 1 == uend);
        if (tile_done || span_done) {
            if (tile_done && from_start) {
                finish_rows<G>(p, cur.prob, cur.I, acc, reinterpret_cast<double*>(stage),
                            wpref ? wbufs + ((u - ubeg) & 1) * (G::WBUF_BYTES / 8) : nullptr);
            } else {
                // tile shared with other CTAs.  Slot 2k: continuation of a tile begun by an earlier CTA;
                // slot 2k+1: tile begun here but not finished here.
                double* dst = p.partial + (2 * k + (from_start ? 1 : 0)) * G::TI;
#pragma unroll
                for (int r = 0; r < R; r += 2)
                    *reinterpret_cast<double2*>(dst + acc_pair_row<G>(r >> 1)) = make_double2(acc[r], acc[r + 1]);
                const int64_t tile_begin = (u - cur.v) * G::SUBS, tile_end = tile_begin + cur.units * G::SUBS;   // in groups
                const int64_t k0 = span_owner(tile_begin, total_groups, NG), k1 = span_owner(tile_end - 1, total_groups, NG);
                __threadfence();
                __syncthreads();
#ifdef FRX_K2_TRACE
                if (u + 1 == uend) FRX_TRACE(8);    // partial slot published
#endif
                if (threadIdx.x == 0) {
                    int* ticket = p.tickets + (cur.prob * p.n_tiles + cur.I);
                    const int old = atomicAdd(ticket, 1);
                    s_last = (old == (int)(k1 - k0));
                    if (s_last) *ticket = 0;  // leave the tickets clean for the next call
                }
                __syncthreads();
#ifdef FRX_K2_TRACE
                if (u + 1 == uend) FRX_TRACE(9);    // ticket taken
                if (u + 1 == uend && threadIdx.x == 0 && blockIdx.x < 1024) frx_k2_trace[blockIdx.x * 16 + 12] = s_last;
#endif
                if (s_last) {  // every other participant has published its slot: add them in CTA order
                    __threadfence();
                    const double* src = p.partial + (2 * k0 + 1) * G::TI;
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = __ldcg(src + acc_pair_row<G>(r >> 1) + (r & 1));
                    // slots are added in CTA order, the loads of kFixupBatch slots in flight together (a small problem can
                    // split one tile over a hundred CTAs: one L2 round trip per slot would dominate it)
                    constexpr int kFixupBatch = FixupBatch<G>::value;
                    for (int64_t kk = k0 + 1; kk <= k1; kk += kFixupBatch) {
                        double part[kFixupBatch][R];
#pragma unroll
                        for (int j = 0; j < kFixupBatch; ++j) {
                            src = p.partial + (2 * (kk + j <= k1 ? kk + j : k1)) * G::TI;
#pragma unroll
                            for (int r = 0; r < R; ++r) part[j][r] = __ldcg(src + acc_pair_row<G>(r >> 1) + (r & 1));
                        }
#pragma unroll
                        for (int j = 0; j < kFixupBatch; ++j)
                            if (kk + j <= k1) {
#pragma unroll
                                for (int r = 0; r < R; ++r) acc[r] += part[j][r];
                            }
                    }
#ifdef FRX_K2_TRACE
                    if (u + 1 == uend) FRX_TRACE(10);   // slots summed
#endif
                    finish_rows<G>(p, cur.prob, cur.I, acc, reinterpret_cast<double*>(stage),
                            wpref ? wbufs + ((u - ubeg) & 1) * (G::WBUF_BYTES / 8) : nullptr);
#ifdef FRX_K2_TRACE
                    if (u + 1 == uend) FRX_TRACE(11);   // rows stored
#endif
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0.0;
            from_start = true;
        }
        cur = nxt;
    }
#ifdef FRX_K2_TRACE
    FRX_TRACE(5);
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        frx_k2_trace[blockIdx.x * 16 + 6] = smid | ((unsigned long long)trace_groups << 32);
        frx_k2_trace[blockIdx.x * 16 + 7] = (unsigned long long)trace_compute | ((unsigned long long)(clock64() - trace_c0) << 32);
    }
#endif
}

// prepare kernel: blocks [0, n_tab) generate the weight tables (K1); the remaining blocks build the padded
// copies of the fields, gp[(f*n_pass + pass)*ldgp + padf + j'] = g_{j'+1} for 0 <= j' < N (for 'simpson' only the
// nodes of the pass's parity: pass 0 even, pass 1 odd) and 0 elsewhere, and clear the tickets.
__global__ void __launch_bounds__(FRX_TAB_BLOCK, 8) frx_k12_prepare(int rule, const double* __restrict__ alpha, double h, int64_t N,
                                                          const frx_alpha_scalars* __restrict__ sc,
                                                       double* __restrict__ tabs, int64_t ldc, int n_tab_per_alpha, int dec,
                                                       double* __restrict__ w0, double* __restrict__ mult, int64_t ldw,
                                                       double* __restrict__ cm1, int c_pad, int c_perm, int64_t n_tab,
                                                       int64_t tab_per_alpha,
                                                       const double* __restrict__ g, int64_t ldg, int n_pass, int64_t n_fp,
                                                       int64_t ldgp, int padf, double* __restrict__ gp, int* __restrict__ tickets,
                                                       int64_t n_tickets, int reverse, const double* __restrict__ raw_c,
                                                       int64_t raw_ldc, double* __restrict__ g0) {
    __shared__ frx_table_smem S;
    const int64_t bx = blockIdx.x;
    if (raw_c) frx_grid_dependency_wait();   // caller-supplied sequences may come from the previous kernel in the stream
    frx_grid_launch_dependents();   // the history kernel may be scheduled as this grid drains; it waits before reading
    if (bx < n_tab) {
        const int64_t a = bx / tab_per_alpha;
        if (raw_c) {   // copy sequence a into the padded (and permuted) table layout
            const int t = threadIdx.x;
            if (t < FRX_TAB_HALO || t >= FRX_TAB_THREADS - FRX_TAB_HALO) return;
            const int64_t X = (bx - a * tab_per_alpha) * FRX_TAB_ENTRIES + (t - FRX_TAB_HALO);
            if (X >= ldc) return;
            const int64_t x = X - c_pad;
            double* cA = tabs + a * ldc;
            if (x < 0) cA[X] = 0.0;
            else
                cA[frx_table_slot(c_pad, c_perm, x)] = x < N ? raw_c[a * raw_ldc + x] : 0.0;
            return;
        }
        // tables of alpha a: n_tab_per_alpha consecutive tables of ldc entries (1; 'simpson': the 4 decimated ones)
        frx_table_block(S, sc + a, rule, alpha[a], h, N, (bx - a * tab_per_alpha) * FRX_TAB_ENTRIES, c_pad, c_perm, dec,
                        tabs + a * n_tab_per_alpha * ldc, nullptr, ldc, w0 + a * ldw, mult + a * ldw, cm1 + a);
        return;
    }
    const int64_t n_pad_blocks = gridDim.x - n_tab, pb = bx - n_tab;
    const int64_t total = n_fp * ldgp, stride = n_pad_blocks * blockDim.x;
    // batches of 8 elements per thread: all loads of a batch are issued before the first store, so a field that lives in
    // pinned host memory (frx_history_host) costs one PCIe round trip per batch, not per element
    for (int64_t e0 = pb * blockDim.x + threadIdx.x; e0 < total; e0 += 8 * stride) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t e = e0 + q * stride;
            v[q] = 0.0;
            if (e < total) {
                const int64_t fp = e / ldgp, x = e - fp * ldgp;
                const int64_t f = fp / n_pass, pass = fp - f * n_pass, j = x - padf;
                // slot padf + j' holds column j = j' + 1 (shifted coordinates, see tile_chunks); 'simpson' (n_pass = 2):
                // pass e holds every second column, slot padf + k <-> j' = 2k + e ('simpson' also keeps node N+1, column
                // j' = N: no row reaches it, the epilogue of the last odd row reads it).  Mirrored field: node N+1 would be
                // node -1 of the caller's field -- it stays 0 and the one row that needs it, the right-sided 'simpson' row
                // of node 0 with odd N, is NaN in the epilogue.
                const int64_t jj = (n_pass == 1 ? j : 2 * j + pass) + 1;
                if (j >= 0 && jj <= N + (n_pass == 2 && !reverse ? 1 : 0)) v[q] = g[f * ldg + (reverse ? N - jj : jj)];
                if (j == 0 && pass == 0 && !raw_c) g0[f] = g[f * ldg + (reverse ? N : 0)];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t e = e0 + q * stride;
            if (e < total) gp[e] = v[q];
        }
    }
    for (int64_t e = pb * blockDim.x + threadIdx.x; e < n_tickets; e += stride) tickets[e] = 0;
}

struct PeerTargets {
    int n_peers;
    int64_t row0;
    double* y[FRX_MAX_PEERS];
};
struct RawSpec {
    const double* c;   // sequences [n_alpha][ldc], natural order
    int64_t ldc;
    double scale;
    int paired;
    int64_t n_cols;    // > 0: the fields are zero from column n_cols on and only rows > n_cols are wanted
};

template <class G>
int history_impl(frx_ctx* ctx, int rule, int side, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                 int64_t n_fields, const double* d_g, int64_t ldg, double* d_y, int64_t ldy, const PeerTargets* peers,
                 const RawSpec* raw = nullptr) {
    HistParams p;
    p.raw = raw != nullptr;
    p.paired = raw ? raw->paired : 0;
    p.raw_scale = raw ? raw->scale : 1.0;
    p.rule = rule;
    p.side = side;
    p.n_peers = peers ? peers->n_peers : 0;
    p.y_row0 = peers ? peers->row0 : 0;
    for (int q = 0; q < FRX_MAX_PEERS; ++q) p.peer_y[q] = (peers && q < peers->n_peers) ? peers->y[q] : nullptr;
    const bool simpson = rule == FRX_RULE_SIMPSON;
    p.n_pass = simpson ? 2 : 1;
    p.n_sub = simpson ? 2 : 1;
    p.N = N;
    p.NR = simpson ? (N + 1) / 2 : N;       // rows i' = 2r + q, q = 0, 1: at most ceil(N/2) per sub-problem
    p.n_alpha = n_alpha;
    p.n_fields = n_fields;
    p.n_tiles = (p.NR + G::TI - 1) / G::TI;
    p.units_per_prob = tile_prefix<G>(p.n_tiles - 1, p.n_pass) + p.n_pass * tile_chunks<G>(p.n_tiles - 1, p.n_tiles, p.NR);
    p.rect_I0 = 0;
    p.rect_chunks = 0;
    if (raw && raw->n_cols > 0 && raw->n_cols % G::TI == 0 && raw->n_cols < N) {
        p.rect_I0 = raw->n_cols / G::TI;
        // + 1: column block m holds the columns j' = 8m + r - s, so the last columns below n_cols also belong to the
        // first block of the next chunk
        p.rect_chunks = raw->n_cols / G::TD + 1;
        p.units_per_prob = (p.n_tiles - p.rect_I0) * p.rect_chunks * p.n_pass;
    }
    p.total_units = p.units_per_prob * n_alpha * n_fields * p.n_sub;
    // tables are zero beyond the last distance up to the end of the last tile: whole-tile reads stay in bounds and finite
    p.ldc = G::C_PAD + (frx_table_ldc(p.NR) > p.n_tiles * G::TI ? frx_table_ldc(p.NR) : p.n_tiles * G::TI);
    const int n_tabs = p.n_sub * p.n_pass;  // tables per alpha
    p.ldw = (N + 1 + 1) / 2 * 2;
    p.ldgp = G::PADF + p.n_tiles * G::TI;
    int64_t NG = (int64_t)ctx->sm_count * G::CTAS_PER_SM;
    // small problems: at least 4 sub-steps per CTA (fewer participants in the split-tile fix-up)
    if (NG > (p.total_units * G::SUBS + 3) / 4) NG = (p.total_units * G::SUBS + 3) / 4;
    const int64_t n_tickets = n_alpha * n_fields * p.n_sub * p.n_tiles;

    // workspace carve-up (all offsets multiples of 256 bytes)
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off += frx_align_up(bytes, 256);
        return o;
    };
    const size_t o_cA = carve(sizeof(double) * (size_t)n_alpha * n_tabs * p.ldc);
    const size_t o_w0 = carve(sizeof(double) * (size_t)n_alpha * p.ldw);
    const size_t o_mult = carve(sizeof(double) * (size_t)n_alpha * p.ldw);
    const size_t o_cm1 = carve(sizeof(double) * (size_t)n_alpha);
    const int64_t n_fields_all = p.paired ? n_alpha * n_fields : n_fields;
    const size_t o_gp = carve(sizeof(double) * (size_t)n_fields_all * p.n_pass * p.ldgp);
    const size_t o_part = carve(sizeof(double) * (size_t)2 * NG * G::TI);
    const size_t o_tick = carve(sizeof(int) * (size_t)n_tickets);
    const size_t o_scal = carve(frx_alpha_scalars_bytes() * (size_t)n_alpha);
    const size_t o_g0 = carve(sizeof(double) * (size_t)n_fields_all);
    int rc = frx_ws_reserve(ctx, off);
    if (rc) return rc;
    char* ws = (char*)ctx->ws;
    double* cA = (double*)(ws + o_cA);
    double* w0 = (double*)(ws + o_w0);
    double* mult = (double*)(ws + o_mult);
    double* cm1 = (double*)(ws + o_cm1);
    double* gp = (double*)(ws + o_gp);
    p.tabs = cA;
    p.w0 = w0;
    p.mult = mult;
    p.cm1 = cm1;
    p.gp = gp;
    p.g = d_g;
    p.ldg = ldg;
    p.g0 = (const double*)(ws + o_g0);
    p.y = d_y;
    p.ldy = ldy;
    p.sc = (const frx_alpha_scalars*)(ws + o_scal);
    p.partial = (double*)(ws + o_part);
    p.tickets = (int*)(ws + o_tick);

    {
        const bool gl = rule == FRX_RULE_GRUNWALD_LETNIKOV;   // extension: its tables come from their own scan kernel
        // table CTAs cover the storage slots of one table (the base sequences d = 0..N for 'simpson')
        const int64_t cover = simpson ? G::C_PAD + N + 1 : p.ldc;
        const int64_t tab_per_alpha = (cover + FRX_TAB_ENTRIES - 1) / FRX_TAB_ENTRIES, n_tab = gl ? 0 : tab_per_alpha * n_alpha;
        if (simpson)   // the four decimated tables are scattered into: padding, tails and T(0,1)[0] must be zero
            FRX_CUDA(cudaMemsetAsync(cA, 0, sizeof(double) * (size_t)n_alpha * n_tabs * p.ldc, ctx->stream));
        const int64_t n_fp = n_fields_all * p.n_pass;
        int64_t n_pad = (n_fp * p.ldgp + 8 * kNT - 1) / (8 * kNT);
        if (n_pad > 4096) n_pad = 4096;
        FRX_REQUIRE(n_tab + n_pad < ((int64_t)1 << 31), FRX_ERR_INVALID_ARG, "problem too large for one launch");
        if (!raw) {
            rc = frx_launch_alpha_scalars(ctx, rule, n_alpha, d_alpha, h, ws + o_scal);
            if (rc) return rc;
        }
        if (gl) {
            rc = frx_launch_gl_table(ctx, n_alpha, d_alpha, h, N, ws + o_scal, G::C_PAD, G::C_PERM, cA, p.ldc, w0, mult, p.ldw, cm1);
            if (rc) return rc;
        }
        FRX_LAUNCH(ctx, FRX_K_WEIGHTS_TABLE,
                   FRX_CUDA(frx_launch_dependent(frx_k12_prepare, dim3((unsigned)(n_tab + n_pad)), dim3(FRX_TAB_BLOCK), 0,
                                                 ctx->stream, rule, d_alpha, h, N, (const frx_alpha_scalars*)(ws + o_scal), cA, p.ldc, n_tabs, (int)simpson, w0, mult, p.ldw, cm1, G::C_PAD, G::C_PERM, n_tab, tab_per_alpha, d_g, ldg, p.n_pass,
                                                 n_fp, p.ldgp, (int)G::PADF, gp, p.tickets, n_tickets,
                                                 (int)(side != FRX_SIDE_LEFT), raw ? raw->c : (const double*)nullptr,
                                                 raw ? raw->ldc : (int64_t)0, (double*)(ws + o_g0))));
    }
    // dynamic shared memory: at least the double buffer, padded so that exactly CTAS_PER_SM CTAs fit per SM
    int smem_sm = 0;
    FRX_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    // exactly CTAS_PER_SM CTAs per SM: the request is padded only if one more CTA would fit otherwise.  What is left of
    // the SM's shared memory lets CTAs of the (small) prepare kernel stay resident next to the early-launched CTAs of
    // this one (programmatic dependent launch), so the CTA launch ramp hides behind the prepare kernel.
    const int need = 2 * G::STAGE_BYTES + 2 * G::WBUF_BYTES, per_cta_overhead = 1024 + 128;
    int dyn = need;
    if (smem_sm / (need + per_cta_overhead) > G::CTAS_PER_SM)
        dyn = (smem_sm / (G::CTAS_PER_SM + 1) - per_cta_overhead + 128) / 128 * 128;   // too big for CTAS_PER_SM + 1
    if (dyn < need) dyn = need;
    FRX_CUDA(cudaFuncSetAttribute(frx_k2_history_main<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    FRX_LAUNCH(ctx, FRX_K_HISTORY_MAIN,
               FRX_CUDA(frx_launch_dependent(frx_k2_history_main<G>, dim3((unsigned)NG), dim3(kNT), (size_t)dyn, ctx->stream, p)));
    return FRX_OK;
}

}  // namespace

extern "C" int64_t frx_table_ldc(int64_t N) { return (N + 1 + kLdcQuantum - 1) / kLdcQuantum * kLdcQuantum; }

static int history_dispatch(frx_ctx* ctx, int rule, int side, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                            int64_t n_fields, const double* d_g, int64_t ldg, double* d_y, int64_t ldy,
                            const PeerTargets* peers = nullptr, const void* raw_spec = nullptr) {
    if (raw_spec)
        return history_impl<MGeo<4, 1024, 2, true>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, nullptr,
                                                    (const RawSpec*)raw_spec);
    // tile geometry: FRX_K2_VARIANT is a tuning/experiment switch, the default is the measured best
    static const int variant = getenv("FRX_K2_VARIANT") ? atoi(getenv("FRX_K2_VARIANT")) : 11;
    switch (variant) {
        case 0: return history_impl<Geo<8, 256, 4>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 1: return history_impl<Geo<16, 128, 3>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 4: return history_impl<MGeo<4, 256, 4>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 5: return history_impl<MGeo<4, 1024, 3>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 6: return history_impl<MGeo<4, 512, 3>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 7: return history_impl<MGeo<4, 512, 2>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 8: return history_impl<MGeo<4, 512, 1>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 10: return history_impl<MGeo<4, 1024, 2>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        case 9: return history_impl<MGeo<4, 512, 4>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
        default: return history_impl<MGeo<4, 1024, 2, true>>(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy, peers);
    }
}

extern "C" int frx_history_sided(frx_ctx* ctx, int rule, int side, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                                 int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, double* d_y, int64_t ldy) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_LAST, FRX_ERR_UNKNOWN_RULE, "unknown approximation key");
    FRX_REQUIRE(side >= FRX_SIDE_LEFT && side <= FRX_SIDE_RIESZ_CAPUTO, FRX_ERR_INVALID_ARG, "bad side");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30), FRX_ERR_INVALID_ARG, "N out of range");
    FRX_REQUIRE(n_alpha >= 0 && n_alpha <= 65535 && n_fields >= 0 && n_fields <= 32767, FRX_ERR_INVALID_ARG,
                "n_alpha (<= 65535) / n_fields (<= 32767) out of range");
    FRX_REQUIRE(side == FRX_SIDE_LEFT || rule != FRX_RULE_GRUNWALD_LETNIKOV, FRX_ERR_DOMAIN,
                "the Gruenwald-Letnikov extension offers the left operator only");
    const int64_t need_g = N + 1 + (rule == FRX_RULE_SIMPSON ? 1 : 0);
    FRX_REQUIRE(n_g >= need_g && ldg >= n_g && ldy >= N + 1, FRX_ERR_INVALID_ARG,
                "field too short (needs N+1 nodes, N+2 for 'simpson') or leading dimension too small");
    if (n_alpha == 0 || n_fields == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_g && d_y, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    if (side == FRX_SIDE_RIESZ_CAPUTO) {   // K * (left + (-1.)**n * right): the left pass leaves its result in y
        int rc = history_dispatch(ctx, rule, FRX_SIDE_LEFT, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy);
        if (rc) return rc;
    }
    return history_dispatch(ctx, rule, side, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, d_y, ldy);
}

// Fused compute + gather for sweeps sharded over the GPUs of one box: the history kernel's epilogue stores every
// result row straight into each peer's copy of the gathered array (peer pointers obtained by the caller through CUDA
// IPC / symmetric memory), so the transfer rides on NVLink while the remaining tiles are still being computed and no
// collective follows -- only a barrier before the gathered array is read.
extern "C" int frx_history_gather(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                                  int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, int32_t n_peers,
                                  const uint64_t* h_peer_ptrs, int64_t row0, int64_t ldy) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(rule >= FRX_RULE_CAPUTO && rule <= FRX_RULE_LAST, FRX_ERR_UNKNOWN_RULE, "unknown approximation key");
    FRX_REQUIRE(n_peers >= 1 && n_peers <= FRX_MAX_PEERS && h_peer_ptrs != nullptr, FRX_ERR_INVALID_ARG, "bad peer list");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30) && row0 >= 0, FRX_ERR_INVALID_ARG, "N / row offset out of range");
    FRX_REQUIRE(n_alpha >= 0 && n_alpha <= 65535 && n_fields >= 0 && n_fields <= 32767, FRX_ERR_INVALID_ARG,
                "n_alpha (<= 65535) / n_fields (<= 32767) out of range");
    const int64_t need_g = N + 1 + (rule == FRX_RULE_SIMPSON ? 1 : 0);
    FRX_REQUIRE(n_g >= need_g && ldg >= n_g && ldy >= N + 1, FRX_ERR_INVALID_ARG,
                "field too short (needs N+1 nodes, N+2 for 'simpson') or leading dimension too small");
    if (n_alpha == 0 || n_fields == 0) return FRX_OK;
    FRX_REQUIRE(d_alpha && d_g, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    PeerTargets pt;
    pt.n_peers = n_peers;
    pt.row0 = row0;
    for (int q = 0; q < FRX_MAX_PEERS; ++q) pt.y[q] = q < n_peers ? (double*)(uintptr_t)h_peer_ptrs[q] : nullptr;
    for (int q = 0; q < n_peers; ++q) FRX_REQUIRE(pt.y[q] != nullptr, FRX_ERR_INVALID_ARG, "NULL peer pointer");
    return history_dispatch(ctx, rule, FRX_SIDE_LEFT, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, pt.y[0], ldy, &pt);
}

extern "C" int frx_history(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                           int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, double* d_y, int64_t ldy) {
    return frx_history_sided(ctx, rule, FRX_SIDE_LEFT, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, n_g, d_y, ldy);
}

// Plain lower-triangular Toeplitz products with caller-supplied sequences on the same kernel (used by frx_history_solve):
//     y[a][f][i] = scale * sum_{j=1..i} c[a][i-j] * g[.][j],   i = 1..N        (y[..][0] is not written)
// paired != 0: sequence a is applied to fields a*n_fields .. a*n_fields+n_fields-1 only (y has n_seq*n_fields rows);
// otherwise every sequence is applied to every field.  n_cols > 0 promises g[.][j] = 0 for j > n_cols and asks for rows
// i > n_cols only (the rows up to n_cols are left untouched when n_cols is a multiple of the row tile, else computed).
int frx_toeplitz_raw(frx_ctx* ctx, int64_t n_seq, const double* d_c, int64_t ldc, int64_t N, int64_t n_fields, int paired,
                     const double* d_g, int64_t ldg, double scale, double* d_y, int64_t ldy, int64_t n_cols) {
    RawSpec rs{d_c, ldc, scale, paired, n_cols};
    return history_dispatch(ctx, FRX_RULE_CAPUTO, FRX_SIDE_LEFT, n_seq, nullptr, 0.0, N, n_fields, d_g, ldg, d_y, ldy, nullptr, &rs);
}

extern "C" int frx_toeplitz_product(frx_ctx* ctx, int64_t n_seq, const double* d_c, int64_t ldc, int64_t N, int64_t n_fields,
                                    int32_t paired, const double* d_g, int64_t ldg, double scale, double* d_y, int64_t ldy,
                                    int64_t n_cols) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(N >= 1 && N < ((int64_t)1 << 30) && n_cols >= 0, FRX_ERR_INVALID_ARG, "N / n_cols out of range");
    FRX_REQUIRE(n_seq >= 0 && n_seq <= 65535 && n_fields >= 0 && n_fields <= 32767, FRX_ERR_INVALID_ARG,
                "n_seq (<= 65535) / n_fields (<= 32767) out of range");
    FRX_REQUIRE(ldc >= N && ldg >= N && ldy >= N, FRX_ERR_INVALID_ARG, "leading dimension too small");
    if (n_seq == 0 || n_fields == 0) return FRX_OK;
    FRX_REQUIRE(d_c && d_g && d_y, FRX_ERR_INVALID_ARG, "NULL buffer");
    FRX_ON_DEVICE(ctx);
    // the kernel's rows and columns are numbered from 1 (node 0 is the first-column term, absent here)
    return frx_toeplitz_raw(ctx, n_seq, d_c, ldc, N, n_fields, paired, d_g - 1, ldg, scale, d_y - 1, ldy, n_cols);
}

// K3: one alpha, many fields -- the same tensor-core kernel with the fields as independent problems (the c window
// of a unit is shared by all fields through L2)
extern "C" int frx_toeplitz_apply(frx_ctx* ctx, int rule, double alpha, double h, int64_t N, int64_t n_fields,
                                  const double* d_G, int64_t ldg, int64_t n_g, double* d_Y, int64_t ldy) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    int rc = frx_check_rule_alpha(rule, alpha);
    if (rc) return rc;
    FRX_ON_DEVICE(ctx);
    FRX_CUDA(cudaMemcpyAsync(ctx->scalars, &alpha, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return frx_history(ctx, rule, 1, (const double*)ctx->scalars, h, N, n_fields, d_G, ldg, n_g, d_Y, ldy);
}

// Host-buffer entry point.  Pinned (page-locked, device-mapped) host buffers are used in place: the prepare kernel
// reads the field straight from host memory while its other CTAs build the weight tables, and the history kernel's
// epilogue stores every finished row straight into the host result as the tiles complete, so neither transfer is a
// separate step on the critical path.  Pageable buffers go through copies to a context-owned device buffer.
static const void* mapped_device_pointer(const void* h) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
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
    // FRX_HOST_PROFILE=1: accumulate where the host-side time of this call goes (printed every 64 calls)
    static const bool prof = getenv("FRX_HOST_PROFILE") != nullptr;
    static double prof_acc[3];
    static int prof_calls;
    auto now = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
    };
    const double pt0 = prof ? now() : 0.0;
    FRX_ON_DEVICE(ctx);
    // FRX_HOST_ZERO_COPY: bit 0 = read the field in place, bit 1 = store the result in place
    static const int zero_copy = getenv("FRX_HOST_ZERO_COPY") ? atoi(getenv("FRX_HOST_ZERO_COPY")) : 3;
    const double* map_g = (zero_copy & 1) ? (const double*)mapped_device_pointer(h_g) : nullptr;
    double* map_y = (zero_copy & 2) ? (double*)mapped_device_pointer(h_y) : nullptr;
    // caller-facing device buffers: a grow-only allocation owned by the context (separate from the
    // workspace, which frx_history re-carves)
    const size_t b_alpha = frx_align_up(sizeof(double) * n_alpha, 256);
    const size_t b_g = map_g ? 0 : frx_align_up(sizeof(double) * (size_t)n_fields * ldg, 256);
    const size_t b_y = sizeof(double) * (size_t)n_alpha * n_fields * ldy;
    int rc = frx_io_reserve(ctx, b_alpha + b_g + (map_y ? 0 : b_y));
    if (rc) return rc;
    char* io = (char*)ctx->io;
    double* d_alpha = (double*)io;
    const double* d_g = map_g ? map_g : (const double*)(io + b_alpha);
    double* d_y = map_y ? map_y : (double*)(io + b_alpha + b_g);
    FRX_CUDA(cudaMemcpyAsync(d_alpha, h_alpha, sizeof(double) * n_alpha, cudaMemcpyHostToDevice, ctx->stream));
    if (!map_g)
        FRX_CUDA(cudaMemcpyAsync(io + b_alpha, h_g, sizeof(double) * (size_t)n_fields * ldg, cudaMemcpyHostToDevice, ctx->stream));
    const double pt1 = prof ? now() : 0.0;
    rc = frx_history(ctx, rule, n_alpha, d_alpha, h, N, n_fields, d_g, ldg, n_g, d_y, ldy);
    if (rc) return rc;
    if (!map_y) FRX_CUDA(cudaMemcpyAsync(h_y, d_y, b_y, cudaMemcpyDeviceToHost, ctx->stream));
    const double pt2 = prof ? now() : 0.0;
    FRX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (prof) {
        const double pt3 = now();
        prof_acc[0] += pt1 - pt0;
        prof_acc[1] += pt2 - pt1;
        prof_acc[2] += pt3 - pt2;
        if (++prof_calls % 64 == 0) {
            fprintf(stderr, "frx_history_host: setup + alpha copy %.1f us, launches %.1f us, wait %.1f us (mean of 64 calls)\n",
                    prof_acc[0] / 64, prof_acc[1] / 64, prof_acc[2] / 64);
            prof_acc[0] = prof_acc[1] = prof_acc[2] = 0.0;
        }
    }
    return FRX_OK;
}

#ifdef FRX_K2_TRACE
extern "C" int frx_debug_k2_mode(int mode) {
    return cudaMemcpyToSymbol(frx_k2_debug_mode, &mode, sizeof(int)) == cudaSuccess ? 0 : -1;
}
extern "C" int frx_debug_k2_trace(unsigned long long* h_out /*[1024*16]*/) {
    return cudaMemcpyFromSymbol(h_out, frx_k2_trace, sizeof(unsigned long long) * 1024 * 16) == cudaSuccess ? 0 : -1;
}
#endif
