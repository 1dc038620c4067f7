// K4, narrowest band-width bin (nb = res + 1 <= 7): ONE THREAD PER SYSTEM.
//
// The block LDL^T of frx_solve_ldl.cuh spends a warp on a system whatever its band width; at nb <= 7 its 8 x 8 tile is
// the whole band and the warp's ~450 instructions per 8 columns are bookkeeping around 8 * 28 useful multiply-adds.
// Here a lane owns a system: the (nb + 1) x (nb + 1) window of the symmetric band (lower triangle, 36 doubles), the
// window of the right-hand side and the coefficients live in registers, every index is a compile-time constant (the
// window is a circular buffer over the residues mod 8 in both directions, so the rank-1 update of column k IS the shift to
// column k + 1: no moves), and a column costs 47 FP64 instructions per WARP, i.e. per 32 systems.
//
//   forward   for k = 0 .. n-1:  d = W[k][k] (sign test);  l_i = W[k+i][k] / d;  W[k+i][k+j] -= l_i W[k+j][k];
//             z_k = b_k;  b_{k+i} -= l_i z_k;  row k + 8 of the Toeplitz matrix enters the window;
//             l_1..l_7 and y_k = z_k / d go to the scratch, [column][8][lane]: coalesced
//   backward  x_k = y_k - sum_i l_i(k) x_{k+i}   (the term in x_{k+1} last: one dependent DFMA per column)
//
// Right-hand sides and solutions are row-major per system: they pass through a 32 x 32 shared-memory tile (padded), so
// that global memory sees 256-byte rows (the buffers may be page-locked host memory, frx_assemble_solve_host).
// A warp keeps ONE scratch slab (n x 8 x 32 doubles) for all the groups of 32 systems it works through: it is written
// and read back (in reverse) while it is still in L2, and discarded at the end.
//
// Only DEFINITE systems end here (every pivot of the sign of a_0: elimination without interchanges is backward stable).
// A system with a pivot of the other sign is appended to a retry list on the device; frx_k4_band_bldl<1> is launched
// behind this kernel with that list as its work (refinement + residual test, then the pivoted LU for what fails).
#pragma once
#include "frx_solve_ldl.cuh"

namespace frx_k4 {

struct ScalarGeo {
    static constexpr int NB = 7, W = NB + 1;
    static constexpr int TILE = 32 * 33;                                   // doubles of the transposition tile
    __host__ __device__ static inline int npad(int n) { return (n + 7) & ~7; }
    __host__ __device__ static inline int64_t ustride(int n) { return (int64_t)npad(n) * W * 32; }
    __host__ __device__ static constexpr int tri(int p, int q) { return p >= q ? p * (p + 1) / 2 + q : q * (q + 1) / 2 + p; }
};

__global__ void __launch_bounds__(32) frx_k4_band_scalar(const BandParams p) {
    using G = ScalarGeo;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = G::W;
    __shared__ double tiles[2 * G::TILE];            // right-hand-side entries 32t .. 32t + 31 of the 32 systems in tiles[(t & 1) * TILE]
    double* const tile = tiles;                        // (the backward pass stages the solutions in the first one)
    const int lane = threadIdx.x, n = p.n, npad = G::npad(n);
    double* U = p.uscratch + (int64_t)blockIdx.x * p.ustride_cta + lane;
#pragma unroll 1
    for (int64_t base = (int64_t)blockIdx.x * 32; base < p.n_sys; base += (int64_t)gridDim.x * 32) {
        const int n_live = p.n_sys - base < 32 ? (int)(p.n_sys - base) : 32;
        const bool live = lane < n_live;
        const int64_t s = p.order[base + (live ? lane : 0)];              // idle lanes shadow the group's first system
        const int res = (int)((p.wptr[s + 1] - p.wptr[s] - 1) / 2);
        const double ih = 1.0 / p.h[s];
        const double* rc = p.rcw + p.wptr[s] + res;
        const int64_t row_b = s * p.ldb, row_x = s * p.ldx;
        double a[W];                                                      // a_m = A[i + m][i], zero beyond the band
#pragma unroll
        for (int m = 0; m < W; ++m) a[m] = m <= res + 1 ? band_coefficient(-m, res, ih, rc) : 0.0;
        const double sgn = a[0] < 0.0 ? -1.0 : 1.0;
        const int sgn_bit = a[0] < 0.0 ? (int)0x80000000 : 0;
        // window: entry (k + i, k + j) of the Schur complement at the residues ((k + i) % 8, (k + j) % 8); rows beyond n are
        // identity rows of the sign of a_0 (their pivots pass the sign test, their multipliers are zero)
        double Wd[W * (W + 1) / 2], bw[W];
#pragma unroll
        for (int i = 0; i < W; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Wd[G::tri(i, j)] = i < n ? a[i - j] : (i == j ? sgn : 0.0);
        int piv_min = 0x7ff00000;
        double r_last = 0.0;
        // 32 entries of the 32 right-hand sides -> one tile: 32 asynchronous 8-byte copies per lane, all in flight at once
        // (row r = system r, 256 contiguous bytes; zero beyond n and for idle lanes)
        auto issue_tile = [&](const int t) {
            double* dst = tiles + (t & 1) * G::TILE + lane * 33;
            const int c = 32 * t + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int64_t rb = __shfl_sync(FULL, row_b, r);
                if (r < n_live && c < n) cp_async8(dst + r, p.rhs + rb + c);
                else dst[r] = 0.0;
            }
        };
        __syncwarp();
        issue_tile(0);
        cp_async_wait_all();
        __syncwarp();
        issue_tile(1);                                                    // travels while the first 24 columns are eliminated
#pragma unroll
        for (int i = 0; i < W; ++i) bw[i] = tiles[i * 33 + lane];
        // ---- forward ------------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int k0 = 0; k0 < npad; k0 += 8) {
            if ((k0 & 31) == 24) {          // columns from here on let in entries of the next tile; the one before it is dead
                cp_async_wait_all();
                __syncwarp();
                issue_tile((k0 >> 5) + 2);
            }
            const double* tin = tiles + (((k0 + 8) >> 5) & 1) * G::TILE + lane;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int k = k0 + kk;
                // residues: column k is kk, row k + i is (kk + i) % 8
                const double d = Wd[G::tri(kk, kk)];
                piv_min = min(piv_min, __double2hiint(d) ^ sgn_bit);
                const double r = pivot_reciprocal(d);
                if (kk == 7) r_last = r;
                double l[W];
#pragma unroll
                for (int i = 1; i < W; ++i) l[i] = Wd[G::tri((kk + i) & 7, kk)] * r;
                const double z = bw[kk];
                double* u = U + (int64_t)k * (W * 32);
#pragma unroll
                for (int i = 1; i < W; ++i) {
                    bw[(kk + i) & 7] = fma(-l[i], z, bw[(kk + i) & 7]);
                    __stcg(u + (i - 1) * 32, l[i]);
                }
                __stcg(u + (W - 1) * 32, z * r);
#pragma unroll
                for (int i = 1; i < W; ++i)
#pragma unroll
                    for (int j = 1; j <= i; ++j)
                        Wd[G::tri((kk + i) & 7, (kk + j) & 7)] =
                            fma(-l[i], Wd[G::tri((kk + j) & 7, kk)], Wd[G::tri((kk + i) & 7, (kk + j) & 7)]);
                // row k + 8 enters at the residue of column k: its entries against the columns k + 1 .. k + 8
                const bool in = k + 8 < n;
#pragma unroll
                for (int j = 1; j < W; ++j) Wd[G::tri(kk, (kk + j) & 7)] = in ? a[W - j] : 0.0;
                Wd[G::tri(kk, kk)] = in ? a[0] : sgn;
                bw[kk] = tin[((k + 8) & 31) * 33];                       // right-hand side entry k + 8 (zero beyond n)
            }
        }
        // a pivot p with p * sgn(a_0) <= 1e-300 (0x01a56e1f = high word of 1e-300) marks the system as not definite; a NaN
        // (or infinite) pivot passes that integer test but poisons every later column: the last reciprocal shows it
        const bool ok = piv_min > 0x01a56e1f && r_last == r_last;
        cp_async_wait_all();                                              // (a tile beyond the last column may still be in flight)
        __syncwarp();
        // ---- backward ------------------------------------------------------------------------------------------------
        double xw[W];                                                     // x_{k+1 .. k+7} at the residues mod 8
#pragma unroll
        for (int i = 0; i < W; ++i) xw[i] = 0.0;
        const unsigned okmask = __ballot_sync(FULL, live && ok);
#pragma unroll 1
        for (int k0 = npad - 8; k0 >= 0; k0 -= 8) {
#pragma unroll
            for (int kk = 7; kk >= 0; --kk) {
                const int k = k0 + kk;
                const double* u = U + (int64_t)k * (W * 32);
                double acc = __ldcg(u + (W - 1) * 32);
#pragma unroll
                for (int i = W - 1; i >= 1; --i) acc = fma(-__ldcg(u + (i - 1) * 32), xw[(kk + i) & 7], acc);
                xw[kk] = acc;
                tile[(k & 31) * 33 + lane] = acc;
            }
            if ((k0 & 31) == 0) {                                         // entries k0 .. k0 + 31 of the 32 solutions are complete
                __syncwarp();
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int64_t rx = __shfl_sync(FULL, row_x, r);
                    const int c = k0 + lane;
                    if (((okmask >> r) & 1u) && c < n) p.x[rx + c] = tile[lane * 33 + r];
                }
                __syncwarp();
            }
        }
        if (live) {
            if (ok) {
                if (p.info) p.info[s] = 0;
            } else {
                p.fail_list[atomicAdd(p.fail_count, 1)] = (int32_t)s;
            }
        }
        __syncwarp();
        // the slab is dead: its lines must not be written back when the next group overwrites or evicts them
        discard_l2(p.uscratch + (int64_t)blockIdx.x * p.ustride_cta, p.ustride_cta * 8, lane, 32);
        __syncwarp();
    }
}

}  // namespace frx_k4
