// frx_compose.cu -- stencil o stencil composition on the device: the expansion that the
// reference's call sites perform through `fdm.Operator(stencil_a, fdm.Operator(stencil_b))`
// (reference fractulus/test/integration/test_equation.py:61-66), i.e. for every node of A (in A's own order) the whole
// stencil B shifted to that node, weights multiplied and merged per resulting node:
//     C[o] = sum_{ia = 0, 1, ...} wa[ia] * wb[ib(ia)],        off_a[ia] + off_b[ib(ia)] == o,
// the first term assigned and the later ones added in the order of ia -- exactly the floating-point operations of the
// dict-merging expansion, so the result is bit-identical to it (oracle/fdm_shim.py Operator.expand).  Offsets are
// integers in a unit the caller chooses (the finest grid both stencils live on, e.g. h/2 when a central difference of
// span h is involved).  One thread per output node; B's offsets must ascend (binary search), A's order is the
// summation order.  Ragged batches: item k owns entries [ptr[k], ptr[k+1]) of its arrays.
#include "frx_internal.h"

namespace {

__global__ void frx_k6_stencil_compose(int64_t n_items, int64_t total_c, const int64_t* __restrict__ ptr_a,
                                       const int32_t* __restrict__ off_a, const double* __restrict__ w_a,
                                       const int64_t* __restrict__ ptr_b, const int32_t* __restrict__ off_b,
                                       const double* __restrict__ w_b, const int64_t* __restrict__ ptr_c,
                                       const int32_t* __restrict__ off_c, double* __restrict__ w_c) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total_c; e += (int64_t)gridDim.x * blockDim.x) {
        // item that owns output entry e: last k with ptr_c[k] <= e
        int64_t lo = 0, hi = n_items;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr_c[mid] <= e) lo = mid; else hi = mid;
        }
        const int64_t k = lo, a0 = ptr_a[k], a1 = ptr_a[k + 1], b0 = ptr_b[k], b1 = ptr_b[k + 1];
        const int32_t o = off_c[e];
        double acc = 0.0;
        bool any = false;
        for (int64_t ia = a0; ia < a1; ++ia) {
            const int64_t want = (int64_t)o - off_a[ia];
            int64_t l = b0, r = b1;            // first entry of B with offset >= want
            while (l < r) {
                const int64_t m = (l + r) >> 1;
                if (off_b[m] < want) l = m + 1; else r = m;
            }
            if (l < b1 && off_b[l] == want) {
                const double t = __dmul_rn(w_a[ia], w_b[l]);
                acc = any ? __dadd_rn(acc, t) : t;
                any = true;
            }
        }
        w_c[e] = acc;
    }
}

}  // namespace

extern "C" int frx_stencil_compose(frx_ctx* ctx, int64_t n_items, const int64_t* d_ptr_a, const int32_t* d_off_a,
                                   const double* d_w_a, const int64_t* d_ptr_b, const int32_t* d_off_b, const double* d_w_b,
                                   const int64_t* d_ptr_c, const int32_t* d_off_c, int64_t total_c, double* d_w_c) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_REQUIRE(n_items >= 0 && total_c >= 0, FRX_ERR_INVALID_ARG, "bad sizes");
    if (n_items == 0 || total_c == 0) return FRX_OK;
    FRX_REQUIRE(d_ptr_a && d_off_a && d_w_a && d_ptr_b && d_off_b && d_w_b && d_ptr_c && d_off_c && d_w_c, FRX_ERR_INVALID_ARG,
                "NULL buffer");
    FRX_ON_DEVICE(ctx);
    int64_t blocks = (total_c + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    FRX_LAUNCH(ctx, FRX_K_STENCIL, frx_k6_stencil_compose<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                                       n_items, total_c, d_ptr_a, d_off_a, d_w_a, d_ptr_b, d_off_b, d_w_b, d_ptr_c, d_off_c, d_w_c));
    return FRX_OK;
}
