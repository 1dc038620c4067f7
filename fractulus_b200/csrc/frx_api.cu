// frx_api.cu -- context management, error reporting and the host-side integer/domain logic of the C ABI
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "frx_internal.h"
#include "frx_rules.cuh"

static thread_local char g_err[512] = "";

void frx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int frx_version(void) { return FRX_VERSION; }
extern "C" const char* frx_last_error_string(void) { return g_err; }

extern "C" int frx_device_count(int* n) {
    FRX_REQUIRE(n != nullptr, FRX_ERR_INVALID_ARG, "n_devices is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *n = c;
    return FRX_OK;
}

extern "C" int frx_ctx_create(int device, frx_ctx** out) {
    FRX_REQUIRE(out != nullptr, FRX_ERR_INVALID_ARG, "out is NULL");
    int n = 0;
    frx_device_count(&n);
    if (n <= 0) {
        frx_set_error("no CUDA device visible: libfrx_b200 has no CPU path");
        return FRX_ERR_NO_DEVICE;
    }
    FRX_REQUIRE(device >= 0 && device < n, FRX_ERR_INVALID_ARG, "device index out of range");
    frx_ctx* c = (frx_ctx*)calloc(1, sizeof(frx_ctx));
    FRX_REQUIRE(c != nullptr, FRX_ERR_ALLOC, "host allocation failed");
    c->device = device;
    c->stream = 0;
    c->k4_n_sys = -1;
    frx_device_guard guard(device);
    cudaError_t e = guard.err;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&c->scalars, 512);   // [0, 256): by-value scalars; [256, 512): K4's fail counters
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->switch_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        frx_set_error("context creation on device %d failed: %s", device, cudaGetErrorString(e));
        if (c->scalars) cudaFree(c->scalars);
        if (c->switch_ev) cudaEventDestroy(c->switch_ev);
        free(c);
        return FRX_ERR_CUDA;
    }
    *out = c;
    return FRX_OK;
}

extern "C" int frx_ctx_destroy(frx_ctx* ctx) {
    if (!ctx) return FRX_OK;
    frx_device_guard guard(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->io) cudaFree(ctx->io);
    if (ctx->scalars) cudaFree(ctx->scalars);
    if (ctx->switch_ev) cudaEventDestroy(ctx->switch_ev);
    if (ctx->bin_fork) cudaEventDestroy(ctx->bin_fork);
    for (int k = 0; k < 2; ++k) {
        if (ctx->k4_hst[k]) cudaFreeHost(ctx->k4_hst[k]);
        if (ctx->k4_img[k]) cudaFree(ctx->k4_img[k]);
        if (ctx->k4_copied[k]) cudaEventDestroy(ctx->k4_copied[k]);
        if (ctx->k4_done[k]) cudaEventDestroy(ctx->k4_done[k]);
    }
    if (ctx->k4_copy_stream) cudaStreamDestroy(ctx->k4_copy_stream);
    for (int b = 0; b < 4; ++b) {
        if (ctx->bin_join[b]) cudaEventDestroy(ctx->bin_join[b]);
        if (ctx->bin_stream[b]) cudaStreamDestroy(ctx->bin_stream[b]);
    }
    for (int i = 0; i < FRX_EV_POOL; ++i) {
        if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
    }
    free(ctx);
    return FRX_OK;
}

// The context owns one workspace, one io buffer and one scalars block, so everything it queues must execute in
// order.  Within one stream that is the stream's order; when the caller moves the context to another stream,
// an event recorded on the old stream (covering all work queued so far) is made a dependency of the new one.
extern "C" int frx_ctx_set_stream(frx_ctx* ctx, void* s) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    cudaStream_t ns = (cudaStream_t)s;
    if (ns == ctx->stream) return FRX_OK;
    FRX_ON_DEVICE(ctx);
    FRX_CUDA(cudaEventRecord(ctx->switch_ev, ctx->stream));
    FRX_CUDA(cudaStreamWaitEvent(ns, ctx->switch_ev, 0));
    ctx->stream = ns;
    return FRX_OK;
}

extern "C" int frx_host_register(void* h_ptr, uint64_t bytes) {
    FRX_REQUIRE(h_ptr != nullptr && bytes > 0, FRX_ERR_INVALID_ARG, "NULL pointer or zero size");
    cudaError_t e = cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return FRX_OK;
    }
    FRX_CUDA(e);
    return FRX_OK;
}

extern "C" int frx_host_unregister(void* h_ptr) {
    FRX_REQUIRE(h_ptr != nullptr, FRX_ERR_INVALID_ARG, "NULL pointer");
    cudaError_t e = cudaHostUnregister(h_ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return FRX_OK;
    }
    FRX_CUDA(e);
    return FRX_OK;
}

extern "C" int frx_ctx_synchronize(frx_ctx* ctx) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_ON_DEVICE(ctx);
    FRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

extern "C" int frx_ctx_launch_count(frx_ctx* ctx, int64_t* n) {
    FRX_REQUIRE(ctx != nullptr && n != nullptr, FRX_ERR_INVALID_ARG, "NULL argument");
    *n = ctx->launches;
    return FRX_OK;
}

extern "C" int frx_solve_stats(frx_ctx* ctx, int64_t* n_systems, int64_t* n_pivoted, int64_t* pivoted_per_bin,
                               int64_t* refined_per_bin) {
    FRX_REQUIRE(ctx != nullptr && n_systems && n_pivoted, FRX_ERR_INVALID_ARG, "NULL argument");
    FRX_ON_DEVICE(ctx);
    *n_systems = ctx->k4_n_sys < 0 ? 0 : ctx->k4_n_sys;
    *n_pivoted = *n_systems;
    for (int b = 0; b < 4; ++b) {
        if (pivoted_per_bin) pivoted_per_bin[b] = -1;
        if (refined_per_bin) refined_per_bin[b] = -1;
    }
    if (ctx->k4_n_sys < 0 || !ctx->k4_ldl) return FRX_OK;
    int32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    FRX_CUDA(cudaMemcpyAsync(cnt, (const char*)ctx->scalars + 256, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
    FRX_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_pivoted = (int64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3];
    for (int b = 0; b < 4; ++b) {
        if (pivoted_per_bin) pivoted_per_bin[b] = cnt[b];
        if (refined_per_bin) refined_per_bin[b] = cnt[4 + b];
    }
    return FRX_OK;
}

// ---- K4's parameter image: two pinned host blocks, two device blocks, one copy stream -------------------------------
// Call i fills host block i & 1, uploads it on the copy stream into device block i & 1 and makes the context's stream wait
// for that copy only.  The copy itself waits for the kernels of call i - 2 (the last readers of that device block), not for
// those of call i - 1: it runs under them.  (One block on each side put 4.4 MB of upload, ~0.1 ms, between the kernels of
// consecutive 1e5-system calls.)
int frx_k4_stage_begin(frx_ctx* ctx, size_t bytes, int* slot, void** host_image) {
    const int k = (int)(ctx->k4_calls & 1u);
    *slot = k;
    if (ctx->k4_copied_pending[k]) {             // the copy that last read this host block
        FRX_CUDA(cudaEventSynchronize(ctx->k4_copied[k]));
        ctx->k4_copied_pending[k] = 0;
    }
    if (bytes > ctx->k4_hst_bytes[k]) {
        if (ctx->k4_hst[k]) {
            FRX_CUDA(cudaFreeHost(ctx->k4_hst[k]));
            ctx->k4_hst[k] = nullptr;
            ctx->k4_hst_bytes[k] = 0;
        }
        const size_t want = frx_align_up(bytes + bytes / 8, (size_t)1 << 16);
        cudaError_t e = cudaMallocHost(&ctx->k4_hst[k], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->k4_hst[k] = nullptr;
            frx_set_error("pinned staging allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return FRX_ERR_ALLOC;
        }
        ctx->k4_hst_bytes[k] = want;
    }
    *host_image = ctx->k4_hst[k];
    return FRX_OK;
}

int frx_k4_stage_upload(frx_ctx* ctx, int k, size_t bytes, void** device_image) {
    if (!ctx->k4_copy_stream) {
        FRX_CUDA(cudaStreamCreateWithFlags(&ctx->k4_copy_stream, cudaStreamNonBlocking));
        for (int j = 0; j < 2; ++j) {
            FRX_CUDA(cudaEventCreateWithFlags(&ctx->k4_copied[j], cudaEventDisableTiming));
            FRX_CUDA(cudaEventCreateWithFlags(&ctx->k4_done[j], cudaEventDisableTiming));
        }
    }
    if (bytes > ctx->k4_img_bytes[k]) {
        if (ctx->k4_img[k]) {                    // its last readers must be through before it goes
            if (ctx->k4_done_recorded[k]) FRX_CUDA(cudaEventSynchronize(ctx->k4_done[k]));
            FRX_CUDA(cudaFree(ctx->k4_img[k]));
            ctx->k4_img[k] = nullptr;
            ctx->k4_img_bytes[k] = 0;
        }
        const size_t want = frx_align_up(bytes + bytes / 8, (size_t)1 << 16);
        cudaError_t e = cudaMalloc(&ctx->k4_img[k], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->k4_img[k] = nullptr;
            frx_set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return FRX_ERR_ALLOC;
        }
        ctx->k4_img_bytes[k] = want;
    }
    if (ctx->k4_done_recorded[k]) FRX_CUDA(cudaStreamWaitEvent(ctx->k4_copy_stream, ctx->k4_done[k], 0));
    FRX_CUDA(cudaMemcpyAsync(ctx->k4_img[k], ctx->k4_hst[k], bytes, cudaMemcpyHostToDevice, ctx->k4_copy_stream));
    FRX_CUDA(cudaEventRecord(ctx->k4_copied[k], ctx->k4_copy_stream));
    ctx->k4_copied_pending[k] = 1;
    FRX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->k4_copied[k], 0));
    *device_image = ctx->k4_img[k];
    return FRX_OK;
}

// after the last kernel of the call that reads device block `k` has been queued on the context's stream
int frx_k4_stage_done(frx_ctx* ctx, int k) {
    FRX_CUDA(cudaEventRecord(ctx->k4_done[k], ctx->stream));
    ctx->k4_done_recorded[k] = 1;
    ctx->k4_calls += 1;
    return FRX_OK;
}

int frx_bin_streams(frx_ctx* ctx) {
    if (ctx->bin_fork) return FRX_OK;
    for (int b = 0; b < 4; ++b) {
        FRX_CUDA(cudaStreamCreateWithFlags(&ctx->bin_stream[b], cudaStreamNonBlocking));
        FRX_CUDA(cudaEventCreateWithFlags(&ctx->bin_join[b], cudaEventDisableTiming));
    }
    FRX_CUDA(cudaEventCreateWithFlags(&ctx->bin_fork, cudaEventDisableTiming));
    return FRX_OK;
}

int frx_timing_begin(frx_ctx* ctx, int kid) {
    if (!ctx->timing) return FRX_OK;
    if (ctx->ev_n == FRX_EV_POOL) {
        int rc = frx_timing_flush(ctx);
        if (rc) return rc;
    }
    int i = ctx->ev_n;
    if (!ctx->ev0[i]) {
        FRX_CUDA(cudaEventCreate(&ctx->ev0[i]));
        FRX_CUDA(cudaEventCreate(&ctx->ev1[i]));
    }
    ctx->ev_kid[i] = kid;
    FRX_CUDA(cudaEventRecord(ctx->ev0[i], ctx->stream));
    return FRX_OK;
}

int frx_timing_end(frx_ctx* ctx) {
    if (!ctx->timing) return FRX_OK;
    FRX_CUDA(cudaEventRecord(ctx->ev1[ctx->ev_n], ctx->stream));
    ctx->ev_n += 1;
    return FRX_OK;
}

int frx_timing_flush(frx_ctx* ctx) {
    for (int i = 0; i < ctx->ev_n; ++i) {
        float ms = 0.f;
        FRX_CUDA(cudaEventSynchronize(ctx->ev1[i]));
        FRX_CUDA(cudaEventElapsedTime(&ms, ctx->ev0[i], ctx->ev1[i]));
        ctx->k_ms[ctx->ev_kid[i]] += ms;
        ctx->k_calls[ctx->ev_kid[i]] += 1;
    }
    ctx->ev_n = 0;
    return FRX_OK;
}

extern "C" int frx_ctx_set_timing(frx_ctx* ctx, int on) {
    FRX_REQUIRE(ctx != nullptr, FRX_ERR_INVALID_ARG, "ctx is NULL");
    FRX_ON_DEVICE(ctx);
    int rc = frx_timing_flush(ctx);
    if (rc) return rc;
    ctx->timing = on ? 1 : 0;
    if (on) {
        for (int k = 0; k < FRX_K_COUNT; ++k) {
            ctx->k_ms[k] = 0.0;
            ctx->k_calls[k] = 0;
        }
    }
    return FRX_OK;
}

extern "C" int frx_ctx_kernel_times(frx_ctx* ctx, double* ms, int64_t* calls) {
    FRX_REQUIRE(ctx && ms && calls, FRX_ERR_INVALID_ARG, "NULL argument");
    FRX_ON_DEVICE(ctx);
    int rc = frx_timing_flush(ctx);
    if (rc) return rc;
    for (int k = 0; k < FRX_K_COUNT; ++k) {
        ms[k] = ctx->k_ms[k];
        calls[k] = ctx->k_calls[k];
    }
    return FRX_OK;
}

int frx_ws_reserve(frx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return FRX_OK;
    FRX_ON_DEVICE(ctx);
    if (ctx->ws) {
        FRX_CUDA(cudaStreamSynchronize(ctx->stream));
        FRX_CUDA(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
    }
    size_t want = frx_align_up(bytes + bytes / 8, (size_t)1 << 20);
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        frx_set_error("device workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return FRX_ERR_ALLOC;
    }
    ctx->ws_bytes = want;
    return FRX_OK;
}

int frx_io_reserve(frx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->io_bytes) return FRX_OK;
    FRX_ON_DEVICE(ctx);
    if (ctx->io) {
        FRX_CUDA(cudaStreamSynchronize(ctx->stream));
        FRX_CUDA(cudaFree(ctx->io));
        ctx->io = nullptr;
        ctx->io_bytes = 0;
    }
    size_t want = frx_align_up(bytes + bytes / 8, (size_t)1 << 20);
    cudaError_t e = cudaMalloc(&ctx->io, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        frx_set_error("device buffer allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return FRX_ERR_ALLOC;
    }
    ctx->io_bytes = want;
    return FRX_OK;
}

// ---- domain checks mirroring what the reference does with such settings -------------------------
int frx_check_rule_alpha(int rule, double alpha) {
    if (rule < FRX_RULE_CAPUTO || rule > FRX_RULE_LAST) {
        frx_set_error("unknown approximation key (reference: KeyError at equation.py:29)");
        return FRX_ERR_UNKNOWN_RULE;
    }
    if (!(alpha >= 0.0) || !isfinite(alpha)) {
        frx_set_error("alpha must be finite and >= 0");
        return FRX_ERR_DOMAIN;
    }
    if (rule == FRX_RULE_GRUNWALD_LETNIKOV) return FRX_OK;
    if (rule == FRX_RULE_RECTANGLE && alpha > 1.0) {
        frx_set_error("'rectangle' with alpha > 1: 0.0 ** negative (reference: ZeroDivisionError, equation.py:68)");
        return FRX_ERR_ZERO_DIVISION;
    }
    // 'trapezoidal' and 'simpson' give real weights for every alpha < 2 (all pow bases are >= 0 for resolution >= 2 and
    // (1.-alpha)**2 has an integer exponent); beyond that the reference fails:
    if (rule == FRX_RULE_TRAPEZOIDAL && alpha > 2.0) {
        frx_set_error("'trapezoidal' with alpha > 2: 0.0 ** negative (reference: ZeroDivisionError, equation.py:95)");
        return FRX_ERR_ZERO_DIVISION;
    }
    if (rule == FRX_RULE_SIMPSON && alpha >= 2.0) {
        frx_set_error("'simpson' with alpha >= 2: gamma(2 - alpha) at a pole / 0.0 ** negative / complex weights in the "
                      "reference (equation.py:114-139)");
        return FRX_ERR_DOMAIN;
    }
    return FRX_OK;
}

// create_riesz_caputo_stencil multiplies by gamma(2. - alpha) (equation.py:23): a pole for alpha = 2, 3, ...
// (reference: ValueError "math domain error").  For non-integer alpha > 2 the reference evaluates CPython's
// reflection branch of math.gamma through libm sin/cos; that branch is not restated here -- rejected (documented).
int frx_check_riesz_alpha(double alpha) {
    if (alpha >= 2.0) {
        frx_set_error("Riesz-Caputo with alpha >= 2: gamma(2 - alpha) is not positive-argument (reference: ValueError "
                      "at integer alpha, equation.py:23)");
        return FRX_ERR_DOMAIN;
    }
    return FRX_OK;
}

int frx_check_settings(int rule, double alpha, double resolution) {
    int rc = frx_check_rule_alpha(rule, alpha);
    if (rc) return rc;
    if (resolution == 0.0) {
        frx_set_error("resolution 0 (reference: ZeroDivisionError, equation.py:56,64,99,105)");
        return FRX_ERR_ZERO_DIVISION;
    }
    if (!(resolution > 0.0) || resolution != floor(resolution) || resolution > 2.0e9) {
        frx_set_error("resolution must be a positive integer value");
        return FRX_ERR_DOMAIN;
    }
    if (rule == FRX_RULE_RECTANGLE && resolution < 2.0) {
        frx_set_error("'rectangle' with resolution 1 asks for a uniform stencil with zero intervals");
        return FRX_ERR_ZERO_DIVISION;
    }
    if (rule == FRX_RULE_SIMPSON && resolution < 2.0) {
        frx_set_error("'simpson' with resolution 1: (m-2)**(3-alpha) has a negative base (complex in the reference)");
        return FRX_ERR_DOMAIN;
    }
    return FRX_OK;
}

extern "C" int frx_stencil_layout(int rule, int side, double alpha, double resolution, int32_t* first_offset,
                                  int32_t* count) {
    FRX_REQUIRE(first_offset && count, FRX_ERR_INVALID_ARG, "NULL output pointer");
    FRX_REQUIRE(side >= FRX_SIDE_LEFT && side <= FRX_SIDE_RIESZ_CAPUTO, FRX_ERR_INVALID_ARG, "bad side");
    FRX_REQUIRE(rule != FRX_RULE_GRUNWALD_LETNIKOV || side == FRX_SIDE_LEFT, FRX_ERR_DOMAIN,
                "the Gruenwald-Letnikov extension offers the left stencil only");
    int rc = frx_check_settings(rule, alpha, resolution);
    if (rc) return rc;
    if (side == FRX_SIDE_RIESZ_CAPUTO) {
        rc = frx_check_riesz_alpha(alpha);
        if (rc) return rc;
        FRX_REQUIRE(resolution <= 1.0e9, FRX_ERR_DOMAIN, "Riesz-Caputo stencil of 2*resolution+1 nodes exceeds int32");
    }
    long long f, c;
    frx_left_layout(rule, (long long)resolution, &f, &c);
    long long last = f + c - 1;
    if (side == FRX_SIDE_RIGHT) {              // o -> -o                                     equation.py:38-39
        f = -last;
    } else if (side == FRX_SIDE_RIESZ_CAPUTO) {  // union of left and mirrored left           equation.py:20-25
        long long lo = f < -last ? f : -last;
        f = lo;
        c = -2 * lo + 1;
    }
    *first_offset = (int32_t)f;
    *count = (int32_t)c;
    return FRX_OK;
}

extern "C" int frx_stencil_layout_batch(int rule, int side, int64_t n_settings, const double* h_alpha, const double* h_resolution,
                                        int32_t* h_first_offset, int64_t* h_row_ptr, int64_t* bad_index) {
    FRX_REQUIRE(n_settings >= 0 && h_row_ptr && (n_settings == 0 || (h_alpha && h_resolution)), FRX_ERR_INVALID_ARG,
                "NULL array or negative count");
    h_row_ptr[0] = 0;
    for (int64_t s = 0; s < n_settings; ++s) {
        int32_t first, count;
        const int rc = frx_stencil_layout(rule, side, h_alpha[s], h_resolution[s], &first, &count);
        if (rc) {
            if (bad_index) *bad_index = s;
            return rc;
        }
        if (h_first_offset) h_first_offset[s] = first;
        h_row_ptr[s + 1] = h_row_ptr[s] + count;
    }
    return FRX_OK;
}

extern "C" int64_t frx_table_ldc(int64_t N);  // defined in frx_history.cu (depends on the tile sizes)
