// frx_internal.h -- context object and error plumbing shared by the translation units of libfrx_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/frx.h"

#define FRX_EV_POOL 64
struct frx_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    int64_t launches;      // kernels launched through this context
    // grow-only device workspace (weight tables, padded fields, stream-K partials)
    void* ws;
    size_t ws_bytes;
    // grow-only device buffers behind the *_host entry points (caller-facing inputs / outputs)
    void* io;
    size_t io_bytes;
    void* scalars;         // 256 bytes of device memory for by-value scalar arguments, then K4's four fail counters
    int64_t k4_n_sys;      // systems of the last frx_assemble_solve (-1: none yet), k4_ldl: it went through the pivot-free attempt
    int k4_ldl;
    // cross-stream ordering: everything this context has queued is ordered before work on a newly set stream
    cudaEvent_t switch_ev;
    // K4: the per-call image of the host-side parameters (settings, row pointers, bin order), double-buffered on both
    // sides and uploaded on a copy stream, so that the upload of call i+1 runs under the kernels of call i
    // (frx_k4_stage_begin / _upload / _done)
    void* k4_hst[2];
    size_t k4_hst_bytes[2];
    void* k4_img[2];
    size_t k4_img_bytes[2];
    cudaEvent_t k4_copied[2], k4_done[2];
    int k4_copied_pending[2], k4_done_recorded[2];
    cudaStream_t k4_copy_stream;
    unsigned k4_calls;
    // K4: one stream per band-width bin, forked from / joined to `stream` inside frx_assemble_solve (frx_bin_streams)
    cudaStream_t bin_stream[4];
    cudaEvent_t bin_fork, bin_join[4];
    // optional per-kernel CUDA-event timing (frx_ctx_set_timing): bench.py's roofline numbers
    int timing;
    int ev_n;
    cudaEvent_t ev0[FRX_EV_POOL], ev1[FRX_EV_POOL];
    int ev_kid[FRX_EV_POOL];
    double k_ms[FRX_K_COUNT];
    int64_t k_calls[FRX_K_COUNT];
};

// the double-buffered parameter image of K4 (see frx_ctx)
int frx_k4_stage_begin(frx_ctx* ctx, size_t bytes, int* slot, void** host_image);
int frx_k4_stage_upload(frx_ctx* ctx, int slot, size_t bytes, void** device_image);
int frx_k4_stage_done(frx_ctx* ctx, int slot);
// creates the bin streams and their events on first use
int frx_bin_streams(frx_ctx* ctx);
// a kernel launch bracketed by events when timing is on; counts the launch either way
int frx_timing_begin(frx_ctx* ctx, int kid);
int frx_timing_end(frx_ctx* ctx);
int frx_timing_flush(frx_ctx* ctx);
#define FRX_LAUNCH(ctx, kid, ...)                 \
    do {                                          \
        int rc_ = frx_timing_begin((ctx), (kid)); \
        if (rc_) return rc_;                      \
        __VA_ARGS__;                              \
        (ctx)->launches += 1;                     \
        FRX_CUDA(cudaGetLastError());             \
        rc_ = frx_timing_end((ctx));              \
        if (rc_) return rc_;                      \
    } while (0)

void frx_set_error(const char* fmt, ...);

#define FRX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            frx_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FRX_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define FRX_REQUIRE(cond, code, msg)         \
    do {                                     \
        if (!(cond)) {                       \
            frx_set_error("%s", msg);        \
            return code;                     \
        }                                    \
    } while (0)

// ensure ctx->ws holds at least `bytes`; contents are NOT preserved on growth
int frx_ws_reserve(frx_ctx* ctx, size_t bytes);
int frx_io_reserve(frx_ctx* ctx, size_t bytes);
// pinned staging: reserve waits for the previous upload out of the block (if any) and grows it; release records the
// event after the caller has queued its copy on ctx->stream
// host-side domain check of one Settings triple against the reference's behaviour
int frx_check_settings(int rule, double alpha, double resolution);
int frx_check_rule_alpha(int rule, double alpha);
int frx_check_riesz_alpha(double alpha);

static inline size_t frx_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __cplusplus
// Entry points run on the context's device and leave the caller's current device as they found it.
struct frx_device_guard {
    int prev, target;
    cudaError_t err;
    explicit frx_device_guard(int device) : prev(-1), target(device), err(cudaSuccess) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        if (err != cudaSuccess) cudaGetLastError();
    }
    bool ok() const { return err == cudaSuccess; }
    ~frx_device_guard() {
        if (prev >= 0 && prev != target) cudaSetDevice(prev);
    }
};
#define FRX_ON_DEVICE(ctx)                                                            \
    frx_device_guard frx_guard_((ctx)->device);                                       \
    do {                                                                              \
        if (!frx_guard_.ok()) {                                                       \
            frx_set_error("cudaSetDevice(%d) failed: %s", (ctx)->device, cudaGetErrorString(frx_guard_.err)); \
            return FRX_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)
#endif

#ifdef __CUDACC__
#include <utility>
// launch `kernel` as a programmatic dependent of the previous kernel in the stream: it may start while its
// predecessor is still running and must execute griddepcontrol.wait before touching the predecessor's output
template <typename... KArgs, typename... Args>
static inline cudaError_t frx_launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                               cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif
