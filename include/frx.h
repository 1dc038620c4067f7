/* frx.h -- C ABI of libfrx_b200.so: the B200 (sm_100a) implementation of the fractional-operator hot
 * path of szajek/fractulus.  Plain C, plain pointers and sizes; no torch / C++ types.
 *
 * The reference is pure Python and has no FFI of its own; each entry point below names the reference
 * interface it replaces (file:line under the reference checkout).  The Python host layer
 * (fractulus_b200/equation.py, fractulus_b200/batched.py) binds these with ctypes; INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns FRX_OK (0) or a negative frx_status; frx_last_error_string() gives the
 *     thread-local message of the last failure.
 *   - "d_" pointers are device pointers on the context's device, "h_" pointers are host pointers.
 *   - device-pointer entry points are asynchronous on the context's stream (frx_ctx_set_stream);
 *     "_host" entry points copy in, compute, copy out and synchronise before returning.
 *   - rule  : the reference's approximation key (fractulus/equation.py:179-184).
 *   - side  : which of the reference's three stencil factories (equation.py:13,28,32).
 *   - every weight is produced on the GPU; there is no CPU fallback anywhere in this library.
 */
#ifndef FRX_H
#define FRX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRX_VERSION 100 /* 0.1.0 */

typedef enum {
    FRX_OK = 0,
    FRX_ERR_CUDA = -1,          /* a CUDA runtime call failed (message has the CUDA error)          */
    FRX_ERR_INVALID_ARG = -2,   /* null pointer, negative size, bad leading dimension ...            */
    FRX_ERR_UNKNOWN_RULE = -3,  /* reference: KeyError from _factory[approximation], equation.py:29  */
    FRX_ERR_ZERO_DIVISION = -4, /* reference: ZeroDivisionError (resolution 0, equation.py:56,64,99,105;
                                   'rectangle' with alpha > 1, equation.py:68; 'rectangle' resolution 1) */
    FRX_ERR_DOMAIN = -5,        /* settings for which the reference yields complex / undefined weights */
    FRX_ERR_NO_DEVICE = -6,     /* no CUDA device: this library never computes on the CPU             */
    FRX_ERR_ALLOC = -7          /* device workspace allocation failed                                 */
} frx_status;

typedef enum {                  /* keys of _factory, reference fractulus/equation.py:179-184 */
    FRX_RULE_CAPUTO = 0,        /* _create_left_caputo_stencil,           equation.py:42-59   */
    FRX_RULE_RECTANGLE = 1,     /* _create_left_rectangle_rule_stencil,   equation.py:62-73   */
    FRX_RULE_TRAPEZOIDAL = 2,   /* _create_left_trapezoidal_rule_stencil, equation.py:84-100  */
    FRX_RULE_SIMPSON = 3,       /* _create_left_simpson_rule_stencil,     equation.py:103-176 */
    /* EXTENSION, not a key of the reference: Gruenwald-Letnikov weights of the Riemann-Liouville
     * derivative, w_0 = 1, w_k = w_{k-1} * (1 - (alpha+1)/k), stencil weight h**(-alpha) * w_k at offset -k,
     * k = 0..resolution.  Left side only.  Python key 'grunwald_letnikov'. */
    FRX_RULE_GRUNWALD_LETNIKOV = 4
} frx_rule;
#define FRX_RULE_LAST FRX_RULE_GRUNWALD_LETNIKOV

typedef enum {
    FRX_SIDE_LEFT = 0,          /* create_left_stencil,         equation.py:28-29 */
    FRX_SIDE_RIGHT = 1,         /* create_right_stencil,        equation.py:32-39 */
    FRX_SIDE_RIESZ_CAPUTO = 2   /* create_riesz_caputo_stencil, equation.py:13-25 */
} frx_side;

typedef struct frx_ctx frx_ctx; /* one per (host thread, device); not thread-safe across threads */

/* ---- library / context ------------------------------------------------------------------------ */
int frx_version(void);
const char* frx_last_error_string(void);
int frx_device_count(int* n_devices);
int frx_ctx_create(int device, frx_ctx** out);
int frx_ctx_destroy(frx_ctx* ctx);
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = legacy default.
 * A context owns ONE workspace: everything it queues executes in order.  Moving it to another stream inserts an event
 * dependency (work queued so far on the old stream happens-before work queued afterwards on the new one). */
int frx_ctx_set_stream(frx_ctx* ctx, void* cuda_stream);
int frx_ctx_synchronize(frx_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int frx_ctx_launch_count(frx_ctx* ctx, int64_t* n_launches);

/* Page-lock and map an existing host allocation (cudaHostRegister, portable + mapped) so that the "_host" entry points
 * use it IN PLACE: the prepare kernel reads fields from it and the history kernel's epilogue stores result rows into it
 * over PCIe, no staging copies.  Meant for buffers this library's caller did not allocate with cudaHostAlloc -- in
 * particular a POSIX shared-memory segment that every process of a multi-GPU sweep maps, so that each GPU stores its
 * share of the result straight into the one host array (fractulus_b200/sweep.py).  No counterpart in the reference. */
int frx_host_register(void* h_ptr, uint64_t bytes);
int frx_host_unregister(void* h_ptr);

/* Per-kernel device timing with CUDA events on the context's stream (off by default; adds two event
 * records per launch).  frx_ctx_kernel_times synchronises, then fills ms[k] / calls[k] for the kernel
 * classes below, accumulated since timing was last switched on. */
typedef enum {
    FRX_K_WEIGHTS_TABLE = 0,   /* K1 */
    FRX_K_PAD_FIELDS = 1,      /* K2 prep */
    FRX_K_HISTORY_MAIN = 2,    /* K2 */
    FRX_K_HISTORY_FIXUP = 3,   /* K2 split-tile fix-up */
    FRX_K_STENCIL = 4,         /* K5 */
    FRX_K_TOEPLITZ = 5,        /* K3 */
    FRX_K_SOLVE = 6,           /* K4 */
    FRX_K_OTHER = 7,
    FRX_K_COUNT = 8
} frx_kernel_class;
int frx_ctx_set_timing(frx_ctx* ctx, int on);
int frx_ctx_kernel_times(frx_ctx* ctx, double* ms /*[FRX_K_COUNT]*/, int64_t* calls /*[FRX_K_COUNT]*/);

/* ---- integer layout (host-only integer logic, no GPU work) -------------------------------------
 * The nodes of every reference stencil sit at consecutive integer multiples of lf/resolution:
 * offsets first_offset .. first_offset+count-1 (reference equation.py:59,70-71,96-100,173-176 through
 * _create_parametrized_stencil, equation.py:76-81; mirrored by equation.py:38-39; merged by :20-25). */
int frx_stencil_layout(int rule, int side, double alpha, double resolution, int32_t* first_offset,
                       int32_t* count);

/* The same for n_settings settings at once: row_ptr[0] = 0, row_ptr[s+1] = row_ptr[s] + count_s (host arrays;
 * first_offset may be NULL).  Stops at the first invalid setting: its status is returned and *bad_index set (may be NULL). */
int frx_stencil_layout_batch(int rule, int side, int64_t n_settings, const double* h_alpha, const double* h_resolution,
                             int32_t* h_first_offset, int64_t* h_row_ptr, int64_t* bad_index);

/* ---- K5: stencil weights for a batch of Settings(alpha, lf, resolution) ------------------------
 * Replaces create_left_stencil / create_right_stencil / create_riesz_caputo_stencil
 * (equation.py:13-39) for n_settings settings at once.  Stencil s occupies
 * [row_ptr[s], row_ptr[s+1]) of d_offsets / d_weights, row_ptr[s+1]-row_ptr[s] == its layout count
 * (host array, from frx_stencil_layout).  d_offsets may be NULL. */
int frx_stencil_weights(frx_ctx* ctx, int rule, int side, int64_t n_settings, const double* d_alpha,
                        const double* d_lf, const double* d_resolution, const int64_t* h_row_ptr,
                        int32_t* d_offsets, double* d_weights);
/* one Settings triple, host buffers of `capacity` entries; *count receives the stencil length */
int frx_stencil_weights_host(frx_ctx* ctx, int rule, int side, double alpha, double lf, double resolution,
                             int32_t capacity, int32_t* h_offsets, double* h_weights, int32_t* count);

/* ---- K1: Toeplitz weight tables of the growing-history operator -------------------------------
 * Row i of the history operator is the reference stencil of Settings(alpha, lf = i*h, resolution = i)
 * (call pattern of reference fractulus/test/integration/test_equation.py:51-66).  Its raw weights
 * depend on the distance d = i - j only, except node number 0 (equation.py:49,90,113-117), so for
 * each alpha the operator is:   y_i = mult[i] * ( w0[i]*g_0 + sum_{d=0}^{i-1} c[d][j parity]*g_{i-d}
 *                                                  + [simpson, i odd] cm1*g_{i+1} ).
 *   d_cA[a*ldc + d], d = 0..N-1 : raw Toeplitz weights ('simpson': those multiplying even nodes j)
 *   d_cB[a*ldc + d]             : 'simpson' only (odd nodes j); may be NULL for the other rules
 *   d_w0[a*ldw + i], d_mult[a*ldw + i], i = 0..N : first-column raw weight and row multiplier
 *   d_cm1[a]                    : 'simpson' right-of-point raw weight (equation.py:161-162), else 0
 * ldc >= frx_table_ldc(N) (entries d >= N are written as 0), ldw >= N+1. */
int64_t frx_table_ldc(int64_t N);
int frx_weights_table(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                      double* d_cA, double* d_cB, int64_t ldc, double* d_w0, double* d_mult, int64_t ldw,
                      double* d_cm1);

/* ---- K2: history sum (lower-triangular Toeplitz mat-vec), few fields, many alpha ---------------
 * y[(a*n_fields + f)*ldy + i] = D^alpha_a of field f at node i, i = 0..N, for the growing-history
 * operator above; field f is g[f*ldg + j], j = 0..n_g-1 with n_g >= N+1 (N+2 for 'simpson').
 * y_0 = 0; rows for which the reference cannot build a stencil ('rectangle' i=1, 'simpson' i=1) = NaN.
 * Weight tables are generated on the fly inside the call (K1) into the context workspace. */
int frx_history(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, double* d_y, int64_t ldy);
/* The same for the right-sided operator (row i = create_right_stencil(rule, Settings(alpha, (N-i)*h, N-i)),
 * equation.py:32-39; y_N = 0) and for the two-sided Riesz-Caputo combination of both growing histories,
 * 1./2.*gamma(2.-alpha)/gamma(2.) * (left + (-1.)**n * right) (equation.py:20-25; nodes 0 and N = NaN).
 * side: frx_side.  'simpson' fields carry N+2 nodes on every side (the left operator's odd rows read node i+1; the
 * mirrored rows of the right operator read node i-1, and the right-sided row of node 0 with odd N is NaN). */
int frx_history_sided(frx_ctx* ctx, int rule, int side, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                      int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, double* d_y, int64_t ldy);
/* Fused compute + gather (sweeps sharded over the GPUs of one box): like frx_history for
 * FRX_SIDE_LEFT, but result row (a*n_fields + f) is stored at row row0 + a*n_fields + f of EVERY buffer in
 * h_peer_ptrs (host array of n_peers device pointers: this GPU's and its NVLink peers' copies of the gathered
 * [rows][ldy] array, mapped by the caller through CUDA IPC / symmetric memory).  The stores are issued by the
 * history kernel's epilogue tile by tile; the caller only needs a barrier before reading the gathered array. */
#define FRX_MAX_PEERS 8
int frx_history_gather(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                       int64_t n_fields, const double* d_g, int64_t ldg, int64_t n_g, int32_t n_peers,
                       const uint64_t* h_peer_ptrs, int64_t row0, int64_t ldy);
int frx_history_host(frx_ctx* ctx, int rule, int64_t n_alpha, const double* h_alpha, double h, int64_t N,
                     int64_t n_fields, const double* h_g, int64_t ldg, int64_t n_g, double* h_y, int64_t ldy);

/* ---- inverse of the growing-history operator (EXTENSION: the reference has no solver) -------------
 * For every alpha: given f_1..f_N (d_f[a*ldf + i]) and the initial value u0[a], find u with
 *   frx_history(rule, alpha, h, u)[i] = f_i, i = 1..N     (u_0 = u0; rows = Settings(alpha, i*h, i) stencils),
 * by Newton iteration for the inverse series on the tensor-core Toeplitz kernel (csrc/frx_ode.cu).  One implicit step of a
 * fractional ODE / Volterra equation; 'caputo', 'trapezoidal' and the 'grunwald_letnikov' extension (the implicit GL scheme).  d_u[a*ldu + i], i = 0..N. */
int frx_history_solve(frx_ctx* ctx, int rule, int64_t n_alpha, const double* d_alpha, double h, int64_t N,
                      const double* d_f, int64_t ldf, const double* d_u0, double* d_u, int64_t ldu);

/* ---- K3: batched Toeplitz application (one alpha, many fields) ----------------------------------
 * Y[f*ldy + i] = sum_j T[i][j] * G[f*ldg + j], T the (N+1)x(N+1) lower-triangular growing-history operator of
 * (rule, alpha, h) -- the dense contraction of BASELINE cfg 3 (N = 16384 x 4096 fields).  Fields are stored
 * field-major (each field contiguous along the node index, like frx_history); T is never materialised: the
 * kernel builds its FP64 tensor-core (DMMA m8n8k4) A-fragments straight from the 1-D weight table.
 * Same call pattern as above: row i of T is the reference stencil of Settings(alpha, i*h, i). */
int frx_toeplitz_apply(frx_ctx* ctx, int rule, double alpha, double h, int64_t N, int64_t n_fields,
                       const double* d_G, int64_t ldg, int64_t n_g, double* d_Y, int64_t ldy);

/* ---- plain lower-triangular Toeplitz products with caller-supplied sequences (extension) --------
 * y[(s, f)][i] = scale * sum_{j=0..i} c[s*ldc + i-j] * g[.][j],   i = 0..N-1      (0-based: a truncated product of series)
 * on the same FP64 tensor-core kernel as frx_history -- for weight sequences the four rules do not cover, and the
 * building block of frx_history_solve.  paired == 0: every sequence s < n_seq is applied to every field f < n_fields
 * (output row s*n_fields + f, field f); paired != 0: sequence s is applied to fields s*n_fields .. s*n_fields+n_fields-1
 * (g holds n_seq*n_fields fields).  n_cols > 0 promises g[.][j] == 0 for j >= n_cols and asks for rows i >= n_cols only
 * (rows below may be left untouched): the rectangular block of the product, as Newton's iteration on series needs it.
 * No counterpart in the reference. */
int frx_toeplitz_product(frx_ctx* ctx, int64_t n_seq, const double* d_c, int64_t ldc, int64_t N, int64_t n_fields,
                         int32_t paired, const double* d_g, int64_t ldg, double scale, double* d_y, int64_t ldy,
                         int64_t n_cols);

/* ---- stencil o stencil composition on the device ---------------------------
 * Batched expansion of fdm.Operator(stencil_a, fdm.Operator(stencil_b)) as the reference's call sites build it
 * (test/integration/test_equation.py:61-66): C[o] = sum over A's nodes, in A's order, of wa * wb with off_a + off_b == o,
 * first term assigned, later ones added -- the floating-point operations of the dict-merging expansion, hence
 * bit-identical to it.  Ragged arrays: item k owns entries [ptr[k], ptr[k+1]).  Offsets are integers in a unit chosen by
 * the caller (the finest grid both stencils live on); d_off_b must ascend within an item; the caller provides the
 * output structure (d_ptr_c, d_off_c: the distinct sums, any order), the weights come back in d_w_c. */
int frx_stencil_compose(frx_ctx* ctx, int64_t n_items, const int64_t* d_ptr_a, const int32_t* d_off_a, const double* d_w_a,
                        const int64_t* d_ptr_b, const int32_t* d_off_b, const double* d_w_b, const int64_t* d_ptr_c,
                        const int32_t* d_off_c, int64_t total_c, double* d_w_c);

/* ---- K4: assemble + batch-solve small fractional boundary-value systems -------------------------
 * No counterpart in the reference (upstream `fdm` would do it): PARITY UNPINNED, see DESIGN.md section 6.
 * System s has n unknowns u_1..u_n at x_k = k*h[s] and the matrix of
 *   Operator(central(h), Operator(create_riesz_caputo_stencil(rule, Settings(alpha[s], res[s]*h[s], res[s])),
 *            Operator(central(h))))
 * (reference equation.py:13-25 for the stencil, test/integration/test_equation.py:64 for the central difference)
 * expanded at every interior node, with u = 0 at and beyond both ends.  The matrix is built in shared memory from
 * the 2*res+1 stencil weights and solved there by LU with partial pivoting; it never exists in HBM.
 * Parameter arrays are HOST arrays (3 doubles per system); rhs / x are device arrays, system s at s*ldb / s*ldx.
 * d_info[s] = 0, or k+1 if the k-th pivot is exactly zero (may be NULL).  All systems of a call share n. */
int frx_assemble_solve(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                       const double* h_res, const double* d_rhs, int64_t ldb, double* d_x, int64_t ldx,
                       int32_t* d_info);
/* The same with right-hand sides and solutions in PAGE-LOCKED host memory (pinned or cudaHostRegister'ed), used in place:
 * the kernels read each right-hand side once and write each solution once over PCIe while they solve; returns when the
 * work is queued (frx_ctx_synchronize before reading h_x).  Pageable buffers: FRX_ERR_INVALID_ARG.
 * Same kernels and same bits as frx_assemble_solve, except that systems of band width res + 1 <= 7 skip the
 * thread-per-system kernel (another elimination order of the same pivot-free LDL^T: equal to rounding). */
int frx_assemble_solve_host(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                            const double* h_res, const double* h_rhs, int64_t ldb, double* h_x, int64_t ldx);
/* After frx_assemble_solve: what the last call on this context did with its systems.  Every system is first factorised
 * WITHOUT pivoting (LDL^T: one thread per system at band width <= 7, a block LDL^T per warp above and for what the former
 * finds indefinite): a DEFINITE matrix (every pivot of the sign of the first) ends there -- elimination without
 * interchanges is backward stable for it; an INDEFINITE one is accepted only on its residual, after at most six steps of
 * iterative refinement with the same factors:  ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||)  in max norms -- refined_per_bin
 * counts those; what fails that test is solved again by the partially pivoted band LU -- n_pivoted / pivoted_per_bin count
 * those.  Bins: band width res + 1 <= 7, 15, 23, 31; the per-bin arrays may be NULL and hold -1 when the attempt did not
 * run.  Synchronises the context's stream.  n_pivoted == n_systems when the attempt is switched off (FRX_K4_LDL=0) or the
 * batch took another kernel (bands wider than 31, n > 256).  No counterpart in the reference (it has no solver). */
int frx_solve_stats(frx_ctx* ctx, int64_t* n_systems, int64_t* n_pivoted, int64_t* pivoted_per_bin /*[4]*/,
                    int64_t* refined_per_bin /*[4]*/);
/* verification aid: the assembled matrices, row-major n x n per system, written to d_A (no solve) */
int frx_assemble_dense(frx_ctx* ctx, int rule, int64_t n_sys, int32_t n, const double* h_alpha, const double* h_h,
                       const double* h_res, double* d_A);

/* ---- measurement aid ------------------------------------------------------------------------------
 * Register-only probes of the FP64 DFMA pipe and of mma.sync m8n8k4 f64 (DMMA): the denominators of the
 * FP64 rooflines bench.py reports (MEASURED_PEAKS.json carries no FP64 figure).  Best of `repeats`. */
int frx_measure_fp64_peak(frx_ctx* ctx, int repeats, double* tflops_dfma, double* tflops_dmma);

#ifdef __cplusplus
}
#endif
#endif /* FRX_H */
