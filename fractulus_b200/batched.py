"""Tensor-level API over the C ABI (include/frx.h): the batched forms of the reference's path that the
GPU actually runs.  PyTorch tensors are only the buffer carrier (device memory + streams); every
number is produced by the hand-written kernels in fractulus_b200/csrc.

    stencil_weights(approximation, side, alpha, lf, resolution)  -> row_ptr, offsets, weights   (K5)
    weights_table(approximation, alpha, h, N)                    -> WeightTables                 (K1)
    history(approximation, alpha, h, g)                          -> y[n_alpha, n_fields, N+1]    (K1+K2)
    toeplitz_apply(approximation, alpha, h, G)                   -> Y[n_fields, N+1]             (K1+K3)
    history_host(approximation, alpha, h, g)                     -> NumPy in / NumPy out, copies inside
"""
import collections
import ctypes

import os

import numpy as np
import torch

from . import _lib

WeightTables = collections.namedtuple('WeightTables', ('cA', 'cB', 'w0', 'mult', 'cm1', 'N'))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _ctx_for(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('fractulus_b200 computes on CUDA devices only (got %s); there is no CPU path' % device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ctx = _lib.context(idx)
    ctx.set_stream(torch.cuda.current_stream(idx).cuda_stream)
    return ctx, idx


def _as_f64(x, device):
    return torch.as_tensor(x, dtype=torch.float64, device=device).contiguous()


def _check_alpha_host(rule_name, alpha, side='left'):
    """Domain of the rules, enforced on the host like the reference would fail (mirrors frx_check_rule_alpha /
    frx_check_riesz_alpha in csrc/frx_api.cu)."""
    a = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
    if not np.all(np.isfinite(a)) or np.any(a < 0):
        raise ValueError('alpha must be finite and >= 0')
    if rule_name == 'rectangle' and np.any(a > 1.0):
        raise ZeroDivisionError("'rectangle' with alpha > 1: 0.0 cannot be raised to a negative power")   # equation.py:68
    if rule_name == 'trapezoidal' and np.any(a > 2.0):
        raise ZeroDivisionError("'trapezoidal' with alpha > 2: 0.0 cannot be raised to a negative power")  # equation.py:95
    if rule_name == 'simpson' and np.any(a >= 2.0):
        raise ValueError("'simpson' is only defined for alpha < 2 (gamma(2 - alpha), equation.py:114-139)")
    if side == 'riesz_caputo' and np.any(a >= 2.0):
        raise ValueError('Riesz-Caputo needs alpha < 2 (gamma(2. - alpha), equation.py:23)')
    return a


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError('no CUDA device: fractulus_b200 has no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def stencil_layouts(approximation, side, alpha, resolution):
    """Integer layouts of a batch of stencils in ONE C call: (row_ptr[int64, n+1], first_offset[int32, n]), host arrays."""
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(alpha, dtype=np.float64)))
    r = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(np.asarray(resolution, dtype=np.float64)), a.shape))
    row_ptr = np.zeros(len(a) + 1, dtype=np.int64)
    first = np.zeros(len(a), dtype=np.int32)
    bad = ctypes.c_int64(-1)
    _lib.check(_lib.lib().frx_stencil_layout_batch(_lib.RULES[approximation], _lib.SIDES[side], len(a),
                                                   a.ctypes.data_as(_lib.c_double_p), r.ctypes.data_as(_lib.c_double_p),
                                                   first.ctypes.data_as(_lib.c_int32_p), row_ptr.ctypes.data_as(_lib.c_int64_p),
                                                   ctypes.byref(bad)))
    return row_ptr, first


def stencil_weights(approximation, side, alpha, lf, resolution, device=None):
    """K5 for a batch of Settings triples (1-D array-likes of equal length, host or device).
    Returns (row_ptr[int64, host numpy], offsets[int32 cuda], weights[float64 cuda])."""
    rule = _lib.RULES[approximation]
    sd = _lib.SIDES[side]
    device = torch.device(device) if device is not None else default_device()
    a_h = _check_alpha_host(approximation, alpha.cpu().numpy() if torch.is_tensor(alpha) else alpha, side)
    r_h = np.atleast_1d(np.asarray(resolution.cpu().numpy() if torch.is_tensor(resolution) else resolution,
                                   dtype=np.float64))
    n = len(a_h)
    L = _lib.lib()
    row_ptr, _ = stencil_layouts(approximation, side, a_h, r_h)
    ctx, idx = _ctx_for(device)
    d_alpha, d_lf, d_res = _as_f64(a_h, device), _as_f64(np.atleast_1d(lf.cpu().numpy() if torch.is_tensor(lf) else lf), device), _as_f64(r_h, device)
    offsets = torch.empty(int(row_ptr[-1]), dtype=torch.int32, device=device)
    weights = torch.empty(int(row_ptr[-1]), dtype=torch.float64, device=device)
    _lib.check(L.frx_stencil_weights(ctx.handle, rule, sd, n, _ptr(d_alpha), _ptr(d_lf), _ptr(d_res),
                                     row_ptr.ctypes.data_as(_lib.c_int64_p), _ptr(offsets), _ptr(weights)))
    return row_ptr, offsets, weights


def weights_table(approximation, alpha, h, N, device=None):
    """K1: Toeplitz tables of the growing-history operator for every alpha (see include/frx.h)."""
    rule = _lib.RULES[approximation]
    device = torch.device(device) if device is not None else default_device()
    a_h = _check_alpha_host(approximation, alpha.cpu().numpy() if torch.is_tensor(alpha) else alpha)
    n_alpha = len(a_h)
    L = _lib.lib()
    ldc = L.frx_table_ldc(int(N))
    ldw = N + 1
    ctx, idx = _ctx_for(device)
    d_alpha = _as_f64(a_h, device)
    cA = torch.empty((n_alpha, ldc), dtype=torch.float64, device=device)
    cB = torch.empty((n_alpha, ldc), dtype=torch.float64, device=device) if approximation == 'simpson' else None
    w0 = torch.empty((n_alpha, ldw), dtype=torch.float64, device=device)
    mult = torch.empty((n_alpha, ldw), dtype=torch.float64, device=device)
    cm1 = torch.empty((n_alpha,), dtype=torch.float64, device=device)
    _lib.check(L.frx_weights_table(ctx.handle, rule, n_alpha, _ptr(d_alpha), float(h), int(N), _ptr(cA), _ptr(cB), ldc,
                                   _ptr(w0), _ptr(mult), ldw, _ptr(cm1)))
    return WeightTables(cA, cB, w0, mult, cm1, N)


def history(approximation, alpha, h, g, N=None, out=None, validate=True, side='left'):
    """K1 + K2: y[a, f, i] = (D^alpha_a g_f)(x_i), i = 0..N, of the growing-history operator whose row i
    is the reference stencil of Settings(alpha, i*h, i) (reference test/integration/test_equation.py:51-66).

    alpha: scalar / 1-D (host or cuda); g: cuda float64 [n_g] or [n_fields, n_g]; N defaults to
    n_g - 1 (n_g - 2 for 'simpson', whose odd rows read one node to the right).
    side: 'left' (default), 'right' (row i = create_right_stencil(Settings(alpha, (N-i)*h, N-i))) or
    'riesz_caputo' (1/2 gamma(2-alpha) (left + (-1)**n right), reference equation.py:20-25)."""
    rule = _lib.RULES[approximation]
    if not torch.is_tensor(g) or not g.is_cuda:
        raise TypeError('g must be a CUDA tensor (use history_host for NumPy buffers)')
    device = g.device
    g2 = g.to(torch.float64)
    squeeze_f = g2.dim() == 1
    g2 = g2.reshape(-1, g2.shape[-1]).contiguous()
    n_fields, n_g = g2.shape
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    if torch.is_tensor(alpha) and alpha.is_cuda and not validate:
        d_alpha = alpha.to(torch.float64).reshape(-1).contiguous()
        squeeze_a = alpha.dim() == 0
    else:
        a_h = _check_alpha_host(approximation, alpha.cpu().numpy() if torch.is_tensor(alpha) else alpha, side)
        squeeze_a = np.ndim(alpha) == 0
        d_alpha = _as_f64(a_h, device)
    n_alpha = d_alpha.numel()
    ctx, idx = _ctx_for(device)
    if out is None:
        out = torch.empty((n_alpha, n_fields, N + 1), dtype=torch.float64, device=device)
    else:
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == n_alpha * n_fields * (N + 1)
    _lib.check(_lib.lib().frx_history_sided(ctx.handle, rule, _lib.SIDES[side], n_alpha, _ptr(d_alpha), float(h), int(N),
                                            n_fields, _ptr(g2), n_g, n_g, _ptr(out), N + 1))
    y = out.view(n_alpha, n_fields, N + 1)
    if squeeze_f:
        y = y[:, 0]
    if squeeze_a:
        y = y[0]
    return y


def history_gather(approximation, alpha, h, g, peer_ptrs, row0, N, validate=False):
    """Fused compute + gather: like history(side='left') but every result row is written into all the buffers in
    `peer_ptrs` (device pointers to each GPU's copy of the gathered [rows, N+1] array) at row `row0 + problem`.
    The caller synchronises the GPUs (sweep.PeerGather.barrier) before reading the gathered array."""
    rule = _lib.RULES[approximation]
    g2 = g.to(torch.float64).reshape(-1, g.shape[-1]).contiguous()
    n_fields, n_g = g2.shape
    if torch.is_tensor(alpha) and alpha.is_cuda and not validate:
        d_alpha = alpha.to(torch.float64).reshape(-1).contiguous()
    else:
        d_alpha = _as_f64(_check_alpha_host(approximation, alpha.cpu().numpy() if torch.is_tensor(alpha) else alpha), g.device)
    ctx, idx = _ctx_for(g.device)
    ptrs = (ctypes.c_uint64 * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    _lib.check(_lib.lib().frx_history_gather(ctx.handle, rule, d_alpha.numel(), _ptr(d_alpha), float(h), int(N), n_fields,
                                             _ptr(g2), n_g, n_g, len(peer_ptrs), ptrs, int(row0), N + 1))


def history_solve(approximation, alpha, h, f, u0):
    """Inverse of history(side='left'): u[a, :] with u[a, 0] = u0[a] and history(approximation, alpha_a, h, u[a])[i] = f[a, i]
    for i = 1..N (f[:, 0] is ignored).  alpha: 1-D (host); f: cuda float64 [n_alpha, N+1]; u0: scalar or [n_alpha].
    'caputo' / 'trapezoidal' / 'grunwald_letnikov' (the implicit GL scheme).  Extension: the reference has no solver."""
    rule = _lib.RULES[approximation]
    a_h = _check_alpha_host(approximation, alpha.cpu().numpy() if torch.is_tensor(alpha) else alpha)
    if not torch.is_tensor(f) or not f.is_cuda or f.dim() != 2 or f.shape[0] != len(a_h):
        raise TypeError('f must be a CUDA tensor [n_alpha, N+1]')
    f = f.to(torch.float64).contiguous()
    n_alpha, n1 = f.shape
    ctx, idx = _ctx_for(f.device)
    d_alpha = _as_f64(a_h, f.device)
    d_u0 = _as_f64(np.broadcast_to(np.asarray(u0.cpu().numpy() if torch.is_tensor(u0) else u0, dtype=np.float64), (n_alpha,)).copy(), f.device)
    u = torch.empty_like(f)
    _lib.check(_lib.lib().frx_history_solve(ctx.handle, rule, n_alpha, _ptr(d_alpha), float(h), n1 - 1, _ptr(f), n1, _ptr(d_u0),
                                            _ptr(u), n1))
    return u


def toeplitz_apply(approximation, alpha, h, G, N=None, out=None):
    """K3: Y[f, i] = sum_j T[i, j] G[f, j] for ONE alpha and many fields (BASELINE cfg 3), T the lower-triangular
    growing-history operator.  G: cuda float64 [n_fields, n_g], field-major; T is never materialised."""
    rule = _lib.RULES[approximation]
    if not torch.is_tensor(G) or not G.is_cuda or G.dim() != 2:
        raise TypeError('G must be a 2-D CUDA tensor [n_fields, n_g]')
    G = G.to(torch.float64).contiguous()
    n_fields, n_g = G.shape
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    ctx, idx = _ctx_for(G.device)
    if out is None:
        out = torch.empty((n_fields, N + 1), dtype=torch.float64, device=G.device)
    _lib.check(_lib.lib().frx_toeplitz_apply(ctx.handle, rule, float(alpha), float(h), int(N), n_fields, _ptr(G), n_g, n_g,
                                             _ptr(out), N + 1))
    return out


def toeplitz_product(c, g, paired=False, scale=1.0, n_cols=0, out=None):
    """Extension: plain lower-triangular Toeplitz products y[s, f, i] = scale * sum_{j<=i} c[s, i-j] g[., j] (0-based) with
    caller-supplied sequences on the tensor-core history kernel.  c: cuda float64 [n_seq, N]; g: [n_fields, N]
    (every sequence with every field -> [n_seq, n_fields, N]) or, paired, [n_seq, n_fields, N] (sequence s with its own
    fields).  n_cols > 0: g is zero from column n_cols on and only rows >= n_cols are wanted."""
    if not (torch.is_tensor(c) and c.is_cuda and torch.is_tensor(g) and g.is_cuda):
        raise TypeError('c and g must be CUDA tensors')
    c = c.to(torch.float64).contiguous()
    g = g.to(torch.float64).contiguous()
    if c.dim() == 1:
        c = c[None]
    n_seq, N = c.shape
    if paired:
        if g.dim() == 2:
            g = g[:, None]
        if g.shape[0] != n_seq or g.shape[-1] != N:
            raise ValueError('paired: g must be [n_seq, n_fields, N]')
        n_fields = g.shape[1]
    else:
        if g.dim() == 1:
            g = g[None]
        if g.shape[-1] != N:
            raise ValueError('g must be [n_fields, N]')
        n_fields = g.shape[0]
    ctx, idx = _ctx_for(c.device)
    if out is None:
        out = torch.zeros((n_seq, n_fields, N), dtype=torch.float64, device=c.device)
    _lib.check(_lib.lib().frx_toeplitz_product(ctx.handle, n_seq, _ptr(c), N, N, n_fields, 1 if paired else 0, _ptr(g), N,
                                               float(scale), _ptr(out), N, int(n_cols)))
    return out


def _compose_structure(d_ptr_a, d_off_a, d_ptr_b, d_off_b):
    """(ptr_c, off_c) of a batched composition with tensor ops on the inputs' device: the distinct offset sums of every item,
    ascending -- every (a, b) pair of every item, then one sort/unique over (item, sum) keys."""
    dev = d_ptr_a.device
    n = d_ptr_a.numel() - 1
    na, nb = d_ptr_a[1:] - d_ptr_a[:-1], d_ptr_b[1:] - d_ptr_b[:-1]
    item_b = torch.repeat_interleave(torch.arange(n, device=dev), nb)
    if d_off_b.numel() > 1:
        same = item_b[1:] == item_b[:-1]
        if bool((same & (d_off_b[1:] <= d_off_b[:-1])).any()):
            raise ValueError("B's offsets must ascend within an item")
    item_a = torch.repeat_interleave(torch.arange(n, device=dev), na)     # entry e of A (item k) meets the nb_k entries of B's item k
    cnt = nb[item_a]
    pair_a = torch.repeat_interleave(torch.arange(item_a.numel(), device=dev), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    pair_b = d_ptr_b[item_a][pair_a] + (torch.arange(pair_a.numel(), device=dev) - start[pair_a])
    sums = d_off_a[pair_a].to(torch.int64) + d_off_b[pair_b].to(torch.int64)
    if sums.numel() == 0:
        return torch.zeros(n + 1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    lo = sums.min()
    span = sums.max() - lo + 1
    keys = torch.unique(item_a[pair_a] * span + (sums - lo))              # sorted: by item, then by offset sum
    item_c = torch.div(keys, span, rounding_mode='floor')
    d_off_c = (keys - item_c * span + lo).to(torch.int32)
    d_ptr_c = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_ptr_c[1:] = torch.cumsum(torch.bincount(item_c, minlength=n), 0)
    return d_ptr_c, d_off_c


def stencil_compose(ptr_a, off_a, w_a, ptr_b, off_b, w_b, device=None):
    """(f)(1): batched stencil o stencil composition, the expansion of fdm.Operator(A, fdm.Operator(B)) with the dict-merge's
    own term order (bit-identical to it).  Ragged inputs (NumPy or CUDA tensors): item k owns entries ptr[k]:ptr[k+1] of
    (off, w); offsets are integers on a common grid; B's offsets ascend within an item.  The output structure -- the distinct
    offset sums of every item, ascending -- is built on the device (pair sums, one sort/unique over (item, sum) keys).
    -> (ptr_c, off_c) NumPy and the composed weights as a CUDA tensor."""
    if len(ptr_a) != len(ptr_b):
        raise ValueError('A and B must hold the same number of items')
    if not torch.is_tensor(off_b) and len(off_b) > 1:      # host inputs: argument errors before any device work
        ob, pb = np.asarray(off_b), np.asarray(ptr_b, dtype=np.int64)
        first = np.zeros(len(ob), dtype=bool)
        first[pb[:-1][pb[:-1] < len(ob)]] = True
        if np.any((np.diff(ob) <= 0) & ~first[1:]):
            raise ValueError("B's offsets must ascend within an item")
    dev = default_device() if device is None else torch.device(device)
    t = lambda x, dt: x.to(device=dev, dtype=dt).contiguous() if torch.is_tensor(x) else \
        torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    d_ptr_a, d_ptr_b = t(ptr_a, torch.int64), t(ptr_b, torch.int64)
    d_off_a, d_off_b = t(off_a, torch.int32), t(off_b, torch.int32)
    d_w_a, d_w_b = t(w_a, torch.float64), t(w_b, torch.float64)
    n = d_ptr_a.numel() - 1
    if d_ptr_b.numel() - 1 != n:
        raise ValueError('A and B must hold the same number of items')
    if n <= 0:
        return np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), torch.empty((0,), dtype=torch.float64, device=dev)
    d_ptr_c, d_off_c = _compose_structure(d_ptr_a, d_off_a, d_ptr_b, d_off_b)
    ctx, idx = _ctx_for(dev)
    w_c = torch.empty((d_off_c.numel(),), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().frx_stencil_compose(ctx.handle, n, _ptr(d_ptr_a), _ptr(d_off_a), _ptr(d_w_a), _ptr(d_ptr_b), _ptr(d_off_b),
                                              _ptr(d_w_b), _ptr(d_ptr_c), _ptr(d_off_c), d_off_c.numel(), _ptr(w_c)))
    return d_ptr_c.cpu().numpy(), d_off_c.cpu().numpy(), w_c


def _solve_params(alpha, h, res, n_sys=None):
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(alpha, dtype=np.float64)))
    n_sys = len(a) if n_sys is None else n_sys
    hh = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.float64), (n_sys,)))
    rr = np.ascontiguousarray(np.broadcast_to(np.asarray(res, dtype=np.float64), (n_sys,)))
    a = np.ascontiguousarray(np.broadcast_to(a, (n_sys,)))
    return a, hh, rr


def assemble_solve(approximation, alpha, h, res, rhs, return_info=False, out=None):
    """K4: for every system s assemble the matrix of d/dx o D^alpha_RC(lf = res*h) o d/dx on n interior nodes
    (u = 0 outside) in shared memory and solve A x = rhs[s] by in-kernel LU with partial pivoting.

    alpha, h, res: scalars or 1-D arrays of length n_sys (host); rhs: cuda float64 [n_sys, n].  -> x [n_sys, n]."""
    rule = _lib.RULES[approximation]
    if not torch.is_tensor(rhs) or not rhs.is_cuda or rhs.dim() != 2:
        raise TypeError('rhs must be a 2-D CUDA tensor [n_sys, n]')
    rhs = rhs.to(torch.float64).contiguous()
    n_sys, n = rhs.shape
    a, hh, rr = _solve_params(alpha, h, res, n_sys)
    ctx, idx = _ctx_for(rhs.device)
    if out is None:
        x = torch.empty_like(rhs)
    else:
        x = out
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and tuple(x.shape) == (n_sys, n)
    info = torch.empty((n_sys,), dtype=torch.int32, device=rhs.device)
    _lib.check(_lib.lib().frx_assemble_solve(ctx.handle, rule, n_sys, n, a.ctypes.data_as(_lib.c_double_p),
                                             hh.ctypes.data_as(_lib.c_double_p), rr.ctypes.data_as(_lib.c_double_p),
                                             _ptr(rhs), n, _ptr(x), n, _ptr(info)))
    return (x, info) if return_info else x


def assemble_dense(approximation, alpha, h, res, n, device=None):
    """Verification aid for K4: the assembled matrices [n_sys, n, n] (row-major) as the solve kernel builds them."""
    rule = _lib.RULES[approximation]
    device = torch.device(device) if device is not None else default_device()
    a, hh, rr = _solve_params(alpha, h, res)
    ctx, idx = _ctx_for(device)
    A = torch.empty((len(a), n, n), dtype=torch.float64, device=device)
    _lib.check(_lib.lib().frx_assemble_dense(ctx.handle, rule, len(a), n, a.ctypes.data_as(_lib.c_double_p),
                                             hh.ctypes.data_as(_lib.c_double_p), rr.ctypes.data_as(_lib.c_double_p), _ptr(A)))
    return A


def pinned_empty(shape):
    """Page-locked float64 host array (NumPy view of a pinned torch tensor): frx_history_host reads / writes such
    buffers in place from the kernels instead of staging them through copies."""
    return torch.empty(tuple(shape), dtype=torch.float64).pin_memory().numpy()


def _f64_c(x, ndim_min=0):
    """x as a C-contiguous float64 ndarray (no copy, no work when it already is one)."""
    if type(x) is np.ndarray and x.dtype == np.float64 and x.flags.c_contiguous and x.ndim >= ndim_min:
        return x
    x = np.ascontiguousarray(x, dtype=np.float64)
    return x.reshape(1) if x.ndim < ndim_min else x


def history_host(approximation, alpha, h, g, N=None, device=None, out=None):
    """Same operator through the host-buffer entry point frx_history_host: NumPy in, NumPy out, the
    host<->device transfers happen inside the call.  `out`: optional C-contiguous float64 array [n_alpha, n_fields, N+1]
    to receive the result (a pinned one, see pinned_empty, is written directly by the kernel)."""
    rule = _lib.RULES[approximation]
    a_h = _f64_c(alpha, 1)
    g_h = _f64_c(g)
    n_g = g_h.shape[-1]
    n_fields = g_h.size // n_g if n_g else 0
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    n_alpha = a_h.shape[0]
    ctx = _lib.context(device)
    if out is None:
        y = np.empty((n_alpha, n_fields, N + 1), dtype=np.float64)
    else:
        y = out
        if not (type(y) is np.ndarray and y.dtype == np.float64 and y.flags.c_contiguous and y.flags.writeable
                and y.size == n_alpha * n_fields * (N + 1)):
            raise ValueError('out must be a writeable C-contiguous float64 array with n_alpha*n_fields*(N+1) elements')
        if y.ndim != 3:
            y = y.reshape(n_alpha, n_fields, N + 1)
    # raw addresses: ndarray.ctypes.data_as costs microseconds per argument, as much as the rest of the call
    _lib.check(_lib.lib().frx_history_host(ctx.handle, rule, n_alpha, a_h.ctypes.data, float(h), int(N), n_fields,
                                           g_h.ctypes.data, n_g, n_g, y.ctypes.data, N + 1))
    if g_h.ndim == 1:
        y = y[:, 0]
    if np.ndim(alpha) == 0:
        y = y[0]
    return y


# ---- host-buffer forms of the many-fields / many-systems calls: chunked, copies overlapped with the kernels -----------------

def _host_tensor(a):
    """torch view of a host array without a copy (pinned / registered memory stays pinned: async copies)."""
    return torch.as_tensor(a) if a.flags.c_contiguous and a.dtype == np.float64 else \
        torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))


def _pipelined(host_in, host_out, chunk, width_in, width_out, run, dev):
    """Rows of host_in -> run(device block) -> rows of host_out in chunks on three streams: chunk k+1 is uploaded while
    chunk k computes and chunk k-1 is downloaded (double-buffered device blocks).  run(d_in, d_out) works on the CURRENT
    stream.  With page-locked host arrays both copies are asynchronous DMA; pageable ones are staged by the driver."""
    n = host_in.shape[0]
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    t_in, t_out = _host_tensor(host_in), torch.as_tensor(host_out)
    d_in = [torch.empty((chunk, width_in), dtype=torch.float64, device=dev) for _ in range(2)]
    d_out = [torch.empty((chunk, width_out), dtype=torch.float64, device=dev) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    free_in = [torch.cuda.Event() for _ in range(2)]
    free_out = [torch.cuda.Event() for _ in range(2)]
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    for k, b in enumerate(range(0, n, chunk)):
        e, slot = min(n, b + chunk), k % 2
        with torch.cuda.stream(s_in):
            if k >= 2:
                s_in.wait_event(free_in[slot])
            d_in[slot][:e - b].copy_(t_in[b:e], non_blocking=True)
            ev_in[slot].record(s_in)
        cur.wait_event(ev_in[slot])
        if k >= 2:
            cur.wait_event(free_out[slot])
        run(d_in[slot][:e - b], d_out[slot][:e - b], b, e)
        free_in[slot].record(cur)
        ev_cmp[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[slot])
            t_out[b:e].copy_(d_out[slot][:e - b], non_blocking=True)
            free_out[slot].record(s_out)
    cur.wait_stream(s_out)
    cur.synchronize()


def toeplitz_apply_host(approximation, alpha, h, G, N=None, out=None, chunk=512, device=None):
    """K3 with host buffers: G NumPy [n_fields, n_g] in, Y NumPy [n_fields, N+1] out (`out`: e.g. a pinned or CUDA-registered
    array, written in place).  Fields travel in chunks of `chunk`: upload, kernel and download of consecutive chunks overlap."""
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    n_fields, n_g = G.shape
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    if out is None:
        out = pinned_empty((n_fields, N + 1))
    if n_fields:
        _pipelined(G, out, max(1, min(n_fields, chunk)), n_g, N + 1,
                   lambda d_in, d_out, b, e: toeplitz_apply(approximation, alpha, h, d_in, N=N, out=d_out), dev)
    return out


def _is_page_locked(a):
    """True for a C-contiguous float64 NumPy array in pinned / CUDA-registered host memory"""
    return (type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous and a.size > 0
            and torch.as_tensor(a).is_pinned())


def assemble_solve_host(approximation, alpha, h, res, rhs, out=None, chunk=32768, device=None):
    """K4 with host buffers: rhs NumPy [n_sys, n] in, x NumPy [n_sys, n] out.

    Page-locked buffers (pinned_empty, a CUDA-registered shared segment) are used IN PLACE by one call of
    frx_assemble_solve_host: every kernel reads a right-hand side once and writes a solution once over PCIe while it solves.
    Pageable buffers (or FRX_HOST_ZERO_COPY=0) travel in chunks of `chunk` systems, copies overlapped with the solves."""
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    n_sys, n = rhs.shape
    a, hh, rr = _solve_params(alpha, h, res, n_sys)
    if out is None:
        out = pinned_empty((n_sys, n))
    if not n_sys:
        return out
    if os.environ.get('FRX_HOST_ZERO_COPY', '1') != '0' and _is_page_locked(rhs) and _is_page_locked(out) \
            and tuple(out.shape) == (n_sys, n):
        ctx, idx = _ctx_for(dev)
        _lib.check(_lib.lib().frx_assemble_solve_host(ctx.handle, _lib.RULES[approximation], n_sys, n,
                                                      a.ctypes.data_as(_lib.c_double_p), hh.ctypes.data_as(_lib.c_double_p),
                                                      rr.ctypes.data_as(_lib.c_double_p), rhs.ctypes.data, n, out.ctypes.data, n))
        ctx.synchronize()
        return out
    _pipelined(rhs, out, max(1, min(n_sys, chunk)), n, n,
               lambda d_in, d_out, b, e: assemble_solve(approximation, a[b:e], hh[b:e], rr[b:e], d_in, out=d_out), dev)
    return out
