"""ctypes binding of libfrx_b200.so (C ABI declared in include/frx.h).

There is no CPU path: if the shared object has not been built (python -m fractulus_b200.build, or
__graft_entry__.build()) importing this module's ``lib()`` raises, and creating a context without a
CUDA device raises ``RuntimeError`` (FRX_ERR_NO_DEVICE)."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# FRX_B200_LIB: alternative build of the same library (tuning aids such as tools/build_trace_lib.sh)
LIB_PATH = os.environ.get('FRX_B200_LIB') or os.path.join(_HERE, '_build', 'libfrx_b200.so')

FRX_OK = 0
RULES = {'caputo': 0, 'rectangle': 1, 'trapezoidal': 2, 'simpson': 3,   # reference equation.py:179-184
         'grunwald_letnikov': 4}   # extension, not a key of the reference (include/frx.h)
SIDES = {'left': 0, 'right': 1, 'riesz_caputo': 2}

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_void_p = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/frx.h declares
SIGNATURES = {
    'frx_version': (ctypes.c_int, []),
    'frx_last_error_string': (ctypes.c_char_p, []),
    'frx_device_count': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    'frx_ctx_create': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_void_p)]),
    'frx_ctx_destroy': (ctypes.c_int, [c_void_p]),
    'frx_ctx_set_stream': (ctypes.c_int, [c_void_p, c_void_p]),
    'frx_ctx_synchronize': (ctypes.c_int, [c_void_p]),
    'frx_ctx_launch_count': (ctypes.c_int, [c_void_p, c_int64_p]),
    'frx_host_register': (ctypes.c_int, [c_void_p, ctypes.c_uint64]),
    'frx_host_unregister': (ctypes.c_int, [c_void_p]),
    'frx_ctx_set_timing': (ctypes.c_int, [c_void_p, ctypes.c_int]),
    'frx_ctx_kernel_times': (ctypes.c_int, [c_void_p, c_double_p, c_int64_p]),
    'frx_stencil_layout': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                          c_int32_p, c_int32_p]),
    'frx_stencil_layout_batch': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_double_p, c_double_p, c_int32_p,
                                                c_int64_p, c_int64_p]),
    'frx_stencil_weights': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_void_p, c_void_p,
                                           c_void_p, c_int64_p, c_void_p, c_void_p]),
    'frx_stencil_weights_host': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                ctypes.c_double, ctypes.c_int32, c_int32_p, c_double_p, c_int32_p]),
    'frx_table_ldc': (ctypes.c_int64, [ctypes.c_int64]),
    'frx_weights_table': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double,
                                         ctypes.c_int64, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p,
                                         ctypes.c_int64, c_void_p]),
    'frx_history': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double, ctypes.c_int64,
                                   ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    'frx_history_sided': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                         ctypes.c_int64]),
    'frx_history_gather': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double, ctypes.c_int64,
                                          ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64, ctypes.c_int64]),
    'frx_history_host': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double,
                                        ctypes.c_int64, ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                        c_void_p, ctypes.c_int64]),   # host arrays by raw address
    'frx_history_solve': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, c_void_p, ctypes.c_double, ctypes.c_int64,
                                         c_void_p, ctypes.c_int64, c_void_p, c_void_p, ctypes.c_int64]),
    'frx_toeplitz_apply': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                          ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    'frx_toeplitz_product': (ctypes.c_int, [c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int32, c_void_p, ctypes.c_int64, ctypes.c_double, c_void_p, ctypes.c_int64,
                                            ctypes.c_int64]),
    'frx_stencil_compose': (ctypes.c_int, [c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, ctypes.c_int64, c_void_p]),
    'frx_assemble_solve': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int32, c_double_p, c_double_p,
                                          c_double_p, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64, c_void_p]),
    'frx_assemble_solve_host': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int32, c_double_p, c_double_p,
                                               c_double_p, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    'frx_solve_stats': (ctypes.c_int, [c_void_p, c_int64_p, c_int64_p, c_int64_p, c_int64_p]),
    'frx_assemble_dense': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int32, c_double_p, c_double_p,
                                          c_double_p, c_void_p]),
    'frx_measure_fp64_peak': (ctypes.c_int, [c_void_p, ctypes.c_int, c_double_p, c_double_p]),
}

_lib = None
_lock = threading.Lock()


class FrxError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        'fractulus_b200: %s is missing -- build the CUDA library with '
                        '`python -m fractulus_b200.build` (there is no CPU fallback)' % LIB_PATH)
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)     # AttributeError if the library does not export it
                    fn.restype, fn.argtypes = res, args
                _lib = L
    return _lib


def check(status):
    """Map a frx_status to the exception the reference raises in the same situation."""
    if status == FRX_OK:
        return
    msg = lib().frx_last_error_string().decode('utf-8', 'replace')
    if status == -3:
        raise KeyError(msg)                  # reference: _factory[approximation], equation.py:29
    if status == -4:
        raise ZeroDivisionError(msg)         # reference: equation.py:56,64,68,99,105
    if status in (-5, -2):
        raise ValueError(msg)
    if status == -7:
        raise MemoryError(msg)
    raise FrxError('frx status %d: %s' % (status, msg))


class Context(object):
    """One frx_ctx per (thread, device)."""

    def __init__(self, device=0):
        self._h = c_void_p()
        self.device = device
        check(lib().frx_ctx_create(int(device), ctypes.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        check(lib().frx_ctx_set_stream(self._h, c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def synchronize(self):
        check(lib().frx_ctx_synchronize(self._h))

    def launch_count(self):
        n = ctypes.c_int64()
        check(lib().frx_ctx_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def solve_stats(self, per_bin=False):
        """(systems of the last assemble_solve on this context, how many of them went to the pivoted LU
        [, that count per band-width bin, the indefinite systems accepted on their residual after refinement per bin])"""
        n, piv = ctypes.c_int64(), ctypes.c_int64()
        bins, refined = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        check(lib().frx_solve_stats(self._h, ctypes.byref(n), ctypes.byref(piv), bins, refined))
        return (n.value, piv.value, list(bins), list(refined)) if per_bin else (n.value, piv.value)

    KERNEL_CLASSES = ('weights_table', 'pad_fields', 'history_main', 'history_fixup', 'stencil', 'toeplitz',
                      'solve', 'other')

    def set_timing(self, on):
        check(lib().frx_ctx_set_timing(self._h, 1 if on else 0))

    def kernel_times(self):
        """{kernel class: (total ms, launches)} since timing was switched on (synchronises)."""
        ms = (ctypes.c_double * 8)()
        calls = (ctypes.c_int64 * 8)()
        check(lib().frx_ctx_kernel_times(self._h, ms, calls))
        return {k: (ms[i], calls[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def fp64_peak(self, repeats=5):
        """(DFMA TFLOP/s, DMMA TFLOP/s) measured by the register-only probes."""
        a, b = ctypes.c_double(), ctypes.c_double()
        check(lib().frx_measure_fp64_peak(self._h, repeats, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def close(self):
        if self._h:
            lib().frx_ctx_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()


def context(device=None):
    """Thread-local context for ``device`` (default: torch's current device when torch is loaded, else 0)."""
    if device is None:
        device = 0
        import sys
        torch = sys.modules.get('torch')
        if torch is not None and torch.cuda.is_available():
            device = torch.cuda.current_device()
    cache = getattr(_tls, 'ctx', None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        cache[device] = Context(device)
    return cache[device]
