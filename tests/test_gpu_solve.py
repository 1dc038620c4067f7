"""GPU checks for K4 (assemble + batch-solve).  PARITY UNPINNED: the reference has no assembler / solver; the
oracle (oracle/assemble.py) composes the reference's Riesz-Caputo stencils with central differences through
the fdm shim exactly like the reference's tests compose operators, and solves with numpy.linalg.solve."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from fractulus_b200 import batched
from oracle import assemble


def test_assembled_matrix_equals_oracle_composition():
    for rule, res_list in (('caputo', (1, 2, 5, 16)), ('trapezoidal', (3,)), ('rectangle', (4,)), ('simpson', (4,))):
        for res in res_list:
            for alpha in (0.3, 0.8):
                n, h = 40, 1. / 41
                got = batched.assemble_dense(rule, [alpha], h, res, n)[0].cpu().numpy()
                want = assemble.assemble(rule, alpha, h, res, n)
                scale = np.abs(want).max()
                assert np.abs(got - want).max() <= 4 * np.finfo(float).eps * scale, (rule, res, alpha)
                assert np.array_equal(got, got.T)         # Riesz-Caputo operator: symmetric Toeplitz


@pytest.mark.parametrize('n', [1, 2, 16, 64, 128, 200, 256])
def test_solve_vs_numpy_on_oracle_matrix(n):
    rng = np.random.default_rng(n)
    n_sys = 24
    alpha = rng.uniform(0.05, 0.95, n_sys)
    res = rng.integers(1, min(30 if n % 2 == 0 else 40, max(1, n)) + 1, n_sys).astype(np.float64)   # <= 30: streaming kernel; odd n: CTA kernel too
    h = 1. / (n + 1)
    x_nodes = np.arange(1, n + 1) * h
    rhs = np.stack([np.sin(np.pi * x_nodes * (1 + s % 3)) + 0.25 * s for s in range(n_sys)])
    x, info = batched.assemble_solve('caputo', alpha, h, res, torch.as_tensor(rhs, device='cuda'), return_info=True)
    x = x.cpu().numpy()
    assert int(info.abs().sum()) == 0
    for s in range(n_sys):
        want, A = assemble.solve('caputo', alpha[s], h, int(res[s]), n, rhs[s])
        cond = np.linalg.cond(A)
        err = np.abs(x[s] - want).max() / np.abs(want).max()
        assert err <= 1e-12 + 20 * np.finfo(float).eps * cond, (n, s, alpha[s], res[s], err, cond)


@pytest.mark.parametrize('n', [3, 9, 30, 50, 64, 100, 201, 256])
def test_definite_systems_end_in_the_pivot_free_attempt_indefinite_ones_only_on_their_residual(n):
    """Every system is first factorised by the block LDL^T without pivoting (csrc/frx_solve_ldl.cuh).  A definite matrix ends
    there; a pivot of the wrong sign marks the system INDEFINITE: it is kept only if its residual passes
    ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||) after at most six refinement steps, otherwise the partially pivoted band LU
    solves it again.  refined + pivoted must be the number of indefinite matrices (eigenvalues of the kernel's own dense
    assembly), and every path must leave LAPACK-quality residuals and errors; n runs over sizes that are not multiples of
    the panel width."""
    from fractulus_b200 import _lib
    alphas, ress = np.linspace(0.05, 0.95, 19), np.arange(1, 31)
    alpha = np.repeat(alphas, len(ress))
    res = np.tile(ress, len(alphas)).astype(np.float64)
    h = 1. / (n + 1)
    g = torch.Generator(device='cuda').manual_seed(n)
    rhs = torch.randn(len(alpha), n, dtype=torch.float64, device='cuda', generator=g)
    x, info = batched.assemble_solve('caputo', alpha, h, res, rhs, return_info=True)
    total, pivoted, piv_bins, refined_bins = _lib.context(0).solve_stats(per_bin=True)
    assert int(info.abs().sum()) == 0 and torch.isfinite(x).all()
    A = batched.assemble_dense('caputo', alpha, h, res, n)
    ev = torch.linalg.eigvalsh(A)
    indefinite = int(((ev[:, 0] < 0) & (ev[:, -1] > 0)).sum())
    assert total == len(alpha)
    if pivoted != total:                       # (the attempt is on: default configuration)
        assert pivoted == sum(piv_bins) and pivoted + sum(refined_bins) == indefinite, (n, pivoted, refined_bins, indefinite)
    r = torch.bmm(A, x.unsqueeze(-1)).squeeze(-1) - rhs
    scale = (A.abs().amax(dim=(1, 2)) * x.abs().amax(dim=1) + rhs.abs().amax(dim=1)).unsqueeze(-1)
    assert (r.abs() / scale).max().item() < 1e-14
    want = torch.linalg.solve(A, rhs)
    cond = ev.abs().amax(dim=1) / ev.abs().amin(dim=1)
    err = (x - want).abs().amax(dim=1) / want.abs().amax(dim=1)
    assert bool((err <= 1e-12 + 20 * np.finfo(float).eps * cond).all())


def test_solve_other_rules_and_dense_storage_path():
    n, h = 48, 1. / 49
    rhs = torch.ones((3, n), dtype=torch.float64, device='cuda')
    for rule, res in (('trapezoidal', 4), ('rectangle', 5), ('simpson', 6), ('caputo', 40)):   # res=40: 3nb+1 >= n -> dense
        x = batched.assemble_solve(rule, [0.2, 0.5, 0.9], h, res, rhs).cpu().numpy()
        for s, alpha in enumerate((0.2, 0.5, 0.9)):
            want, A = assemble.solve(rule, alpha, h, res, n, np.ones(n))
            assert np.abs(x[s] - want).max() / np.abs(want).max() <= 1e-12 + 20 * np.finfo(float).eps * np.linalg.cond(A)


@pytest.mark.parametrize('n', [64, 128, 256])
def test_full_batch_cfg4_properties(n):
    """BASELINE cfg 4 at full size: 1e5 systems, 400 alpha x 250 length scales (resolution 1..30 -> every band-width bin of
    the blocked kernel), n = 64 / 128 / 256: residual of a sample of systems from EVERY bin through the kernel's own dense
    assembly (checked against the oracle above), finiteness of all solutions, bitwise reproducibility."""
    n_alpha, n_ls = 400, 250
    h = 1. / (n + 1)
    alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
    res = np.tile(1. + (np.arange(n_ls) % 30), n_alpha)
    xg = torch.arange(1, n + 1, dtype=torch.float64, device='cuda') * h
    rhs = torch.sin(np.pi * xg).repeat(n_alpha * n_ls, 1).contiguous()
    x1, info = batched.assemble_solve('caputo', alpha, h, res, rhs, return_info=True)
    x2 = batched.assemble_solve('caputo', alpha, h, res, rhs)
    assert torch.equal(x1, x2)
    assert int(info.abs().sum()) == 0 and torch.isfinite(x1).all()
    idx = np.arange(0, n_alpha * n_ls, 331)        # 302 systems; every resolution 1..30 is among them
    assert set(res[idx]) == set(range(1, 31))
    A = batched.assemble_dense('caputo', alpha[idx], h, res[idx], n)
    r = torch.bmm(A, x1[idx].unsqueeze(-1)).squeeze(-1) - rhs[idx]
    scale = (A.abs().amax(dim=(1, 2)) * x1[idx].abs().amax(dim=1)).unsqueeze(-1)
    assert (r.abs() / scale).max().item() < 1e-13
    # a few systems per bin against the oracle's assembly + LAPACK
    for s in (0, 7, 15, 23, 29, 99999 - 5):
        want, Amat = assemble.solve('caputo', alpha[s], h, int(res[s]), n, rhs[s].cpu().numpy())
        err = np.abs(x1[s].cpu().numpy() - want).max() / np.abs(want).max()
        assert err <= 1e-12 + 20 * np.finfo(float).eps * np.linalg.cond(Amat), (n, s, err)


def test_solve_errors():
    rhs = torch.ones((1, 8), dtype=torch.float64, device='cuda')
    with pytest.raises(ZeroDivisionError):
        batched.assemble_solve('caputo', [0.5], 0.1, 0, rhs)
    with pytest.raises(KeyError):
        batched.assemble_solve('nope', [0.5], 0.1, 2, rhs)
    with pytest.raises(ValueError):
        batched.assemble_solve('simpson', [0.5], 0.1, 3, rhs)     # odd Simpson stencil is not symmetric-closed


def _same_bits_beyond_the_narrowest_bin(host, dev, res):
    """Page-locked host buffers used IN PLACE give the device call's bits, except for the systems of the narrowest band
    bin (nb = res + 1 <= 7): on device buffers those are first tried by the thread-per-system kernel
    (frx_k4_band_scalar), which the in-place host form leaves out (FRX_K4_SCALAR=2 turns it on there too); both
    eliminations are pivot-free LDL^T of the same matrix in a different order, so the solutions agree to rounding."""
    wide = np.asarray(res) + 1 > 7
    assert np.array_equal(host[wide], dev[wide])
    scale = np.abs(dev[~wide]).max(axis=1, keepdims=True)
    assert np.all(np.abs(host[~wide] - dev[~wide]) <= 1e-10 * scale)


def test_host_buffer_forms_equal_device_calls():
    """batched.assemble_solve_host (pageable buffers: chunked, copies overlapped with the kernels; page-locked buffers: used
    in place by the kernels) and toeplitz_apply_host = the device calls."""
    n, n_sys = 64, 5000
    rng = np.random.default_rng(2)
    alpha = rng.uniform(0.1, 0.9, n_sys)
    res = rng.integers(1, 31, n_sys).astype(np.float64)
    h = 1. / (n + 1)
    rhs = rng.standard_normal((n_sys, n))
    want = batched.assemble_solve('caputo', alpha, h, res, torch.as_tensor(rhs, device='cuda')).cpu().numpy()
    got = batched.assemble_solve_host('caputo', alpha, h, res, rhs, chunk=1024)            # pageable in, pinned out
    rhs_pin = batched.pinned_empty(rhs.shape)
    rhs_pin[:] = rhs
    out = batched.pinned_empty(rhs.shape)
    got2 = batched.assemble_solve_host('caputo', alpha, h, res, rhs_pin, out=out, chunk=700)
    assert np.shares_memory(got2, out)
    # (per-chunk batches bin the systems differently, the arithmetic per system is the same)
    assert np.array_equal(got, want)
    _same_bits_beyond_the_narrowest_bin(got2, want, res)
    N, B = 1500, 37
    G = rng.standard_normal((B, N + 2))
    Yd = batched.toeplitz_apply('caputo', 0.4, 1. / N, torch.as_tensor(G, device='cuda'), N=N).cpu().numpy()
    Yh = batched.toeplitz_apply_host('caputo', 0.4, 1. / N, G, N=N, chunk=8)
    assert np.allclose(Yh[:, 1:], Yd[:, 1:], rtol=1e-12, atol=1e-12 * np.abs(Yd[:, 1:]).max())


def test_assemble_solve_host_entry_point_needs_page_locked_buffers():
    """frx_assemble_solve_host uses its buffers in place: pageable memory is refused (the Python layer stages that through
    chunked copies instead), page-locked memory gives the device call's bits."""
    import ctypes
    from fractulus_b200 import _lib
    n, n_sys = 40, 64
    alpha, hh, rr = np.full(n_sys, 0.4), np.full(n_sys, 1. / (n + 1)), 1. + np.arange(n_sys) % 9
    rhs = np.random.default_rng(0).standard_normal((n_sys, n))
    out = np.empty_like(rhs)
    ctx = _lib.context(0)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    rc = _lib.lib().frx_assemble_solve_host(ctx.handle, _lib.RULES['caputo'], n_sys, n, dp(alpha), dp(hh), dp(rr), rhs.ctypes.data, n,
                                            out.ctypes.data, n)
    assert rc != 0 and b'page-locked' in _lib.lib().frx_last_error_string()
    rhs_pin, out_pin = batched.pinned_empty(rhs.shape), batched.pinned_empty(rhs.shape)
    rhs_pin[:] = rhs
    rc = _lib.lib().frx_assemble_solve_host(ctx.handle, _lib.RULES['caputo'], n_sys, n, dp(alpha), dp(hh), dp(rr), rhs_pin.ctypes.data, n,
                                            out_pin.ctypes.data, n)
    assert rc == 0
    ctx.synchronize()
    want = batched.assemble_solve('caputo', alpha, hh, rr, torch.as_tensor(rhs, device='cuda')).cpu().numpy()
    _same_bits_beyond_the_narrowest_bin(out_pin, want, rr)


def test_large_batch_equals_small_batches_and_reports_the_first_invalid_setting():
    """A batch is binned by band width on the host and its bins run as concurrent chains of kernels: the solutions are,
    bit for bit, those of the same systems solved in smaller batches (other bin sizes, other launch shapes), and the
    first invalid setting of a large batch is reported like that setting alone.  (A chunked, multi-threaded form of the
    host pass was measured and dropped: 1.5 ms against 0.9 ms per 1e5 systems on the GPU box's host.)"""
    n, n_sys = 24, 40000
    rng = np.random.default_rng(5)
    alpha = rng.uniform(0.1, 0.9, n_sys)
    res = rng.integers(1, 31, n_sys).astype(np.float64)
    h = 1. / (n + 1)
    rhs = torch.as_tensor(rng.standard_normal((n_sys, n)), device='cuda')
    x, info = batched.assemble_solve('caputo', alpha, h, res, rhs, return_info=True)
    assert int((info != 0).sum()) == 0
    for lo in range(0, n_sys, 10000):
        part = batched.assemble_solve('caputo', alpha[lo:lo + 10000], h, res[lo:lo + 10000], rhs[lo:lo + 10000])
        assert torch.equal(part, x[lo:lo + 10000])
    bad = alpha.copy()
    bad[[31000, 9000, 39999]] = [2.5, -1.0, 3.0]           # the first invalid one (index 9000) decides the message
    with pytest.raises(ValueError) as e_big:
        batched.assemble_solve('caputo', bad, h, res, rhs)
    with pytest.raises(ValueError) as e_small:
        batched.assemble_solve('caputo', bad[9000:9001], h, res[9000:9001], rhs[9000:9001])
    assert str(e_big.value) == str(e_small.value)
    hh = np.full(n_sys, h)
    hh[20000] = 0.
    with pytest.raises(ValueError):
        batched.assemble_solve('caputo', alpha, hh, res, rhs)


def test_kernel_selection_switches_agree():
    """The diagnostic switches select other kernel paths for the same systems (the library reads them once per process, so
    each runs in its own interpreter): one stream instead of concurrent bin chains and another queueing order give the
    default's bits; without the thread-per-system kernel, or with the pivoted LU for everything, the solutions agree within
    the solve's accuracy contract (1e-12 + 20 eps cond(A), here through a relative bound on well-conditioned systems)."""
    import os
    import subprocess
    import sys
    import tempfile
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from fractulus_b200 import batched\n"
        "rng = np.random.default_rng(7)\n"
        "n, n_sys = 40, 3000\n"
        "alpha = rng.uniform(0.1, 0.9, n_sys); res = rng.integers(1, 31, n_sys).astype(np.float64)\n"
        "rhs = torch.as_tensor(rng.standard_normal((n_sys, n)), device='cuda')\n"
        "x = batched.assemble_solve('caputo', alpha, 1. / (n + 1), res, rhs)\n"
        "np.save(sys.argv[1], x.cpu().numpy())\n" % repo)
    outs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, env in (('default', {}), ('one_stream', {'FRX_K4_CONCURRENT': '0'}), ('order', {'FRX_K4_BIN_ORDER': '3210'}),
                          ('no_scalar', {'FRX_K4_SCALAR': '0'}), ('pivoted', {'FRX_K4_LDL': '0'})):
            path = os.path.join(tmp, name + '.npy')
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([sys.executable, '-c', script, path], env=e, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, (name, r.stderr[-2000:])
            outs[name] = np.load(path)
    ref = outs['default']
    assert np.isfinite(ref).all()
    assert np.array_equal(outs['one_stream'], ref) and np.array_equal(outs['order'], ref)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    for name in ('no_scalar', 'pivoted'):
        assert np.all(np.abs(outs[name] - ref) <= 1e-9 * scale), name
