"""Multi-GPU parity check, one process per GPU (launched by tests/test_gpu_multi.py through torch.distributed.run):
every rank checks the WHOLE gathered result of the sweeps against the CPU oracle.

  history_sweep, device field   -> fused P2P gather (frx_history_gather: the history kernel's epilogue stores every row into
                                   every GPU's copy of the symmetric-memory result) -- each rank compares all rows
  history_sweep, host field     -> every rank's kernels store into the one shared pinned host array
  toeplitz_apply_sweep / assemble_solve_sweep (host arrays)
prints one line 'SWEEP_CHECK_OK rank=..' per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from fractulus_b200 import batched, sweep
    from oracle import assemble, clib
    from oracle.gen_golden import smooth_field
    eps = np.finfo(float).eps
    for rule, N, n_alpha in (('caputo', 5000, 2 * world + 1), ('simpson', 3001, world), ('caputo', 100000, world)):
        h = 1. / N
        g = np.array(smooth_field(N + 2, h))
        alphas = np.linspace(0.15, 0.85, n_alpha)
        full = sweep.history_sweep(rule, alphas, h, torch.as_tensor(g, device=dev), N=N)
        used_peer = any(k[0] == 'peer' and v is not False for k, v in sweep._plans.items())
        got = full.cpu().numpy()
        assert got.shape == (n_alpha, N + 1)
        lo = 2 if rule == 'simpson' else 1
        for a in range(n_alpha):          # EVERY rank checks EVERY row of the gathered array
            want = clib.history_full(rule, alphas[a], h, N, g, cr=True)
            s = clib.history_full(rule, alphas[a], h, N, g, cr=True, abs_terms=True)
            err = np.abs(got[a, lo:] - want[lo:])
            slack = np.sqrt(np.arange(lo, N + 1) / 4.) * eps * s[lo:]      # see tests/test_gpu_history.py summation_slack
            assert np.all(err <= 1e-12 * np.abs(want[lo:]) + slack), (rank, rule, N, a, err.max())
        # the fused path agrees with the plain single-GPU call up to the split-tile summation grouping (results are bitwise
        # reproducible per launch shape; the stream-K split points depend on the kernel variant)
        b, e = sweep.partition(n_alpha, world, rank)
        if e > b:
            mine = batched.history(rule, alphas[b:e], h, torch.as_tensor(g, device=dev), N=N).reshape(e - b, N + 1)
            d = (mine[:, lo:] - full[b:e, lo:]).abs().max().item()
            assert d <= 1e-13 * mine[:, lo:].abs().max().item(), (rank, rule, N, d)
        again = sweep.history_sweep(rule, alphas, h, torch.as_tensor(g, device=dev), N=N)      # (collective)
        assert torch.equal(again[:, lo:], torch.as_tensor(got[:, lo:], device=dev))            # deterministic
        dist.barrier()
        # host field in, shared pinned host result out
        yh = sweep.history_sweep(rule, alphas, h, g, N=N, host_result='all')
        assert isinstance(yh, np.ndarray) and np.array_equal(yh, got, equal_nan=True), (rank, rule, N)
        dist.barrier()
    # fields (cfg 3 pattern) and systems (cfg 4 pattern) through the shared host result
    N, B, h = 2050, 6 * world + 1, 1. / 2050
    x = np.arange(N + 2) * h
    k = 3. * (1. + np.arange(B) / B)
    G = k[:, None] * np.cos(k[:, None] * x[None, :]) + 2. * x[None, :]
    Y = sweep.toeplitz_apply_sweep('caputo', 0.45, h, G, N=N)
    want = clib.history_full('caputo', 0.45, h, N, G, cr=True)
    s = clib.history_full('caputo', 0.45, h, N, G, cr=True, abs_terms=True)
    assert np.all(np.abs(Y[:, 1:] - want[:, 1:]) <= 1e-12 * np.abs(want[:, 1:]) + np.sqrt(np.arange(1, N + 1) / 4.) * eps * s[:, 1:])
    dist.barrier()
    n, n_sys = 48, 5 * world + 2
    rng = np.random.default_rng(11)
    alpha = rng.uniform(0.1, 0.9, n_sys)
    res = rng.integers(1, 20, n_sys).astype(np.float64)
    rhs = rng.standard_normal((n_sys, n))
    X = sweep.assemble_solve_sweep('caputo', alpha, 1. / (n + 1), res, rhs)
    for sidx in range(n_sys):
        wx, A = assemble.solve('caputo', alpha[sidx], 1. / (n + 1), int(res[sidx]), n, rhs[sidx])
        err = np.abs(X[sidx] - wx).max() / np.abs(wx).max()
        assert err <= 1e-12 + 20 * eps * np.linalg.cond(A), (rank, sidx, err)
    dist.barrier()
    print('SWEEP_CHECK_OK rank=%d world=%d fused_p2p_gather=%s' % (rank, world, used_peer), flush=True)
    sweep.release_plans()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
