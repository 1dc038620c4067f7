"""Host logic of the multi-GPU sweeps on CPU: world_size-2 gloo processes exercise partition() / partition_weighted(),
the tensor gather (fallback path) and the shared host result array of fractulus_b200/sweep.py with stand-in per-rank
computes (no kernels on CPU; the CUDA registration of the shared segment is skipped)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def test_partition_covers_everything_once():
    from fractulus_b200.sweep import partition
    for n in (0, 1, 2, 7, 8, 10000, 10001):
        for world in (1, 2, 3, 4, 8):
            blocks = [partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_partition_weighted_balances_cost():
    from fractulus_b200.sweep import partition_weighted, solve_cost
    rng = np.random.default_rng(5)
    for world in (1, 2, 4, 8):
        for costs in (solve_cost(1. + np.arange(100000) % 30, 128), rng.uniform(0, 1, 1000) ** 4, np.ones(7), np.zeros(5),
                      np.zeros(0)):
            blocks = [partition_weighted(costs, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == len(costs)
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            if len(costs) >= 1000:
                per = np.array([costs[b:e].sum() for b, e in blocks])
                assert per.max() <= 1.02 * costs.sum() / world + costs.max()


def _fake_compute(alphas, g):
    # stand-in for the per-rank kernel call: row a = alpha_a * g (shape [n_local, n_g])
    return torch.as_tensor(np.outer(alphas, np.asarray(g)))


def _worker(rank, world, port, n_alpha, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fractulus_b200 import sweep
    ok = True
    alphas = np.linspace(0.1, 0.9, n_alpha)
    # tensor path (fallback gather): simpson's default N = n_g - 2 must give the same width on every rank (incl. empty ones)
    g = torch.arange(5, dtype=torch.float64)
    full = sweep.history_sweep('caputo', alphas, 0.1, g, compute=_fake_compute)
    ok &= torch.equal(full, torch.as_tensor(np.outer(alphas, g.numpy()))) and tuple(full.shape) == (n_alpha, 5)
    # host path: one shared segment, every rank writes its rows in place, every rank reads all of it
    gh = np.arange(1., 8.)
    y = sweep.history_sweep('simpson', alphas, 0.1, gh, N=6, compute=_fake_compute, host_result='all')
    ok &= isinstance(y, np.ndarray) and y.shape == (n_alpha, 7) and np.array_equal(y, np.outer(alphas, gh))
    dist.barrier()      # the array IS the sweep's shared segment: the next sweep of that shape overwrites it on every rank
    y2 = sweep.history_sweep('simpson', alphas, 0.1, 2. * gh, N=6, compute=_fake_compute, host_result='root')
    ok &= (y2 is None) if rank else np.array_equal(y2, np.outer(alphas, 2. * gh))
    dist.barrier()
    # fields and systems
    G = np.arange(15.).reshape(5, 3)
    Y = sweep.toeplitz_apply_sweep('caputo', 0.5, 0.1, G, N=2, compute=lambda a, Gb: a * Gb)
    ok &= np.array_equal(Y, 0.5 * G)
    dist.barrier()
    res = 1. + np.arange(9) % 5
    rhs = np.arange(36.).reshape(9, 4)
    X = sweep.assemble_solve_sweep('caputo', 0.5, 0.1, res, rhs, compute=lambda a, h, r, b: b * r[:, None])
    ok &= np.array_equal(X, rhs * res[:, None])
    with np.testing.assert_raises(ValueError):
        sweep.history_sweep('caputo', alphas, 0.1, np.zeros((2, 5)), compute=_fake_compute)   # one field only
    dist.barrier()
    sweep.release_plans()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_alpha', [6, 7, 1])      # equal blocks, ragged blocks, a rank without work
def test_sweeps_world2_gloo(n_alpha):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_alpha, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok for _, ok in res)
