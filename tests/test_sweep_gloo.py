"""Host logic of the multi-GPU sweep on CPU: world_size-2 gloo processes exercise partition() and the
result gather of fractulus_b200/sweep.py with a stand-in per-rank compute (no kernels on CPU)."""
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


def _fake_compute(alphas, g):
    # stand-in for the per-rank kernel call: row a = alpha_a * g (shape [n_local, n_g])
    return torch.as_tensor(np.outer(alphas, g))


def _worker(rank, world, port, n_alpha, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fractulus_b200.sweep import history_sweep
    g = torch.arange(5, dtype=torch.float64)
    alphas = np.linspace(0.1, 0.9, n_alpha)
    full = history_sweep('caputo', alphas, 0.1, g, compute=_fake_compute)
    ok = torch.equal(full, torch.as_tensor(np.outer(alphas, g.numpy())))
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_alpha', [6, 7])      # equal blocks and ragged blocks
def test_sweep_gather_world2_gloo(n_alpha):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_alpha, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok and shape == (n_alpha, 5) for _, ok, shape in res)
