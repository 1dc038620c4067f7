"""Parameter sweeps across the GPUs of one box: one process per GPU (torch.distributed), independent problems dealt
to the ranks in contiguous blocks, NO collective on the data path (BASELINE cfg 3 / 4 / 5).

Where the results go decides the plumbing:

  device results (inputs are CUDA tensors)  the gather is FUSED into the history kernel: the gathered array is symmetric
      memory (PeerGather), the kernel's epilogue stores every finished row into every GPU's copy (P2P stores over
      NVLink, tile by tile, while the other tiles are still computing), a device-side barrier ends the step.
      NCCL all_gather is only the fallback when symmetric memory cannot be set up.
  host results (inputs are NumPy arrays)    the result array is ONE page-locked POSIX shared-memory segment that every
      rank maps and registers with CUDA (SharedHostArray); each rank's kernels store their rows straight into their
      slice of it over that GPU's own PCIe link (frx_history_host on a mapped buffer) -- nothing is funnelled through
      rank 0, nothing crosses NVLink, no staging copy.  Every rank then sees the whole result in place.

    partition(n_units, world, rank)                 contiguous block of equal-cost units
    partition_weighted(costs, world, rank)          contiguous block, balanced by a per-unit cost
    history_sweep(approximation, alphas, h, g)      cfg 5: many alpha, one field          -> y[n_alpha, N+1]
    toeplitz_apply_sweep(approximation, alpha, h, G)   cfg 3: one alpha, many fields      -> Y[n_fields, N+1]
    assemble_solve_sweep(approximation, alpha, h, res, rhs)   cfg 4: many systems         -> x[n_sys, n]
"""
import atexit

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, batched


def partition(n_units, world, rank):
    """Contiguous block of units for `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(n_units), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def partition_weighted(costs, world, rank):
    """Contiguous block [begin, end) of units for `rank` such that every rank gets about the same total cost:
    boundaries sit where the running cost crosses k/world of the total (deterministic, the same on every rank)."""
    c = np.asarray(costs, dtype=np.float64).reshape(-1)
    if c.size == 0:
        return 0, 0
    run = np.concatenate([[0.0], np.cumsum(np.maximum(c, 0.0))])
    total = run[-1]
    if not total > 0.0:
        return partition(c.size, world, rank)
    cuts = [int(np.searchsorted(run, total * k / world, side='left')) for k in range(world + 1)]
    cuts[0], cuts[-1] = 0, c.size
    for k in range(1, world + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts[rank], cuts[rank + 1]


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_blocks(local, n_total, group=None, blocks=None):
    """all_gather of per-rank row blocks [n_local, M] into [n_total, M] (fallback path / CPU tests).
    Equal blocks use one all_gather_into_tensor; ragged ones are padded to the largest block."""
    world, rank = _world(group)
    if world == 1:
        return local
    if local.dim() != 2:
        raise ValueError('gather_blocks needs 2-D row blocks, got %r' % (tuple(local.shape),))
    m = local.shape[1]
    sizes = blocks if blocks is not None else [partition(n_total, world, r) for r in range(world)]
    nmax = max(e - b for b, e in sizes)
    if all(e - b == nmax for b, e in sizes):
        out = torch.empty((n_total, m), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((nmax, m), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * nmax, m), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * nmax:r * nmax + (e - b)] for r, (b, e) in enumerate(sizes)], dim=0)


class PeerGather(object):
    """The gathered result array [n_rows, row_len] of a sweep, allocated as symmetric memory so that every rank
    holds device pointers to every peer's copy: batched.history_gather stores results straight into all copies
    from the history kernel's epilogue (P2P stores over NVLink), and barrier() replaces the collective.
    Raises if symmetric memory is unavailable; callers fall back to gather_blocks (NCCL all_gather)."""

    def __init__(self, n_rows, row_len, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        dev = torch.device('cuda', torch.cuda.current_device())
        self.buffer = symm_mem.empty((n_rows, row_len), dtype=torch.float64, device=dev)
        self.handle = symm_mem.rendezvous(self.buffer, group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.rank, self.world = self.handle.rank, self.handle.world_size
        # this rank's own copy first: frx_history_gather treats entry 0 like any other target
        self.ptrs = [self.ptrs[self.rank]] + [p for r, p in enumerate(self.ptrs) if r != self.rank]

    def barrier(self):
        """All ranks' stores issued before this point are visible in every copy afterwards (device-side barrier
        on the current stream)."""
        self.handle.barrier(channel=0)


class SharedHostArray(object):
    """float64 host array of `shape` in ONE POSIX shared-memory segment mapped by every rank of `group` and page-locked /
    mapped for CUDA rank by rank (pin_rows: every rank registers the rows its own kernels store into), every rank
    reads the whole of it as an ordinary NumPy array.  Collective constructor (rank 0 creates, the others attach).
    register=False (CPU tests with the gloo backend) skips the CUDA registration."""

    _counter = 0

    def __init__(self, shape, group=None, register=True):
        import mmap
        import os
        world, rank = _world(group)
        self.shape = tuple(int(x) for x in shape)
        nbytes = max(mmap.PAGESIZE, int(np.prod(self.shape)) * 8)
        # a file in /dev/shm, unlinked as soon as every rank has mapped it: nothing to clean up, even after a crash
        path = [None]
        if rank == 0:
            SharedHostArray._counter += 1
            path[0] = '/dev/shm/frx_sweep_%d_%d' % (os.getpid(), SharedHostArray._counter)
            fd = os.open(path[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, nbytes)
        if world > 1:
            dist.broadcast_object_list(path, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank != 0:
            fd = os.open(path[0], os.O_RDWR)
        self._mm = mmap.mmap(fd, nbytes)
        os.close(fd)
        if world > 1:
            dist.barrier(group)
        if rank == 0:
            os.unlink(path[0])
        self.array = np.frombuffer(self._mm, dtype=np.float64, count=int(np.prod(self.shape))).reshape(self.shape)
        self._ptr = self.array.ctypes.data if self.array.size else 0
        self._nbytes, self._register, self._registered = nbytes, bool(register), []
        atexit.register(self.close)

    def pin_rows(self, begin, end):
        """Page-lock and map for CUDA the rows [begin, end) -- the part THIS rank's kernels store into (every rank pins only
        its own share: 8 ranks x the whole of an 8 GB result would exceed what the OS lets processes lock).  Returns False
        when the OS refuses; the host entry points then stage the rows through a device buffer instead of writing in place."""
        if not self._register or not self._ptr or end <= begin:
            return False
        import mmap
        row_bytes = int(np.prod(self.shape[1:])) * 8
        lo = (begin * row_bytes) // mmap.PAGESIZE * mmap.PAGESIZE
        hi = min(self._nbytes, -(-(end * row_bytes) // mmap.PAGESIZE) * mmap.PAGESIZE)
        if any(a <= lo and hi <= b for a, b in self._registered):
            return True
        try:
            _lib.check(_lib.lib().frx_host_register(self._ptr + lo, hi - lo))
        except Exception:      # noqa: BLE001 -- locked-memory limit: fall back to staged copies
            return False
        self._registered.append((lo, hi))
        return True

    def close(self):
        if getattr(self, '_mm', None) is None:
            return
        for lo, hi in self._registered:
            try:
                _lib.lib().frx_host_unregister(self._ptr + lo)
            except Exception:
                pass
        self._registered = []
        self.array = None
        mm, self._mm = self._mm, None
        try:
            mm.close()
        except BufferError:      # a result array handed to the caller is still alive: the mapping goes with it
            pass


_plans = {}


def _plan(kind, key, make):
    """Sweep buffers are expensive collective allocations: one per (kind, shape, group), reused by later calls."""
    k = (kind,) + tuple(key)
    if k not in _plans:
        for old in [q for q in _plans if q[0] == kind]:      # one live plan per kind: a new shape replaces the old one
            p = _plans.pop(old)
            if hasattr(p, 'close'):
                p.close()
        _plans[k] = make()
    return _plans[k]


def release_plans():
    """Free the cached sweep buffers (collective: call on every rank)."""
    for k in list(_plans):
        p = _plans.pop(k)
        if hasattr(p, 'close'):
            p.close()


def _host_barrier(group, world):
    if world > 1:
        dist.barrier(group)


def _peer_plan(n_rows, row_len, group, world):
    """PeerGather or None (symmetric memory unavailable on some rank -> every rank falls back together)."""
    def make():
        ok, peer = 1, None
        try:
            peer = PeerGather(n_rows, row_len, group)
        except Exception:      # noqa: BLE001 -- any failure means: use NCCL
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=torch.device('cuda', torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return peer if t.item() == 1 else False
    p = _plan('peer', (n_rows, row_len, id(group)), make)
    return p if p is not False else None


def history_sweep(approximation, alphas, h, g, N=None, group=None, compute=None, host_result='all'):
    """Growing-history derivative of ONE field for many alpha (BASELINE cfg 5), sharded over the ranks of `group`.

    alphas: 1-D array-like, the same on every rank.
    g: CUDA tensor [n_g] -> the gathered CUDA result [n_alpha, N+1] on every rank (fused P2P gather; the returned tensor
       is the sweep's symmetric-memory buffer, overwritten by the next sweep of the same shape);
       NumPy array [n_g] (pinned: read in place) -> the result in a shared pinned host array that every rank's kernels
       wrote in place; host_result 'all': every rank returns it, 'root': rank 0 only (the others return None).  The
       array is the sweep's shared segment, overwritten by the next sweep of the same shape.
    `compute` (tests only) replaces the per-rank kernel call so the sharding logic can run on CPU tensors with gloo."""
    world, rank = _world(group)
    alphas = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64).reshape(-1))
    n_total = len(alphas)
    host_io = not torch.is_tensor(g)
    n_g = int(g.shape[-1])
    if (g.ndim if host_io else g.dim()) != 1:
        raise ValueError('history_sweep takes ONE field [n_g]; use toeplitz_apply_sweep for field batches')
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    N = int(N)
    begin, end = partition(n_total, world, rank)
    mine = alphas[begin:end]
    if host_io:
        shared = _plan('host', (n_total, N + 1, id(group)),
                       lambda: SharedHostArray((n_total, N + 1), group, register=compute is None))
        if end > begin:
            if compute is None:
                shared.pin_rows(begin, end)
                batched.history_host(approximation, mine, h, g, N=N, out=shared.array[begin:end])
            else:
                shared.array[begin:end] = np.asarray(compute(mine, g))
        _host_barrier(group, world)
        return shared.array if (host_result == 'all' or rank == 0) else None
    if compute is not None:
        local = compute(mine, g)
        return gather_blocks(local if local.dim() == 2 else local.reshape(end - begin, N + 1), n_total, group)
    if world == 1:
        return batched.history(approximation, mine, h, g, N=N).reshape(n_total, N + 1)
    peer = _peer_plan(n_total, N + 1, group, world)
    if peer is None:
        dev = g.device
        local = batched.history(approximation, mine, h, g, N=N).reshape(end - begin, N + 1) if end > begin else \
            torch.empty((0, N + 1), dtype=torch.float64, device=dev)
        return gather_blocks(local, n_total, group)
    if end > begin:
        batched.history_gather(approximation, mine, h, g, peer.ptrs, begin, N, validate=True)
    peer.barrier()
    return peer.buffer


def toeplitz_apply_sweep(approximation, alpha, h, G, N=None, group=None, compute=None, host_result='all'):
    """ONE alpha applied to many fields (BASELINE cfg 3: the DMMA dense-contraction path), the FIELDS dealt to the ranks
    in contiguous blocks (the weight table is replicated: 128 KB).

    G: NumPy [n_fields, n_g], the same on every rank -> each rank uploads and transforms its block of fields and stores
       the rows into the shared pinned host result [n_fields, N+1] (see history_sweep for host_result);
       CUDA tensor = THIS RANK'S block of fields -> this rank's block of results (fields are independent: nothing to
       exchange; gather with gather_blocks if the whole batch is wanted on every GPU)."""
    world, rank = _world(group)
    host_io = not torch.is_tensor(G)
    if (G.ndim if host_io else G.dim()) != 2:
        raise ValueError('G must be [n_fields, n_g]')
    n_g = int(G.shape[-1])
    if N is None:
        N = n_g - 1 - (1 if approximation == 'simpson' else 0)
    N = int(N)
    if not host_io:
        return batched.toeplitz_apply(approximation, alpha, h, G, N=N) if compute is None else compute(alpha, G)
    n_fields = int(G.shape[0])
    begin, end = partition(n_fields, world, rank)
    shared = _plan('host', (n_fields, N + 1, id(group)),
                   lambda: SharedHostArray((n_fields, N + 1), group, register=compute is None))
    if end > begin:
        if compute is None:
            shared.pin_rows(begin, end)
            batched.toeplitz_apply_host(approximation, alpha, h, G[begin:end], N=N, out=shared.array[begin:end])
        else:
            shared.array[begin:end] = np.asarray(compute(alpha, G[begin:end]))
    _host_barrier(group, world)
    return shared.array if (host_result == 'all' or rank == 0) else None


def solve_cost(res, n):
    """Relative cost of one assemble+solve: the band-limited elimination does ~ n * nb * (2 nb + 1) multiply-adds."""
    nb = np.minimum(np.asarray(res, dtype=np.float64) + 1.0, max(int(n) - 1, 0))
    return nb * (2.0 * nb + 1.0) + 8.0


def assemble_solve_sweep(approximation, alpha, h, res, rhs, group=None, compute=None, host_result='all'):
    """Assemble + batch-solve many small fractional boundary-value systems (BASELINE cfg 4), the SYSTEMS dealt to the ranks
    in contiguous blocks balanced by their band-elimination cost n * nb^2 (solve_cost).

    alpha, h, res: scalars or 1-D arrays of length n_sys (host), the same on every rank.
    rhs: NumPy [n_sys, n] -> solutions in the shared pinned host array [n_sys, n] (see history_sweep for host_result);
         CUDA tensor [n_sys, n], the same shape on every rank -> each rank solves its block and returns
         (x_block, (begin, end))."""
    world, rank = _world(group)
    host_io = not torch.is_tensor(rhs)
    n_sys, n = int(rhs.shape[0]), int(rhs.shape[1])
    a, hh, rr = batched._solve_params(alpha, h, res, n_sys)
    begin, end = (0, n_sys) if world == 1 else partition_weighted(solve_cost(rr, n), world, rank)
    sl = slice(begin, end)
    if not host_io:
        x = batched.assemble_solve(approximation, a[sl], hh[sl], rr[sl], rhs[sl]) if compute is None else \
            compute(a[sl], hh[sl], rr[sl], rhs[sl])
        return x, (begin, end)
    shared = _plan('host', (n_sys, n, id(group)), lambda: SharedHostArray((n_sys, n), group, register=compute is None))
    if end > begin:
        if compute is None:
            shared.pin_rows(begin, end)
            batched.assemble_solve_host(approximation, a[sl], hh[sl], rr[sl], rhs[sl], out=shared.array[sl])
        else:
            shared.array[sl] = np.asarray(compute(a[sl], hh[sl], rr[sl], rhs[sl]))
    _host_barrier(group, world)
    return shared.array if (host_result == 'all' or rank == 0) else None
