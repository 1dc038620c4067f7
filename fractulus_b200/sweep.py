"""Parameter sweeps across the GPUs of one box: one process per GPU (torch.distributed, NCCL over
NVLink), independent problems dealt to ranks in contiguous blocks, no collective on the data path,
one all_gather of the results at the end (BASELINE cfg 5).

    partition(n_units, world, rank)            -> (begin, end) of this rank's contiguous block
    history_sweep(approximation, alphas, h, g) -> y[n_alpha, N+1] for ALL alphas on every rank
"""
import numpy as np
import torch
import torch.distributed as dist

from . import batched


def partition(n_units, world, rank):
    """Contiguous block of units for `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(n_units), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_blocks(local, n_total, group=None):
    """all_gather of per-rank row blocks [n_local, M] of a partition() split into [n_total, M].
    Equal blocks use one all_gather_into_tensor; ragged ones are padded to the largest block."""
    world, rank = _world(group)
    if world == 1:
        return local
    m = local.shape[1]
    sizes = [partition(n_total, world, r) for r in range(world)]
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


_pinned = {}


def _to_host(t):
    """device -> pinned host buffer (cached per shape) -> NumPy view; synchronises."""
    key = (tuple(t.shape), t.dtype)
    buf = _pinned.get(key)
    if buf is None:
        _pinned.clear()
        buf = _pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy()


def history_sweep(approximation, alphas, h, g, N=None, group=None, compute=None, host_result='all'):
    """Growing-history derivative of ONE field for many alpha, sharded over the ranks of `group`.

    alphas: 1-D array-like (the same on every rank); g: NumPy array (host: copied in, result copied
    out, like the reference's user would hold it) or CUDA tensor (result stays on the device).
    host_result (host g only): 'all' = every rank returns the gathered NumPy result; 'root' = only rank 0
    copies it out (the other ranks return None), which avoids world-size-fold PCIe traffic.
    `compute` (tests only) replaces the per-rank kernel call so the sharding/gather logic can run on
    CPU tensors with the gloo backend."""
    world, rank = _world(group)
    alphas = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64).reshape(-1))
    begin, end = partition(len(alphas), world, rank)
    host_io = not torch.is_tensor(g)
    if compute is None:
        dev = torch.device('cuda', torch.cuda.current_device())
        g_dev = torch.as_tensor(np.asarray(g, dtype=np.float64)).to(dev, non_blocking=True) if host_io else g
        local = batched.history(approximation, alphas[begin:end], h, g_dev, N=N) if end > begin else \
            torch.empty((0, (N if N is not None else g_dev.shape[-1] - 1) + 1), dtype=torch.float64, device=dev)
    else:
        local = compute(alphas[begin:end], g)
    full = gather_blocks(local, len(alphas), group)
    if not host_io:
        return full
    if host_result == 'root' and rank != 0:
        return None
    return _to_host(full) if full.is_cuda else full.numpy()
