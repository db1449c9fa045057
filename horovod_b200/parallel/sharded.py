"""Sharding helpers built from reducescatter / allgather / alltoall (the ZeRO-1, sequence- and expert-parallel
primitives named in SURVEY.md 2.11)."""
import torch


def _hvd():
    import horovod_b200.torch as hvd
    return hvd


def shard_range(numel, rank, size):
    """[start, end) of `rank`'s shard of a flat vector of `numel` elements: the split `hvd.reducescatter` uses along
    dim 0 (earlier ranks get the extra element when `numel` is not divisible)."""
    base, extra = divmod(numel, size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_scatter_flat(flat, op=None, process_set=None, name=None):
    """Sum (or average) `flat` over the set and keep only my shard."""
    hvd = _hvd()
    kw = {} if process_set is None else {'process_set': process_set}
    return hvd.reducescatter(flat.reshape(-1), op=hvd.Sum if op is None else op, name=name, **kw)


def all_gather_flat(shard, process_set=None, name=None):
    """Inverse of `reduce_scatter_flat` for the values: concatenates every rank's shard."""
    hvd = _hvd()
    kw = {} if process_set is None else {'process_set': process_set}
    return hvd.allgather(shard.reshape(-1), name=name, **kw)


def alltoall_rows(x, send_rows=None, process_set=None, name=None):
    """Row exchange (`x[k]` rows go to rank k according to `send_rows`, equal split when None) — the dispatch / combine of
    expert parallelism and the head<->sequence exchange of Ulysses-style sequence parallelism.  Returns
    (received rows, rows received from every rank)."""
    hvd = _hvd()
    kw = {} if process_set is None else {'process_set': process_set}
    n = (process_set or hvd.global_process_set).size()
    if send_rows is None:
        if x.shape[0] % n:
            raise ValueError('equal split needs dim 0 divisible by the set size')
        send_rows = [x.shape[0] // n] * n
    out, received = hvd.alltoall(x, splits=send_rows, name=name, **kw)
    return out, received


class ShardedSGD:
    """ZeRO-1 style SGD with momentum: gradients are reduce-scattered, every rank updates only its shard of the flat
    parameter vector (and keeps momentum only for that shard), updated shards are all-gathered back.

    Demonstrates the primitives end to end and is value-equivalent to `DistributedOptimizer(SGD)` with op=Average."""

    def __init__(self, params, lr, momentum=0.0, process_set=None):
        hvd = _hvd()
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.momentum, self.ps = lr, momentum, process_set
        ps = process_set or hvd.global_process_set
        self.rank, self.size = ps.rank(), ps.size()
        self.numel = sum(p.numel() for p in self.params)
        self.lo, self.hi = shard_range(self.numel, self.rank, self.size)
        self.buf = torch.zeros(self.hi - self.lo, dtype=torch.float32, device=self.params[0].device) if momentum else None
        self.steps = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        hvd = _hvd()
        flat_g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in self.params])
        g = reduce_scatter_flat(flat_g, op=hvd.Average, process_set=self.ps, name='sharded_sgd.grad')
        flat_p = torch.cat([p.detach().reshape(-1).float() for p in self.params])
        mine = flat_p[self.lo:self.hi].clone()
        if self.momentum:
            self.buf.mul_(self.momentum).add_(g)
            g = self.buf
        mine.add_(g, alpha=-self.lr)
        full = all_gather_flat(mine, process_set=self.ps, name='sharded_sgd.param')
        off = 0
        for p in self.params:
            p.copy_(full[off:off + p.numel()].view_as(p).to(p.dtype))
            off += p.numel()
        self.steps += 1
