"""Edge cases of the op API: empty / 0-dim / bool tensors, many ops in flight, tensors above the fusion threshold,
name reuse across op types, non-contiguous inputs, uneven allgather with empty contributions."""
import sys

import torch

import horovod_b200.torch as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
dev = torch.device('cuda', hvd.local_rank()) if len(sys.argv) > 1 and sys.argv[1] == 'cuda' else torch.device('cpu')
if dev.type == 'cuda':
    torch.cuda.set_device(dev)

# 0-dim (scalar) tensors
s = hvd.allreduce(torch.tensor(float(r + 1), device=dev), op=hvd.Sum, name='e.scalar')
assert s.dim() == 0 and s.item() == n * (n + 1) / 2
b = hvd.broadcast(torch.tensor(7 + r, device=dev), root_rank=n - 1, name='e.scalar.bc')
assert b.dim() == 0 and b.item() == 7 + n - 1

# empty tensors
e = hvd.allreduce(torch.zeros(0, device=dev), op=hvd.Sum, name='e.empty')
assert e.numel() == 0
e2 = hvd.allreduce(torch.zeros(3, 0, 2, device=dev), name='e.empty3d')
assert tuple(e2.shape) == (3, 0, 2)
g = hvd.allgather(torch.zeros(0, 4, device=dev), name='e.ag.allempty')
assert tuple(g.shape) == (0, 4)
g = hvd.allgather(torch.full((r % 2, 3), float(r), device=dev), name='e.ag.someempty')   # even ranks contribute nothing
assert g.shape[0] == sum(q % 2 for q in range(n)) and all(v % 2 == 1 for v in g[:, 0].tolist())
hvd.broadcast_(torch.zeros(0, device=dev), root_rank=0, name='e.bc.empty')

# bool / uint8 / int8
m = hvd.allreduce(torch.tensor([True, False, r == 0], device=dev), op=hvd.Max, name='e.bool.max')
assert m.dtype == torch.bool and m.tolist() == [True, False, True]
m = hvd.allreduce(torch.tensor([True, r == 0], device=dev), op=hvd.Min, name='e.bool.min')
assert m.tolist() == [True, n == 1]
gb = hvd.allgather(torch.tensor([r % 2 == 0], device=dev), name='e.bool.ag')
assert gb.dtype == torch.bool and gb.tolist() == [q % 2 == 0 for q in range(n)]

# many ops in flight, completion in any order, values by name
hs = {i: hvd.allreduce_async(torch.full((1 + i % 7,), float(i + r), device=dev), op=hvd.Sum, name=f'e.many.{i}') for i in range(300)}
for i in sorted(hs, reverse=True):
    out = hvd.synchronize(hs[i])
    assert out.shape[0] == 1 + i % 7 and out[0].item() == n * i + n * (n - 1) / 2, (i, out)

# tensors above and around the fusion threshold (the test runs with a 64 KiB threshold): no fusion, exact values
big = [torch.full((40000 + 13 * k,), float(k + 1), device=dev) for k in range(4)]           # 160 KB each
hb = [hvd.allreduce_async(t, op=hvd.Sum, name=f'e.big.{k}') for k, t in enumerate(big)]
small = [hvd.allreduce_async(torch.full((10,), float(k), device=dev), op=hvd.Sum, name=f'e.small.{k}') for k in range(20)]
for k, h in enumerate(hb):
    out = hvd.synchronize(h)
    assert out.numel() == 40000 + 13 * k and torch.all(out == (k + 1) * n)
for k, h in enumerate(small):
    assert torch.all(hvd.synchronize(h) == k * n)

# the same user name under different op types does not collide; a name may be reused after completion
x = torch.ones(4, device=dev) * (r + 1)
h1 = hvd.allreduce_async(x, op=hvd.Sum, name='e.same')
h2 = hvd.allgather_async(x, name='e.same')
h3 = hvd.broadcast_async(x, 0, name='e.same')
assert hvd.synchronize(h1)[0].item() == n * (n + 1) / 2 and hvd.synchronize(h2).numel() == 4 * n and hvd.synchronize(h3)[0].item() == 1.0
for _ in range(3):
    assert hvd.allreduce(x, op=hvd.Sum, name='e.same')[0].item() == n * (n + 1) / 2

# non-contiguous input: a clear error, and the documented fix works
nc = torch.ones(4, 6, device=dev).t()
try:
    hvd.allreduce_(nc, name='e.nc')
    raise AssertionError('non-contiguous in-place allreduce must be rejected')
except ValueError as ex:
    assert 'contiguous' in str(ex).lower()
assert hvd.allreduce(nc.contiguous(), op=hvd.Sum, name='e.nc.ok').shape == (6, 4)

# int64 sums beyond 2^53 stay exact (no float detour)
big_i = hvd.allreduce(torch.tensor([2 ** 60 + r], dtype=torch.int64, device=dev), op=hvd.Sum, name='e.i64')
assert big_i.item() == n * 2 ** 60 + n * (n - 1) // 2

hvd.barrier()
if r == 0:
    print('EDGE OK')
hvd.shutdown()
