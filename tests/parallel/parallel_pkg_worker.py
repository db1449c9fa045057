"""np=4 check of horovod_b200.parallel (mesh process sets, local/cross sets, sharded helpers, ShardedSGD)."""
import os

import torch

_r, _n = int(os.environ['HOROVOD_RANK']), int(os.environ['HOROVOD_SIZE'])
if _n == 4:  # present the 4 ranks as 2 hosts x 2
    os.environ.update(HOROVOD_HOSTNAME='fakehost%d' % (_r // 2), HOROVOD_LOCAL_RANK=str(_r % 2), HOROVOD_LOCAL_SIZE='2',
                      HOROVOD_CROSS_RANK=str(_r // 2), HOROVOD_CROSS_SIZE='2')
import horovod_b200.torch as hvd
from horovod_b200 import parallel
from horovod_b200.parallel import sharded

hvd.init()
r, n = hvd.rank(), hvd.size()
assert n == 4
loc = parallel.local_process_set()
cro = parallel.cross_process_set()
assert sorted(loc.ranks) == [2 * (r // 2), 2 * (r // 2) + 1] and sorted(cro.ranks) == [r % 2, r % 2 + 2]
x = torch.ones(3) * (r + 1)
assert hvd.allreduce(x, op=hvd.Sum, process_set=loc, name='loc').tolist() == [float(sum(q + 1 for q in loc.ranks))] * 3
assert hvd.allreduce(x, op=hvd.Sum, process_set=cro, name='cro').tolist() == [float(sum(q + 1 for q in cro.ranks))] * 3
mesh = parallel.mesh_2d(2, 2)
assert (mesh.row, mesh.col) == (r // 2, r % 2)
assert sorted(mesh.row_set.ranks) == [2 * mesh.row, 2 * mesh.row + 1] and sorted(mesh.col_set.ranks) == [mesh.col, mesh.col + 2]
row_sum = hvd.allreduce(x, op=hvd.Sum, process_set=mesh.row_set, name='row')
col_sum = hvd.allreduce(row_sum, op=hvd.Sum, process_set=mesh.col_set, name='col')
assert col_sum.tolist() == [10.0] * 3     # rows then columns = everybody
mesh.release()
try:
    parallel.mesh_2d(3, 2)
    raise AssertionError('bad mesh accepted')
except ValueError:
    pass

# shard ranges tile the vector and agree with reducescatter's split
covered = []
for q in range(n):
    lo, hi = sharded.shard_range(10, q, n)
    covered += list(range(lo, hi))
assert covered == list(range(10)) and sharded.shard_range(10, 0, 4) == (0, 3) and sharded.shard_range(10, 3, 4) == (8, 10)
flat = torch.arange(10, dtype=torch.float32) * (r + 1)
mine = sharded.reduce_scatter_flat(flat, name='rsf')
lo, hi = sharded.shard_range(10, r, n)
assert torch.allclose(mine, torch.arange(10, dtype=torch.float32)[lo:hi] * 10)
back = sharded.all_gather_flat(mine, name='agf')
assert torch.allclose(back, torch.arange(10, dtype=torch.float32) * 10)
rows, got = sharded.alltoall_rows(torch.full((n, 2), float(r)), name='a2a')
assert got.tolist() == [1] * n and rows[:, 0].tolist() == [float(q) for q in range(n)]

# ShardedSGD == DistributedOptimizer(SGD, Average)
def make():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
ma, mb = make(), make()
oa = hvd.DistributedOptimizer(torch.optim.SGD(ma.parameters(), lr=0.1, momentum=0.9), named_parameters=ma.named_parameters())
ob = sharded.ShardedSGD(mb.parameters(), lr=0.1, momentum=0.9)
for step in range(4):
    xb = torch.randn(6, 5, generator=torch.Generator().manual_seed(10 * step + r))
    yb = torch.randn(6, 2, generator=torch.Generator().manual_seed(99 * step + r))
    oa.zero_grad(); torch.nn.functional.mse_loss(ma(xb), yb).backward(); oa.step()
    ob.zero_grad(); torch.nn.functional.mse_loss(mb(xb), yb).backward(); ob.step()
for a, b in zip(ma.parameters(), mb.parameters()):
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
hvd.barrier()
if r == 0:
    print('PARALLEL PKG OK')
hvd.shutdown()
