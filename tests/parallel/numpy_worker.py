"""np=N check of the numpy front end (framework bridge over the torch binding)."""
import numpy as np

import horovod_b200.numpy as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()

x = np.arange(12, dtype=np.float32).reshape(3, 4) * (r + 1)
s = hvd.allreduce(x, op=hvd.Sum, name='np.sum')
assert isinstance(s, np.ndarray) and s.dtype == np.float32
np.testing.assert_allclose(s, np.arange(12, dtype=np.float32).reshape(3, 4) * (n * (n + 1) / 2))
np.testing.assert_allclose(x, np.arange(12, dtype=np.float32).reshape(3, 4) * (r + 1))  # input untouched
a = hvd.allreduce(x.T, name='np.avg.noncontig')  # non-contiguous view
np.testing.assert_allclose(a, x.T * ((n + 1) / 2) / (r + 1), rtol=1e-6)

hs = [hvd.allreduce_async(np.full(5, float(i + r), dtype=np.float64), op=hvd.Sum, name=f'np.async.{i}') for i in range(4)]
for i, h in enumerate(hs):
    np.testing.assert_allclose(hvd.synchronize(h), np.full(5, float(n * i + n * (n - 1) / 2)))

g = hvd.grouped_allreduce([np.ones(3, dtype=np.int32) * (r + 1), np.ones((2, 2), dtype=np.int64)], op=hvd.Sum, name='np.grp')
assert g[0].tolist() == [n * (n + 1) // 2] * 3 and g[1].tolist() == [[n, n], [n, n]]

ag = hvd.allgather(np.full((r + 1, 2), r, dtype=np.int32), name='np.ag')
assert ag.shape == (n * (n + 1) // 2, 2)
assert ag[:, 0].tolist() == [q for q in range(n) for _ in range(q + 1)]

b = hvd.broadcast(np.full(4, r, dtype=np.uint8), root_rank=n - 1, name='np.bc')
assert b.tolist() == [n - 1] * 4
buf = np.full(4, float(r))
hvd.broadcast_(buf, 0, name='np.bc_')
assert buf.tolist() == [0.0] * 4
buf2 = np.full(3, float(r + 1), dtype=np.float32)
hvd.allreduce_(buf2, op=hvd.Max, name='np.max_')
assert buf2.tolist() == [float(n)] * 3

out, rs = hvd.alltoall(np.arange(n * 2, dtype=np.float32) + 100 * r, splits=[2] * n, name='np.a2a')
assert rs.tolist() == [2] * n and out.tolist() == [100.0 * q + 2 * r + j for q in range(n) for j in range(2)]
out2 = hvd.alltoall(np.arange(n, dtype=np.int64) + 10 * r, name='np.a2a.even')
assert out2.tolist() == [10 * q + r for q in range(n)]

rsx = hvd.reducescatter(np.ones((n * 2, 3), dtype=np.float32) * (r + 1), op=hvd.Sum, name='np.rs')
assert rsx.shape == (2, 3) and np.all(rsx == n * (n + 1) / 2)

assert hvd.broadcast_object({'a': r}, root_rank=0) == {'a': 0}
assert hvd.allgather_object(r * 2) == [2 * q for q in range(n)]
hvd.barrier()
if r == 0:
    print('NUMPY OK')
hvd.shutdown()
