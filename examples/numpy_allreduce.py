"""The numpy front end: collectives on ndarrays, no deep-learning framework in the program.

    hvdrun -np 4 python examples/numpy_allreduce.py
"""
import numpy as np

import horovod_b200.numpy as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
x = np.full(5, r + 1, dtype=np.float64)
total = hvd.allreduce(x, op=hvd.Sum, name='total')
mean = hvd.allreduce(x, name='mean')                       # Average is the default
rows = hvd.allgather(np.full((r + 1, 2), r, dtype=np.int32), name='rows')   # ranks contribute different numbers of rows
word = hvd.broadcast_object({'from': r}, root_rank=n - 1)
hvd.allreduce_(x, op=hvd.Max, name='inplace')              # in place
assert total.tolist() == [n * (n + 1) / 2] * 5 and np.allclose(mean, (n + 1) / 2)
assert rows.shape == (n * (n + 1) // 2, 2) and word == {'from': n - 1} and x.tolist() == [float(n)] * 5
if r == 0:
    print('sum', total[0], 'mean', mean[0], 'gathered rows', rows.shape[0], '| NUMPY EXAMPLE OK')
hvd.shutdown()
