"""np=N run of the MXNet front end against tests/fakes/mxnet (see that file's disclaimer)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'fakes'))
import numpy as np
import mxnet as mx

assert mx.__version__.endswith('fake')
import horovod_b200.mxnet as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
tot = n * (n + 1) / 2
x = mx.nd.array(np.ones((2, 3), np.float32) * (r + 1))
np.testing.assert_allclose(hvd.allreduce(x, average=False, name='mx.sum').asnumpy(), np.ones((2, 3)) * tot)
np.testing.assert_allclose(hvd.allreduce(x, name='mx.avg').asnumpy(), np.ones((2, 3)) * tot / n)
np.testing.assert_allclose(x.asnumpy(), np.ones((2, 3)) * (r + 1))
hvd.allreduce_(x, average=False, name='mx.sum_')
np.testing.assert_allclose(x.asnumpy(), np.ones((2, 3)) * tot)
ts = [mx.nd.array(np.ones(2, np.float32) * (r + 1)), mx.nd.array(np.ones(3, np.float64))]
hvd.grouped_allreduce_(ts, average=False, name='mx.grp')
assert ts[0].asnumpy().tolist() == [tot] * 2 and ts[1].asnumpy().tolist() == [float(n)] * 3
g = hvd.allgather(mx.nd.array(np.full((r + 1, 2), r, np.int32)), name='mx.ag')
assert g.shape == (n * (n + 1) // 2, 2)
b = mx.nd.array(np.full(3, float(r)))
hvd.broadcast_(b, root_rank=n - 1, name='mx.bc')
assert b.asnumpy().tolist() == [float(n - 1)] * 3
a2a = hvd.alltoall(mx.nd.array(np.arange(n, dtype=np.float32) + 10 * r), name='mx.a2a')
assert a2a.asnumpy().tolist() == [10.0 * q + r for q in range(n)]

# DistributedOptimizer: rescale_grad carries 1/size, update sums the gradients
base = mx.optimizer.Optimizer(learning_rate=1.0)
opt = hvd.DistributedOptimizer(base)
assert abs(base.rescale_grad - 1.0 / n) < 1e-12
w, gr = mx.nd.array(np.zeros(2)), mx.nd.array(np.ones(2) * (r + 1))
opt.update(0, w, gr, None)
np.testing.assert_allclose(w.asnumpy(), -np.ones(2) * tot / n)
ws, gs = [mx.nd.array(np.zeros(1)), mx.nd.array(np.zeros(1))], [mx.nd.array(np.ones(1) * (r + 1)), mx.nd.array(np.ones(1))]
hvd.DistributedOptimizer(mx.optimizer.Optimizer(learning_rate=1.0), num_groups=1).update([0, 1], ws, gs, [None, None])
np.testing.assert_allclose(ws[0].asnumpy(), [-tot / n])
np.testing.assert_allclose(ws[1].asnumpy(), [-1.0])
opt.set_learning_rate(0.25)
assert base.lr == 0.25 and opt.lr == 0.25

# DistributedTrainer: params sorted by name, grads summed, _scale = 1/size, fp16 compression round trip
P = mx.gluon.parameter.Parameter
params = {'b': P('b', mx.nd.array(np.zeros(2)), mx.nd.array(np.ones(2) * (r + 1))),
          'a': P('a', mx.nd.array(np.zeros(1)), mx.nd.array(np.ones(1) * 2 * (r + 1))),
          'frozen': P('frozen', mx.nd.array(np.ones(1)), mx.nd.array(np.ones(1)), grad_req='null')}
tr = hvd.DistributedTrainer(params, mx.optimizer.Optimizer(learning_rate=1.0), compression=hvd.Compression.fp16)
assert [p.name for p in tr._params] == ['a', 'b', 'frozen'] and abs(tr._scale - 1.0 / n) < 1e-12
tr.step(1)
np.testing.assert_allclose(params['b'].data().asnumpy(), -np.ones(2) * tot / n, rtol=1e-3)
np.testing.assert_allclose(params['a'].data().asnumpy(), [-2 * tot / n], rtol=1e-3)
assert params['frozen'].data().asnumpy().tolist() == [1.0]

# broadcast_parameters: NDArray dict + gluon parameters incl. deferred initialisation
d = {'w2': mx.nd.array(np.full(2, float(r))), 'w1': mx.nd.array(np.full(1, float(r + 5)))}
hvd.broadcast_parameters(d, root_rank=0)
assert d['w2'].asnumpy().tolist() == [0.0, 0.0] and d['w1'].asnumpy().tolist() == [5.0]
late = P('late')
hvd.broadcast_parameters({'late': late, 'now': P('now', mx.nd.array(np.full(1, float(r))))}, root_rank=0)
late._init_impl(mx.nd.array(np.full(2, float(r + 1))))   # materialises later: broadcast happens right after init
assert late.data().asnumpy().tolist() == [1.0, 1.0]
try:
    hvd.broadcast_parameters([1, 2])
    raise AssertionError('list accepted')
except ValueError:
    pass
# module paths of the reference: horovod.mxnet.{mpi_ops,functions,compression}
from horovod_b200.mxnet import compression, functions, mpi_ops
assert mpi_ops.allreduce_ is hvd.allreduce_ and functions.broadcast_object is hvd.broadcast_object
assert hvd.Compression.fp16 is compression.FP16Compressor and issubclass(compression.NoneCompressor, compression.Compressor)
c, ctx = compression.FP16Compressor.compress(mx.nd.array(np.ones(3, np.float32)))
assert 'float16' in str(c.dtype) and 'float32' in str(compression.FP16Compressor.decompress(c, ctx).dtype)
assert hvd.broadcast_object({'k': r}, root_rank=n - 1)['k'] == n - 1 and hvd.allgather_object(r) == list(range(n))
assert hvd.split_list([1, 2, 3], 2) == [[1, 2], [3]]
hvd.barrier()
if r == 0:
    print('MX FAKE OK')
hvd.shutdown()
