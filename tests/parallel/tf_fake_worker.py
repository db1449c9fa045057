"""np=N run of the TensorFlow front end against tests/fakes/tensorflow (see that file's disclaimer)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'fakes'))
import numpy as np
import tensorflow as tf

assert tf.__version__.endswith('fake')
import horovod_b200.tensorflow as hvd
import horovod_b200.tensorflow.keras as hvdk

hvd.init()
r, n = hvd.rank(), hvd.size()

x = tf.constant(np.arange(6, dtype=np.float32).reshape(2, 3) * (r + 1))
np.testing.assert_allclose(hvd.allreduce(x, op=hvd.Sum, name='tf.sum').numpy(), np.arange(6).reshape(2, 3) * n * (n + 1) / 2)
np.testing.assert_allclose(hvd.allreduce(x, name='tf.avg').numpy(), np.arange(6).reshape(2, 3) * (n + 1) / 2, rtol=1e-6)
np.testing.assert_allclose(hvd.allreduce(x, name='tf.avg16', compression=hvd.Compression.fp16).numpy(),
                           np.arange(6).reshape(2, 3) * (n + 1) / 2, rtol=1e-2)
assert hvd.allreduce(x, name='tf.avg16b', compression=hvd.Compression.fp16).dtype == tf.float32
# IndexedSlices -> allgather of values / indices, averaged
sl = tf.IndexedSlices(tf.constant(np.ones((2, 3), np.float32) * (r + 1)), tf.constant(np.array([r, r + 1], np.int64)),
                      dense_shape=tf.constant(np.array([n + 1, 3], np.int64)))
out = hvd.allreduce(sl, name='tf.sparse')
assert isinstance(out, tf.IndexedSlices) and out.values.shape == (2 * n, 3) and out.indices.numpy().tolist() == [q + d for q in range(n) for d in (0, 1)]
np.testing.assert_allclose(out.values.numpy()[:2], np.ones((2, 3)) / n)

g = hvd.grouped_allreduce([tf.constant(np.ones(3, np.float32) * r), tf.constant(np.ones((2, 2), np.float64))], op=hvd.Sum, name='tf.grp')
assert g[0].numpy().tolist() == [n * (n - 1) / 2] * 3 and g[1].numpy().tolist() == [[n, n], [n, n]]

ag = hvd.allgather(tf.constant(np.full((r + 1, 2), r, np.int32)), name='tf.ag')
assert ag.shape == (n * (n + 1) // 2, 2)
b = hvd.broadcast(tf.constant(np.full(3, r, np.int64)), root_rank=n - 1, name='tf.bc')
assert b.numpy().tolist() == [n - 1] * 3
o, rs = hvd.alltoall(tf.constant(np.arange(n, dtype=np.float32) + 10 * r), name='tf.a2a')
assert o.numpy().tolist() == [10.0 * q + r for q in range(n)] and rs.numpy().tolist() == [1] * n
rsx = hvd.reducescatter(tf.constant(np.ones((2 * n, 2), np.float32) * (r + 1)), op=hvd.Sum, name='tf.rs')
assert rsx.shape == (2, 2) and np.all(rsx.numpy() == n * (n + 1) / 2)

vs = [tf.Variable(np.full(4, float(r)), name='a:0'), tf.Variable(np.full((2, 2), float(r * 2)), name='b:0')]
hvd.broadcast_variables(vs, root_rank=0)
assert vs[0].numpy().tolist() == [0.0] * 4 and vs[1].numpy().tolist() == [[0.0, 0.0], [0.0, 0.0]]
assert hvd.rank_op().numpy() == r and hvd.size_op().numpy() == n and hvd.local_rank_op().numpy() == hvd.local_rank()

# DistributedGradientTape: dense + None + sparse-as-dense + a local (unsynchronised) source, grouped
tape = tf.GradientTape()
w = [tf.Variable(np.zeros(3), name='w0:0'), tf.Variable(np.zeros(2), name='w1:0'), tf.Variable(np.zeros(2), name='w2:0'),
     tf.Variable(np.zeros(2), name='local:0')]
tape.canned = [tf.constant(np.ones(3) * (r + 1)), None, tf.constant(np.ones(2) * 2 * (r + 1)), tf.constant(np.ones(2) * n)]
dt = hvd.DistributedGradientTape(tape, groups=2)
dt.register_local_source(w[3])
with dt:
    pass
gr = dt.gradient(None, w)
np.testing.assert_allclose(gr[0].numpy(), np.ones(3) * (n + 1) / 2)
assert gr[1] is None
np.testing.assert_allclose(gr[2].numpy(), np.ones(2) * (n + 1))
np.testing.assert_allclose(gr[3].numpy(), np.ones(2))  # local gradient scaled by 1/size, not reduced
dt2 = hvd.DistributedGradientTape(tape, op=hvd.Sum, sparse_as_dense=True)
tape.canned = [sl]
dense = dt2.gradient(None, [w[0]])[0]
assert dense.shape == (n + 1, 3)

# backward_passes_per_step aggregation helper
calls = []
for cls in (hvd.LocalGradientAggregationHelperEager, hvd.LocalGradientAggregationHelper):
    calls = []
    helper = cls(2, lambda g, v: calls.append(1) or [hvd.allreduce(x, op=hvd.Sum, name='agg') for x in g],
                 average_aggregated_gradients=True)
    helper.register_local_var(w[3])
    part = helper.compute_gradients([tf.constant(np.ones(2) * 2.0), tf.constant(np.ones(2) * n)], [w[1], w[3]])
    assert not helper.synced and not calls                  # first pass of the window: local running sums only
    np.testing.assert_allclose(part[0].numpy(), np.ones(2) * 2.0)
    applied = []
    assert helper.apply_gradients(lambda: applied.append(1), object()) is None and not applied
    res = helper.compute_gradients([tf.constant(np.ones(2) * 4.0), tf.constant(np.ones(2) * n)], [w[1], w[3]])
    assert helper.synced and len(calls) == 1
    np.testing.assert_allclose(res[0].numpy(), np.ones(2) * 3.0 * n)      # (2 + 4) summed over ranks / 2 passes
    np.testing.assert_allclose(res[1].numpy(), np.ones(2))                # local: 2n / n ranks / 2 passes, never reduced
    helper.apply_gradients(lambda: applied.append(1), object())
    assert applied == [1]
    # the next window starts from zero
    part = helper.compute_gradients([tf.constant(np.ones(2) * 1.0), None], [w[1], w[3]])
    np.testing.assert_allclose(part[0].numpy(), np.ones(2)) and part[1] is None
    helper.compute_gradients([tf.constant(np.ones(2) * 1.0), None], [w[1], w[3]])
try:
    hvd.LocalGradientAggregationHelperEager(2, lambda g, v: g).compute_gradients([sl], [w[0]])
    raise SystemExit('IndexedSlices must be rejected without sparse_as_dense')
except ValueError:
    pass
from horovod_b200.tensorflow import functions, gradient_aggregation, gradient_aggregation_eager, mpi_ops, util  # module paths of the reference
assert mpi_ops.allgather is hvd.allgather and functions.broadcast_variables is hvd.broadcast_variables
assert hvd.handle_average_backwards_compatibility(None, True) == hvd.Average and hvd.handle_average_backwards_compatibility(None, None) == hvd.Average
assert util.refs_to_vars(util.vars_to_refs([w[0]]))[0] is w[0] or hasattr(w[0], 'ref')
assert gradient_aggregation.apply_op_to_not_none_tensors(lambda t, k: t * k, [None, 2], 3) == [None, 6]
assert hvd.broadcast_object({'a': r}, root_rank=0, session=None)['a'] == 0 and hvd.allgather_object(r, session=None) == list(range(n))

# callbacks
from horovod_b200.tensorflow.keras.callbacks import MetricAverageCallback, BroadcastGlobalVariablesCallback, LearningRateWarmupCallback
logs = {'loss': float(r), 'acc': float(2 * r), 'note': 'text'}
MetricAverageCallback().on_epoch_end(0, logs)
assert abs(logs['loss'] - (n - 1) / 2) < 1e-9 and abs(logs['acc'] - (n - 1)) < 1e-9 and logs['note'] == 'text'


class _Opt:
    def __init__(self):
        self.learning_rate = tf.Variable(np.array(0.1))
        self.momentum = tf.Variable(np.array(0.9))
        self._v = [tf.Variable(np.full(2, float(r)))]

    def variables(self):
        return self._v


class _Model:
    def __init__(self):
        self.variables = [tf.Variable(np.full(3, float(r + 5)))]
        self.optimizer = _Opt()


m = _Model()
cb = BroadcastGlobalVariablesCallback(0)
cb.set_model(m)
cb.on_batch_end(0)
assert m.variables[0].numpy().tolist() == [5.0] * 3 and m.optimizer._v[0].numpy().tolist() == [0.0] * 2
wu = LearningRateWarmupCallback(initial_lr=0.8, warmup_epochs=2, steps_per_epoch=4)
wu.set_model(m)
wu.on_train_begin()
wu.on_epoch_begin(0)
wu.on_batch_begin(0)
lr0 = float(m.optimizer.learning_rate.numpy())
assert abs(lr0 - 0.8 / n * (0.25 * (n - 1) / 2 + 1)) < 1e-9, lr0
wu.on_batch_end(0)
assert abs(float(m.optimizer.momentum.numpy()) - 0.9) < 1e-9
wu.on_epoch_begin(1)
wu.on_batch_begin(3)
assert abs(float(m.optimizer.learning_rate.numpy()) - 0.8) < 1e-9  # epoch 1 + 3/4 + 1/4 == warmup_epochs
# Keras optimizer wrapper: same class name, gradients averaged before the wrapped apply_gradients, local accumulation
class SGD:
    def __init__(self, learning_rate=0.5):
        self.learning_rate = learning_rate
        self.applied = 0

    def get_config(self):
        return {'learning_rate': self.learning_rate}

    @classmethod
    def from_config(cls, cfg):
        return cls(**cfg)

    def apply_gradients(self, grads_and_vars):
        self.applied += 1
        for g, v in grads_and_vars:
            if g is not None:
                v.assign(v.numpy() - self.learning_rate * g.numpy())


dopt = hvdk.DistributedOptimizer(SGD(0.5))
assert type(dopt).__name__ == 'SGD' and isinstance(dopt, SGD) and dopt.learning_rate == 0.5
v1, v2 = tf.Variable(np.zeros(2), name='v1:0'), tf.Variable(np.ones(1), name='v2:0')
dopt.apply_gradients([(tf.constant(np.ones(2) * (r + 1)), v1), (None, v2)])
np.testing.assert_allclose(v1.numpy(), -0.5 * np.ones(2) * (n + 1) / 2)
assert v2.numpy().tolist() == [1.0] and dopt.applied == 1
acc = hvdk.DistributedOptimizer(SGD(1.0), backward_passes_per_step=2, average_aggregated_gradients=True)
v3 = tf.Variable(np.zeros(1), name='v3:0')
assert acc.apply_gradients([(tf.constant(np.ones(1) * 2.0), v3)]) is None and acc.applied == 0 and v3.numpy().tolist() == [0.0]
acc.apply_gradients([(tf.constant(np.ones(1) * 4.0), v3)])
np.testing.assert_allclose(v3.numpy(), [-3.0])   # mean of (2, 4), identical on every rank, averaged over ranks
assert acc.applied == 1
local_opt = hvdk.DistributedOptimizer(SGD(1.0))
v4, v5 = tf.Variable(np.zeros(1), name='v4:0'), tf.Variable(np.zeros(1), name='v5:0')
local_opt.register_local_var(v5)
local_opt.apply_gradients([(tf.constant(np.ones(1) * (r + 1)), v4), (tf.constant(np.ones(1) * n), v5)])
np.testing.assert_allclose(v4.numpy(), [-(n + 1) / 2])
np.testing.assert_allclose(v5.numpy(), [-1.0])   # local variable: not reduced, scaled by 1/size
# PartialDistributedOptimizer: the variables of `local_layers` are registered as local
class _Layer:
    def __init__(self, *variables):
        self.trainable_weights = list(variables)


v6, v7 = tf.Variable(np.zeros(1), name='v6:0'), tf.Variable(np.zeros(1), name='v7:0')
popt = hvdk.PartialDistributedOptimizer(SGD(1.0), local_layers=[_Layer(v7)])
popt.apply_gradients([(tf.constant(np.ones(1) * (r + 1)), v6), (tf.constant(np.ones(1) * n), v7)])
np.testing.assert_allclose(v6.numpy(), [-(n + 1) / 2])
np.testing.assert_allclose(v7.numpy(), [-1.0])
try:
    hvdk.PartialDistributedOptimizer(SGD(1.0), local_layers=[object()])
    raise AssertionError('local_layers must be layers')
except ValueError:
    pass
import horovod_b200.keras as hk
assert hk.PartialDistributedOptimizer is hvdk.PartialDistributedOptimizer and hk.callbacks.BestModelCheckpoint is hvdk.callbacks.BestModelCheckpoint
assert hk.elastic.KerasState is hvdk.elastic.KerasState
# BestModelCheckpoint: saves only on improvement, mode inferred from the metric name
saved = []
best = hvdk.callbacks.BestModelCheckpoint(monitor='val_loss', filepath='/tmp/ckpt-{epoch}', save_fn=lambda model, path: saved.append(path))
best.set_model(object())
for epoch, value in enumerate([1.0, 0.5, 0.7, 0.4]):
    best.on_epoch_end(epoch, {'val_loss': value})
assert saved == ['/tmp/ckpt-1', '/tmp/ckpt-2', '/tmp/ckpt-4'] and best.best == 0.4 and best.best_epoch == 3
assert hvdk.callbacks.BestModelCheckpoint(monitor='val_acc').mode == 'max'

# legacy tf.compat.v1.train.Optimizer: compute_gradients wrapper, local vars, aggregation window, Adasum delta optimizer
legacy = tf.compat.v1.train.GradientDescentOptimizer(0.5)
dopt = hvd.DistributedOptimizer(legacy, op=hvd.Sum)
assert isinstance(dopt, tf.compat.v1.train.Optimizer) and dopt.get_slot_names() == ['momentum']
lv = [tf.Variable(np.zeros(2), name='lv0:0'), tf.Variable(np.zeros(2), name='lv1:0')]
dopt.register_local_var(lv[1])
legacy.canned = [tf.constant(np.ones(2) * (r + 1)), tf.constant(np.ones(2) * n)]
gv = dopt.compute_gradients(None, var_list=lv)
np.testing.assert_allclose(gv[0][0].numpy(), np.ones(2) * n * (n + 1) / 2)       # summed over ranks
np.testing.assert_allclose(gv[1][0].numpy(), np.ones(2))                         # local: n / n, not reduced
dopt.apply_gradients(gv)
np.testing.assert_allclose(lv[0].numpy(), -0.5 * np.ones(2) * n * (n + 1) / 2)
legacy2 = tf.compat.v1.train.GradientDescentOptimizer(1.0)
dopt2 = hvd.DistributedOptimizer(legacy2, op=hvd.Average, backward_passes_per_step=2, average_aggregated_gradients=True)
v2 = [tf.Variable(np.zeros(3), name='agg:0')]
for step, g in enumerate((2.0, 4.0)):
    legacy2.canned = [tf.constant(np.ones(3) * g * (r + 1))]
    dopt2.apply_gradients(dopt2.compute_gradients(None, var_list=v2))
    assert legacy2.applied == (0 if step == 0 else 1)                           # the first pass of the window applies nothing
np.testing.assert_allclose(v2[0].numpy(), -np.ones(3) * 3.0 * (n + 1) / 2)       # mean over ranks of (2+4)/2 * (r+1)
try:
    hvd.DistributedOptimizer(_Opt() if '_Opt' in dir() else object(), op=hvd.Adasum)
    raise SystemExit('Adasum with a Keras optimizer must be rejected')
except ValueError:
    pass
ada = hvd.DistributedOptimizer(tf.compat.v1.train.GradientDescentOptimizer(1.0), op=hvd.Adasum)
av = [tf.Variable(np.zeros(4), name='ada:0')]
ada._optimizer.canned = [tf.constant(np.eye(4)[r % 4] * -1.0)]                    # orthogonal updates: Adasum adds them up
ada.apply_gradients(ada.compute_gradients(None, var_list=av))
expect = np.zeros(4)
for q in range(n):
    expect[q % 4] += 1.0
np.testing.assert_allclose(av[0].numpy(), expect, atol=1e-6)

# elastic callbacks (logic in _keras/elastic.py, reference _keras/elastic.py): commit cadence, batch / epoch bookkeeping
import horovod_b200.keras.elastic as hke
import horovod_b200.tensorflow.keras.elastic as hvdke
from horovod_b200._keras import elastic as keras_elastic_impl


class _State:
    def __init__(self):
        self.commits, self.batch, self.epoch = 0, 0, 0

    def commit(self):
        self.commits += 1


st = _State()
cb = hvdke.CommitStateCallback(st, batches_per_commit=3)
assert isinstance(cb, keras_elastic_impl.CommitStateCallbackImpl) and hke.CommitStateCallback is hvdke.CommitStateCallback
cb.on_train_begin()
for b in range(7):
    cb.on_batch_end(b)
assert st.commits == 2 and cb.batches_remaining == 2
cb.on_epoch_end(0)
assert st.commits == 3 and cb.batches_remaining == 3
ub = hvdke.UpdateBatchStateCallback(st)
ub.params = {'steps': 10}
ub.on_train_begin()
st.batch = 4                                          # restored from the last commit: 4 batches of this epoch are done
ub.on_epoch_begin(0)
assert ub.params['steps'] == 6
ub.on_batch_end(5)
assert st.batch == 5
ub.on_epoch_end(0)
assert st.batch == 0
ue = hvdke.UpdateEpochStateCallback(st)
st.epoch = 3                                          # resumed job: Keras counts from 0 again, the state keeps counting
ue.on_train_begin()
ue.on_epoch_end(0)
assert st.epoch == 4
ue.on_epoch_end(1)
assert st.epoch == 5
assert hvdke.TensorFlowKerasState is hke.TensorFlowKerasState
from horovod_b200._keras import callbacks as kcb
assert kcb.MetricAverageCallbackImpl is kcb.MetricAverageCallback and kcb.LearningRateWarmupCallbackImpl is kcb.LearningRateWarmupCallback
# TF extras
assert hvd.broadcast_object_fn(root_rank=n - 1, name='bofn')({'from': r}) == {'from': n - 1}
assert hvd.check_num_rank_power_of_2(4) and not hvd.check_num_rank_power_of_2(6) and hvd.gpu_available() in (True, False)
hook = hvd.BroadcastGlobalVariablesHook(0)
hook.begin()
hook.after_create_session(None, None)
try:
    hvdk.DistributedOptimizer(SGD(), op=hvd.Sum, gradient_predivide_factor=2.0)
    raise AssertionError('predivide with op != Average must be rejected')
except ValueError:
    pass
assert hvdk.allreduce(np.array([1.0, 2.0]) * (r + 1), name='k.ar', op=hvd.Sum).tolist() == [n * (n + 1) / 2, n * (n + 1.0)]
# elastic state for Keras models: commit / restore / sync through hvd.elastic.TensorFlowKerasState
from horovod_b200.tensorflow.elastic import TensorFlowKerasState, TensorFlowState


class _KModel:
    def __init__(self, value):
        self.variables = [tf.Variable(np.full(3, float(value)), name='kw:0')]
        self.optimizer = None

    def get_weights(self):
        return [v.numpy() for v in self.variables]

    def set_weights(self, ws):
        for v, w in zip(self.variables, ws):
            v.assign(w)


km = _KModel(r + 1)
kst = TensorFlowKerasState(km, epoch=r, batch=5)
kst.sync()                                   # rank 0's weights and values everywhere
assert km.variables[0].numpy().tolist() == [1.0] * 3 and kst.epoch == 0 and kst.batch == 5
kst.epoch = 4
kst.commit() if False else kst.save()
km.variables[0].assign(np.full(3, 9.0))
kst.epoch = 8
kst.restore()
assert km.variables[0].numpy().tolist() == [1.0] * 3 and kst.epoch == 4
vst = TensorFlowState(variables=[tf.Variable(np.full(2, float(r)))], step=r)
vst.sync()
assert vst.variables[0].numpy().tolist() == [0.0, 0.0] and vst.step == 0

# SyncBatchNormalization moments: mean / variance over the GLOBAL batch
from horovod_b200.tensorflow.sync_batch_norm import SyncBatchNormalization
bn = SyncBatchNormalization(name='sbn')
xb = tf.constant(np.full((4, 2), float(r)))          # rank r contributes 4 rows of value r
mean, var = bn._moments(xb, [0])
allv = np.concatenate([np.full((4, 2), float(q)) for q in range(n)])
np.testing.assert_allclose(mean.numpy(), allv.mean(0))
np.testing.assert_allclose(var.numpy(), allv.var(0), atol=1e-12)
hvd.barrier()
if r == 0:
    print('TF FAKE OK')
hvd.shutdown()
