"""`import horovod_b200.tensorflow as hvd` — TensorFlow 2 front end.

API parity with horovod/tensorflow/__init__.py (allreduce with IndexedSlices→allgather :58-176, grouped_allreduce
:232-380, reducescatter :178-230, broadcast_global_variables :481, DistributedOptimizer :896, DistributedGradientTape
:1125, PartialDistributedGradientTape :1204), functions.py (broadcast_variables :66, broadcast_object :97,
allgather_object :177), mpi_ops.py (rank_op/size_op/... :576-660, join :564), compression.py and elastic.py.

Design: there is NO TensorFlow custom-op library here (the reference's tensorflow/mpi_ops.cc + xla_mpi_ops.cc are ~2500
lines of AsyncOpKernels).  TF tensors are handed to the runtime through the framework bridge (`horovod_b200._bridge`):
DLPack for GPU tensors (the P2P kernels read/write TF's own HBM allocation), the numpy view for host tensors.  Inside
`tf.function` graphs the collectives run as `tf.py_function` nodes; gradients are registered with
`tf.custom_gradient`, with the same backward collectives the reference registers (allreduce↔allreduce,
allgather↔reducescatter-by-slice, broadcast↔reduce-to-root, alltoall↔alltoall with the received splits).

TensorFlow is not part of this image: importing this module without TensorFlow raises ImportError.
"""
try:
    import tensorflow as tf
except ImportError as _e:  # pragma: no cover - exercised only where TF is absent
    raise ImportError('horovod_b200.tensorflow needs TensorFlow >= 2.4 (not installed in this environment); the PyTorch '
                      'front end is horovod_b200.torch') from _e
import warnings

from horovod_b200.common.exceptions import HorovodInternalError, HostsUpdatedInterrupt  # noqa: F401
from horovod_b200.common.util import check_extension, split_list  # noqa: F401
from horovod_b200.tensorflow import mpi_ops as _mpi_ops
from horovod_b200.tensorflow.compression import Compression  # noqa: F401
from horovod_b200.tensorflow.mpi_ops import (  # noqa: F401
    _allreduce, _b, _eager, _normalize_name, _ns, _ops, _process_set_ranks, _reducescatter, _run,
    allgather, alltoall, broadcast, broadcast_, grouped_allgather, handle_average_backwards_compatibility, join,
    local_rank_op, local_size_op, process_set_included_op, rank_op, size_op)
from horovod_b200.tensorflow.util import _cache, _executing_eagerly, _make_subgraph, refs_to_vars, vars_to_refs  # noqa: F401

for _k in _mpi_ops._BASICS:
    globals()[_k] = getattr(_mpi_ops, _k)
del _k


def allreduce(tensor, average=None, device_dense='', device_sparse='', compression=Compression.none, op=None,
              prescale_factor=1.0, postscale_factor=1.0, name=None, process_set=_ops.global_process_set):
    """Dense tensors: allreduce.  tf.IndexedSlices: allgather of values and indices (reference __init__.py:91-120)."""
    if average is not None:
        warnings.warn('`average` is deprecated, use `op`', DeprecationWarning)
        op = _ops.Average if average else _ops.Sum
    op = _ops.Average if op is None else op
    if isinstance(tensor, tf.IndexedSlices):
        if op == _ops.Adasum:
            raise NotImplementedError('The Adasum reduction does not currently support sparse tensors; pass sparse_as_dense=True')
        values = allgather(tensor.values, name=(name + '_values') if name else None, process_set=process_set)
        indices = allgather(tensor.indices, name=(name + '_indices') if name else None, process_set=process_set)
        if op == _ops.Average:
            values = values / tf.cast(process_set.size(), values.dtype)
        return tf.IndexedSlices(values, indices, dense_shape=tensor.dense_shape)
    if op == _ops.Average:
        true_op, post = _ops.Sum, postscale_factor / process_set.size()
    else:
        true_op, post = op, postscale_factor
    comp, ctx = compression.compress(tf.convert_to_tensor(tensor))
    out = _allreduce(comp, name=name, op=true_op, prescale_factor=prescale_factor, postscale_factor=post, process_set=process_set)
    return compression.decompress(out, ctx)


def grouped_allreduce(tensors, average=None, device_dense='', device_sparse='', compression=Compression.none, op=None,
                      prescale_factor=1.0, postscale_factor=1.0, name=None, process_set=_ops.global_process_set):
    """One negotiated group: the members are fused into a single kernel launch (reference __init__.py:232-380)."""
    if not tensors:
        return tensors
    if average is not None:
        op = _ops.Average if average else _ops.Sum
    op = _ops.Average if op is None else op
    if any(isinstance(t, tf.IndexedSlices) for t in tensors):
        return [allreduce(t, compression=compression, op=op, prescale_factor=prescale_factor, postscale_factor=postscale_factor,
                          name=f'{name}_{i}' if name else None, process_set=process_set) for i, t in enumerate(tensors)]
    pairs = [compression.compress(tf.convert_to_tensor(t)) for t in tensors]
    xs = [p[0] for p in pairs]
    name = _normalize_name(name)

    @tf.custom_gradient
    def f(*xs):
        ys = _run(lambda *ts: _b.grouped_allreduce(list(ts), name=name, op=op, prescale_factor=prescale_factor,
                                                   postscale_factor=postscale_factor, process_set=process_set),
                  list(xs), [x.dtype for x in xs], name)
        if not _eager():
            for y, x in zip(ys, xs):
                y.set_shape(x.shape)

        def grad(*dys):
            return grouped_allreduce(list(dys), op=op, prescale_factor=prescale_factor, postscale_factor=postscale_factor,
                                     name=(name + '_grad') if name else None, process_set=process_set)
        return list(ys), grad
    outs = f(*xs)
    return [compression.decompress(o, p[1]) for o, p in zip(outs, pairs)]


def reducescatter(tensor, device_dense='', compression=Compression.none, op=_ops.Average, name=None,
                  process_set=_ops.global_process_set, prescale_factor=1.0, postscale_factor=1.0):
    comp, ctx = compression.compress(tf.convert_to_tensor(tensor))
    out = _reducescatter(comp, name=name, op=op, process_set=process_set, prescale_factor=prescale_factor, postscale_factor=postscale_factor)
    return compression.decompress(out, ctx)


def grouped_reducescatter(tensors, device_dense='', compression=Compression.none, op=_ops.Average, name=None,
                          process_set=_ops.global_process_set, prescale_factor=1.0, postscale_factor=1.0):
    return [reducescatter(t, compression=compression, op=op, name=f'{name}_{i}' if name else None, process_set=process_set,
                          prescale_factor=prescale_factor, postscale_factor=postscale_factor) for i, t in enumerate(tensors)]


from horovod_b200.tensorflow.functions import (  # noqa: E402,F401
    allgather_object, broadcast_global_variables, broadcast_object, broadcast_object_fn, broadcast_variables)


class BroadcastGlobalVariablesHook(getattr(getattr(tf.compat.v1, 'train', None), 'SessionRunHook', object)):
    """TF1 `SessionRunHook` that broadcasts all global variables from `root_rank` once the session exists (reference
    tensorflow/__init__.py:269-303).  With TF2 eager execution use `broadcast_variables` / the Keras callback instead."""

    def __init__(self, root_rank, device=''):
        super().__init__()
        self.root_rank, self.device = root_rank, device
        self.bcast_op = None

    def begin(self):
        graph = tf.compat.v1.get_default_graph() if hasattr(tf.compat.v1, 'get_default_graph') else None
        if self.bcast_op is None or getattr(self.bcast_op, 'graph', None) is not graph:
            variables = tf.compat.v1.global_variables()
            self.bcast_op = tf.group(*[v.assign(broadcast(v, self.root_rank, name='bcast_hook_%d' % i)) for i, v in enumerate(variables)]) \
                if hasattr(tf, 'group') else [v.assign(broadcast(v, self.root_rank, name='bcast_hook_%d' % i)) for i, v in enumerate(variables)]

    def after_create_session(self, session, coord):
        if session is not None and hasattr(session, 'run') and not isinstance(self.bcast_op, list):
            session.run(self.bcast_op)


def check_num_rank_power_of_2(num_rank):
    """Adasum's vector-halving needs a power-of-two number of ranks."""
    from horovod_b200.common.util import num_rank_is_power_2
    return num_rank_is_power_2(num_rank)


def gpu_available():
    from horovod_b200.common.util import gpu_available as _gpu
    return _gpu('torch')


# ---- gradient reduction helpers ---------------------------------------------------------------------------------------------
def _group_indices(variables, groups):
    """Returns a list of index lists.  `groups`: None (one list per variable), int (that many round-robin-free
    contiguous buckets), or list of lists of variables (reference __init__.py:566-629)."""
    n = len(variables)
    if groups is None:
        return [[i] for i in range(n)]
    if isinstance(groups, int):
        if groups <= 0:
            raise ValueError('groups should be a non-negative integer or a list of list of tf.Variable.')
        k = min(groups, n) or 1
        per = (n + k - 1) // k
        return [list(range(s, min(s + per, n))) for s in range(0, n, per)]
    ids = {id(v): i for i, v in enumerate(variables)} if not hasattr(variables[0], 'ref') else {v.ref(): i for i, v in enumerate(variables)}
    key = (lambda v: v.ref()) if hasattr(variables[0], 'ref') else id
    seen, out = set(), []
    for g in groups:
        idx = [ids[key(v)] for v in g if key(v) in ids]
        seen.update(idx)
        if idx:
            out.append(idx)
    out.extend([i] for i in range(n) if i not in seen)
    return out


def _make_allreduce_grads_fn(name, device_dense, device_sparse, compression, sparse_as_dense, op, gradient_predivide_factor,
                             groups, process_set):
    if op == _ops.Average:
        pre, post = 1.0 / gradient_predivide_factor, gradient_predivide_factor
    else:
        pre, post = 1.0, 1.0

    def allreduce_grads(grads, variables=None, use_generic_names=False):
        grads = list(grads)
        if sparse_as_dense:
            grads = [tf.convert_to_tensor(g) if isinstance(g, tf.IndexedSlices) else g for g in grads]
        live = [i for i, g in enumerate(grads) if g is not None]
        if process_set.size() == 1 and process_set.included():
            return grads
        out = list(grads)
        if groups is not None and variables is not None:
            for gi, idx in enumerate(_group_indices([variables[i] for i in live], groups)):
                members = [live[j] for j in idx]
                red = grouped_allreduce([grads[i] for i in members], compression=compression, op=op, prescale_factor=pre,
                                        postscale_factor=post, name=f'{name}_group_{gi}', process_set=process_set)
                for i, r in zip(members, red):
                    out[i] = r
        else:
            for i in live:
                gname = f'{name}_grad_{i}' if use_generic_names or variables is None else f'{name}_{_normalize_name(getattr(variables[i], "name", str(i)))}'
                out[i] = allreduce(grads[i], compression=compression, op=op, prescale_factor=pre, postscale_factor=post,
                                   name=gname, process_set=process_set)
        return out
    return allreduce_grads


from horovod_b200.tensorflow.gradient_aggregation import LocalGradientAggregationHelper  # noqa: E402,F401
from horovod_b200.tensorflow.gradient_aggregation_eager import LocalGradientAggregationHelperEager  # noqa: E402,F401


class _DistributedGradientTape:
    """Delegating wrapper: `.gradient()` reduces what the wrapped tape computed."""

    def __init__(self, tape, allreduce_grads, local_sources=(), scale_local_gradients=True, process_set=_ops.global_process_set):
        self._tape = tape
        self._allreduce_grads = allreduce_grads
        self._local = {(_v.ref() if hasattr(_v, 'ref') else id(_v)) for _v in local_sources}
        self._scale_local = scale_local_gradients
        self._process_set = process_set

    def __getattr__(self, item):
        return getattr(self._tape, item)

    def __enter__(self):
        self._tape.__enter__()
        return self

    def __exit__(self, *exc):
        return self._tape.__exit__(*exc)

    def register_local_source(self, source):
        """Marks a variable whose gradient stays local (reference PartialDistributedGradientTape)."""
        self._local.add(source.ref() if hasattr(source, 'ref') else id(source))

    def gradient(self, target, sources, output_gradients=None, use_generic_names=False):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        grads = list(self._tape.gradient(target, srcs, output_gradients))
        key = (lambda v: v.ref()) if srcs and hasattr(srcs[0], 'ref') else id
        is_local = [key(s) in self._local for s in srcs]
        shared = [i for i, l in enumerate(is_local) if not l]
        reduced = self._allreduce_grads([grads[i] for i in shared], [srcs[i] for i in shared], use_generic_names)
        for i, r in zip(shared, reduced):
            grads[i] = r
        if self._scale_local and any(is_local):
            n = self._process_set.size()
            grads = [g / n if (l and g is not None) else g for g, l in zip(grads, is_local)]
        return grads[0] if single else grads


def DistributedGradientTape(gradtape, device_dense='', device_sparse='', compression=Compression.none, sparse_as_dense=False,
                            op=_ops.Average, gradient_predivide_factor=1.0, num_groups=0, groups=None,
                            process_set=_ops.global_process_set, scale_local_gradients=True):
    if gradient_predivide_factor != 1.0 and op != _ops.Average:
        raise ValueError('gradient_predivide_factor not supported with op != Average')
    if num_groups != 0:
        warnings.warn('Parameter `num_groups` has been replaced by `groups`', DeprecationWarning)
        groups = groups if groups is not None else num_groups
    if groups is not None and not (isinstance(groups, list) or groups > 0):
        raise ValueError('groups should be a non-negative integer or a list of list of tf.Variable.')
    fn = _make_allreduce_grads_fn('DistributedGradientTape', device_dense, device_sparse, compression, sparse_as_dense, op,
                                  gradient_predivide_factor, groups, process_set)
    return _DistributedGradientTape(gradtape, fn, (), scale_local_gradients, process_set)


def PartialDistributedGradientTape(gradtape, device_dense='', device_sparse='', compression=Compression.none,
                                   sparse_as_dense=False, op=_ops.Average, gradient_predivide_factor=1.0, num_groups=0,
                                   groups=None, process_set=_ops.global_process_set, local_layers=None, scale_local_gradients=True):
    """Like DistributedGradientTape, but the variables of `local_layers` keep their local gradients."""
    tape = DistributedGradientTape(gradtape, device_dense, device_sparse, compression, sparse_as_dense, op,
                                   gradient_predivide_factor, num_groups, groups, process_set, scale_local_gradients)
    layers = [] if local_layers is None else (local_layers if isinstance(local_layers, (list, tuple)) else [local_layers])
    for layer in layers:
        for v in layer.trainable_weights:
            tape.register_local_source(v)
    return tape


# ---- legacy (tf.compat.v1.train.Optimizer) wrappers ------------------------------------------------------------------------------
_LegacyOptimizer = getattr(getattr(getattr(tf, 'compat', None), 'v1', None), 'train', None)
_LegacyOptimizer = getattr(_LegacyOptimizer, 'Optimizer', None)

if _LegacyOptimizer is not None:
    class _DistributedOptimizer(_LegacyOptimizer):
        """`compute_gradients` of the wrapped optimizer followed by the allreduce (reference tensorflow/__init__.py:632-735);
        everything else is delegated, so slots and variables are the wrapped optimizer's."""

        def __init__(self, optimizer, name=None, use_locking=False, device_dense='', device_sparse='', compression=Compression.none,
                     sparse_as_dense=False, op=_ops.Average, gradient_predivide_factor=1.0, backward_passes_per_step=1,
                     average_aggregated_gradients=False, groups=None, process_set=_ops.global_process_set,
                     scale_local_gradients=True):
            super().__init__(name=name or 'Distributed%s' % type(optimizer).__name__, use_locking=use_locking)
            self._optimizer = optimizer
            self._process_set, self._scale_local = process_set, scale_local_gradients
            self._allreduce_grads = _make_allreduce_grads_fn(self._name if hasattr(self, '_name') else (name or 'DistributedOptimizer'),
                                                             device_dense, device_sparse, compression, sparse_as_dense, op,
                                                             gradient_predivide_factor, groups, process_set)
            self._local_vars = set()
            self._agg_helper = None
            if backward_passes_per_step > 1:
                self._agg_helper = LocalGradientAggregationHelper(
                    backward_passes_per_step, self._allreduce_grads, sparse_as_dense=sparse_as_dense,
                    average_aggregated_gradients=average_aggregated_gradients, rank=_ops.rank() if _ops.is_initialized() else 0,
                    optimizer_type=LocalGradientAggregationHelper._OPTIMIZER_TYPE_LEGACY, process_set=process_set,
                    scale_local_gradients=scale_local_gradients)

        def register_local_var(self, var):
            """Gradients of `var` stay local (divided by the set size when scale_local_gradients)."""
            if self._agg_helper is not None:
                self._agg_helper.register_local_var(var)
            self._local_vars.add(var.ref() if hasattr(var, 'ref') else id(var))

        def compute_gradients(self, *args, **kwargs):
            pairs = list(self._optimizer.compute_gradients(*args, **kwargs))
            grads, variables = [g for g, _ in pairs], [v for _, v in pairs]
            if self._agg_helper is not None:
                reduced = self._agg_helper.compute_gradients(grads, variables)
            else:
                key = (lambda v: v.ref()) if variables and hasattr(variables[0], 'ref') else id
                shared = [i for i, v in enumerate(variables) if key(v) not in self._local_vars]
                reduced = list(grads)
                for i, r in zip(shared, self._allreduce_grads([grads[i] for i in shared], [variables[i] for i in shared])):
                    reduced[i] = r
                if self._scale_local and len(shared) != len(reduced):
                    n, keep = float(self._process_set.size()), set(shared)
                    reduced = [g if (i in keep or g is None) else g / n for i, g in enumerate(reduced)]
            return list(zip(reduced, variables))

        def apply_gradients(self, *args, **kwargs):
            if self._agg_helper is not None:
                return self._agg_helper.apply_gradients(lambda: self._optimizer.apply_gradients(*args, **kwargs), self._optimizer,
                                                        *args, **kwargs)
            return self._optimizer.apply_gradients(*args, **kwargs)

        def get_slot(self, *args, **kwargs):
            return self._optimizer.get_slot(*args, **kwargs)

        def get_slot_names(self, *args, **kwargs):
            return self._optimizer.get_slot_names(*args, **kwargs)

        def variables(self, *args, **kwargs):
            return self._optimizer.variables(*args, **kwargs)

    class _DistributedAdasumOptimizer(_LegacyOptimizer):
        """Adasum needs the UPDATE each rank would make, not its gradient (reference :738-893): every rank applies its own
        gradients locally, the per-variable delta against the last synchronised value is combined with op=Adasum, and the
        variables are set to that value plus the combined delta.  With backward_passes_per_step > 1 the local optimizer runs
        that many steps between two combinations."""

        def __init__(self, optimizer, name=None, use_locking=False, device_dense='', device_sparse='', compression=Compression.none,
                     backward_passes_per_step=1):
            super().__init__(name=name or 'DistributedDelta%s' % type(optimizer).__name__, use_locking=use_locking)
            self._optimizer, self._compression = optimizer, compression
            self._passes, self._step = int(backward_passes_per_step), 0
            self._start = {}

        def compute_gradients(self, *args, **kwargs):
            return self._optimizer.compute_gradients(*args, **kwargs)

        def apply_gradients(self, grads_and_vars, *args, **kwargs):
            pairs = [(g, v) for g, v in grads_and_vars if g is not None]
            key = (lambda v: v.ref()) if pairs and hasattr(pairs[0][1], 'ref') else id
            for _, v in pairs:
                if key(v) not in self._start:
                    self._start[key(v)] = tf.Variable(v.value() if hasattr(v, 'value') else v, trainable=False)
            out = self._optimizer.apply_gradients(pairs, *args, **kwargs)
            self._step += 1
            if self._step % self._passes:
                return out
            for i, (_, v) in enumerate(pairs):
                start = self._start[key(v)]
                delta = allreduce(v - start, op=_ops.Adasum, compression=self._compression, name='adasum_delta_%d' % i)
                start.assign(start + delta)
                v.assign(start)
            return out

        def get_slot(self, *args, **kwargs):
            return self._optimizer.get_slot(*args, **kwargs)

        def get_slot_names(self, *args, **kwargs):
            return self._optimizer.get_slot_names(*args, **kwargs)

        def variables(self, *args, **kwargs):
            return self._optimizer.variables(*args, **kwargs)


def DistributedOptimizer(optimizer, name=None, use_locking=False, device_dense='', device_sparse='', compression=Compression.none,
                         sparse_as_dense=False, backward_passes_per_step=1, op=_ops.Average, gradient_predivide_factor=1.0,
                         average_aggregated_gradients=False, num_groups=0, groups=None, process_set=_ops.global_process_set,
                         scale_local_gradients=True):
    """Wraps an optimizer so gradients are reduced across ranks before they are applied: a `tf.compat.v1.train.Optimizer` gets
    the `compute_gradients` wrapper (or the delta wrapper for op=Adasum), a Keras optimizer the Keras subclass wrapper."""
    if gradient_predivide_factor != 1.0 and op != _ops.Average:
        raise ValueError('gradient_predivide_factor not supported with op != Average')
    if op == _ops.Adasum and average_aggregated_gradients:
        raise ValueError('Adasum does not support average_aggregated_gradients == True')
    if num_groups != 0:
        warnings.warn('Parameter `num_groups` has been replaced by `groups`', DeprecationWarning)
        groups = groups if groups is not None else num_groups
    if groups is not None and not (isinstance(groups, list) or groups > 0):
        raise ValueError('groups should be a non-negative integer or a list of list of tf.Variable.')
    if _LegacyOptimizer is not None and isinstance(optimizer, _LegacyOptimizer):
        if op == _ops.Adasum:
            if process_set.process_set_id != 0:
                raise NotImplementedError('Adasum does not support process sets yet')
            return _DistributedAdasumOptimizer(optimizer, name, use_locking, device_dense, device_sparse, compression, backward_passes_per_step)
        return _DistributedOptimizer(optimizer, name, use_locking, device_dense, device_sparse, compression, sparse_as_dense, op,
                                     gradient_predivide_factor, backward_passes_per_step, average_aggregated_gradients, groups,
                                     process_set, scale_local_gradients)
    if op == _ops.Adasum:
        raise ValueError('op == Adasum is not supported yet with Keras')
    from horovod_b200._keras import create_distributed_optimizer
    return create_distributed_optimizer(tf.keras, optimizer, name, device_dense, device_sparse, compression, sparse_as_dense,
                                        gradient_predivide_factor, op, backward_passes_per_step, average_aggregated_gradients,
                                        groups, process_set, scale_local_gradients)


from horovod_b200.tensorflow.sync_batch_norm import SyncBatchNormalization  # noqa: E402,F401
from horovod_b200.tensorflow import elastic  # noqa: E402,F401
