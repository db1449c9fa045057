"""Low-level TensorFlow collectives: the bridge into the native runtime, the scalar "ops" and the differentiable
allreduce / allgather / broadcast / alltoall / reducescatter (reference horovod/tensorflow/mpi_ops.py: basics :57-96,
_allreduce :115, allgather :206, broadcast :318, broadcast_ :359, alltoall :396, _reducescatter :457, join :564,
size_op ... local_rank_op :576-660).

There is NO TensorFlow custom-op library here (the reference's tensorflow/mpi_ops.cc + xla_mpi_ops.cc are ~2500 lines of
AsyncOpKernels).  TF tensors are handed to the runtime through the framework bridge (`horovod_b200._bridge`): DLPack for
GPU tensors (the P2P kernels read/write TF's own HBM allocation), the numpy view for host tensors.  Inside `tf.function`
graphs the collectives run as `tf.py_function` nodes; gradients are registered with `tf.custom_gradient`, with the same
backward collectives the reference registers (allreduce<->allreduce, allgather<->reducescatter-by-slice,
broadcast<->reduce-to-root, alltoall<->alltoall with the received splits).
"""
try:
    import tensorflow as tf
except ImportError as _e:  # pragma: no cover - exercised only where TF is absent
    raise ImportError('horovod_b200.tensorflow needs TensorFlow >= 2.4 (not installed in this environment); the PyTorch '
                      'front end is horovod_b200.torch') from _e

import warnings

import torch as _torch

from horovod_b200._bridge import BridgedOps as _BridgedOps, TensorBridge as _TensorBridge
from horovod_b200.common.exceptions import HorovodInternalError, HostsUpdatedInterrupt  # noqa: F401
from horovod_b200.torch import mpi_ops as _ops
from horovod_b200.common.util import check_extension, get_average_backwards_compatibility_fun, split_list  # noqa: F401
from horovod_b200.tensorflow.util import _executing_eagerly

_dlpack = getattr(getattr(tf, 'experimental', None), 'dlpack', None)


class _TFBridge(_TensorBridge):
    name = 'tensorflow'

    def to_torch(self, x):
        if isinstance(x, tf.Variable):
            x = x.value() if hasattr(x, 'value') else x
        x = tf.convert_to_tensor(x)
        dev = (getattr(x, 'device', '') or '').upper()
        if 'GPU' in dev and _dlpack is not None:
            return _torch.utils.dlpack.from_dlpack(_dlpack.to_dlpack(x))
        return _torch.from_numpy(x.numpy().copy())

    def from_torch(self, t, like=None):
        if t.is_cuda and _dlpack is not None:
            return _dlpack.from_dlpack(_torch.utils.dlpack.to_dlpack(t.contiguous()))
        return tf.convert_to_tensor(t.detach().cpu().numpy())


_b = _BridgedOps(_TFBridge())
_ns = {}
_b.export(_ns)
_BASICS = ('init', 'shutdown', 'is_initialized', 'start_timeline', 'stop_timeline', 'size', 'local_size', 'cross_size',
           'rank', 'local_rank', 'cross_rank', 'is_homogeneous', 'mpi_threads_supported', 'mpi_enabled', 'mpi_built',
           'gloo_enabled', 'gloo_built', 'nccl_built', 'ddl_built', 'ccl_built', 'cuda_built', 'rocm_built', 'p2p_built',
           'gpu_topology', 'gpu_backend_info', 'Average', 'Sum', 'Adasum', 'Min', 'Max', 'Product', 'global_process_set',
           'ProcessSet', 'add_process_set', 'remove_process_set', 'barrier', 'broadcast_object', 'allgather_object')
for _k in _BASICS:
    globals()[_k] = _ns[_k]
del _k


def join():
    """Blocks until every rank joined; returns the last rank that joined (reference tensorflow/mpi_ops.py:564)."""
    return _ops.join()


_eager = _executing_eagerly

handle_average_backwards_compatibility = get_average_backwards_compatibility_fun(_ops)


def _run(fn, inputs, out_dtypes, name):
    """Runs `fn(*eager tensors)` now (eager) or as a py_function node (inside tf.function)."""
    if _eager():
        return fn(*inputs)
    return tf.py_function(fn, inputs, out_dtypes, name=name)


def _normalize_name(name):
    import re
    return re.sub('[^a-zA-Z0-9_]', '_', name) if name else name


# ---- scalar "ops" (reference mpi_ops.py:576-660: graph nodes whose value is read at run time, for elastic jobs) ----
def size_op(process_set_id=0, name=None):
    return _run(lambda: tf.constant(_ops.size() if process_set_id == 0 else len(_process_set_ranks(process_set_id)), tf.int32), [], tf.int32, name)


def _process_set_ranks(ps_id):
    from horovod_b200.common.process_sets import _basics as psb
    return psb.process_set_ranks(ps_id)


def process_set_included_op(process_set_id=0, name=None):
    return _run(lambda: tf.constant(int(_ops.rank() in _process_set_ranks(process_set_id)) if process_set_id else 1, tf.int32), [], tf.int32, name)


def local_size_op(name=None):
    return _run(lambda: tf.constant(_ops.local_size(), tf.int32), [], tf.int32, name)


def rank_op(name=None):
    return _run(lambda: tf.constant(_ops.rank(), tf.int32), [], tf.int32, name)


def local_rank_op(name=None):
    return _run(lambda: tf.constant(_ops.local_rank(), tf.int32), [], tf.int32, name)


# ---- differentiable collectives ---------------------------------------------------------------------------------------
def _allreduce(tensor, name=None, op=_ops.Sum, prescale_factor=1.0, postscale_factor=1.0, process_set=_ops.global_process_set):
    name = _normalize_name(name)

    @tf.custom_gradient
    def f(x):
        y = _run(lambda t: _b.allreduce(t, name=name, op=op, prescale_factor=prescale_factor, postscale_factor=postscale_factor,
                                        process_set=process_set), [x], x.dtype, name)
        if not _eager():
            y.set_shape(x.shape)

        def grad(dy):
            return _allreduce(dy, name=(name + '_grad') if name else None, op=op, prescale_factor=prescale_factor,
                              postscale_factor=postscale_factor, process_set=process_set)
        return y, grad
    return f(tf.convert_to_tensor(tensor))


def allgather(tensor, name=None, ignore_name_scope=False, process_set=_ops.global_process_set):
    name = _normalize_name(name)

    @tf.custom_gradient
    def f(x):
        y = _run(lambda t: _b.allgather(t, name=name, process_set=process_set), [x], x.dtype, name)

        def grad(dy):
            # every rank's slice of the summed upstream gradient (reference _allgather_grad, mpi_ops.py:228-257)
            d0 = tf.shape(x, out_type=tf.int64)[:1]
            sizes = tf.reshape(allgather(d0, name=(name + '_sizes') if name else None, process_set=process_set), [-1])
            summed = _allreduce(dy, name=(name + '_grad') if name else None, op=_ops.Sum, process_set=process_set)
            r = process_set.rank()
            start = tf.reduce_sum(sizes[:r])
            return summed[start:start + sizes[r]]
        return y, grad
    return f(tf.convert_to_tensor(tensor))


def grouped_allgather(tensors, name=None, ignore_name_scope=False, process_set=_ops.global_process_set):
    return [allgather(t, name=f'{name}_{i}' if name else None, process_set=process_set) for i, t in enumerate(tensors)]


def broadcast(tensor, root_rank, name=None, ignore_name_scope=False, process_set=_ops.global_process_set):
    name = _normalize_name(name)

    @tf.custom_gradient
    def f(x):
        y = _run(lambda t: _b.broadcast(t, root_rank, name=name, process_set=process_set), [x], x.dtype, name)
        if not _eager():
            y.set_shape(x.shape)

        def grad(dy):
            g = _allreduce(dy, name=(name + '_grad') if name else None, op=_ops.Sum, process_set=process_set)
            return g if process_set.rank() == root_rank else tf.zeros_like(g)
        return y, grad
    return f(tf.convert_to_tensor(tensor))


def broadcast_(variables, root_rank, name=None, process_set=_ops.global_process_set):
    """In-place broadcast of tf.Variables (reference mpi_ops.py:359-394)."""
    for i, v in enumerate(variables):
        v.assign(broadcast(v, root_rank, name=f'{name or "bcast_"}_{i}', process_set=process_set))
    return variables


def alltoall(tensor, splits=None, name=None, ignore_name_scope=False, process_set=_ops.global_process_set):
    name = _normalize_name(name)
    x = tf.convert_to_tensor(tensor)
    if splits is None:
        n = process_set.size()
        splits = tf.fill([n], tf.shape(x)[0] // n)
    splits = tf.cast(tf.convert_to_tensor(splits), tf.int32)

    @tf.custom_gradient
    def f(x, s):
        y, rs = _run(lambda t, sp: _b.alltoall(t, splits=sp, name=name, process_set=process_set), [x, s], [x.dtype, tf.int32], name)

        def grad(dy, _drs):
            g, _ = alltoall(dy, splits=rs, name=(name + '_grad') if name else None, process_set=process_set)
            return g, None
        return (y, rs), grad
    return f(x, splits)


def _reducescatter(tensor, name=None, op=_ops.Sum, ignore_name_scope=False, process_set=_ops.global_process_set,
                   prescale_factor=1.0, postscale_factor=1.0):
    name = _normalize_name(name)

    @tf.custom_gradient
    def f(x):
        y = _run(lambda t: _b.reducescatter(t, name=name, op=op, process_set=process_set, prescale_factor=prescale_factor,
                                            postscale_factor=postscale_factor), [x], x.dtype, name)

        def grad(dy):
            g = allgather(dy, name=(name + '_grad') if name else None, process_set=process_set)
            scale = prescale_factor * postscale_factor / (process_set.size() if op == _ops.Average else 1)
            return g * tf.cast(scale, g.dtype) if scale != 1 else g
        return y, grad
    return f(tf.convert_to_tensor(tensor))
