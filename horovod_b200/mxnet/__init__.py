"""`import horovod_b200.mxnet as hvd` — MXNet front end.

Parity: horovod/mxnet/__init__.py (DistributedOptimizer :44-120, DistributedTrainer :124-235, broadcast_parameters
:245-310), mpi_ops.py (allreduce(_)/grouped/allgather/broadcast(_)/alltoall/reducescatter) and functions.py.

As for TensorFlow there is no MXNet-specific native extension (the reference's mxnet/mpi_ops.cc + adapter.cc +
tensor_util.cc + cuda_util.cc): NDArrays reach the runtime through the framework bridge — DLPack for GPU arrays
(`to_dlpack_for_write`: the kernels write MXNet's own allocation, which makes the in-place ops zero-copy), numpy for host
arrays.  MXNet's engine is asynchronous; `wait_to_read()` before the hand-off replaces the reference's engine-callback
(`MXEnginePushAsync`) ordering.

MXNet is not part of this image: importing this module without it raises ImportError.
"""
try:
    import mxnet as mx
except ImportError as _e:  # pragma: no cover
    raise ImportError('horovod_b200.mxnet needs MXNet (not installed in this environment); the PyTorch front end is '
                      'horovod_b200.torch') from _e

import types
import warnings
from collections import OrderedDict, defaultdict  # noqa: F401

from horovod_b200.mxnet import mpi_ops as _mpi_ops
from horovod_b200.mxnet.compression import Compression  # noqa: F401
from horovod_b200.mxnet.mpi_ops import *  # noqa: F401,F403
from horovod_b200.mxnet.mpi_ops import _assign, _b, _ops, _torch  # noqa: F401
from horovod_b200.mxnet.functions import allgather_object, broadcast_object  # noqa: F401
from horovod_b200.common.util import check_extension, split_list  # noqa: F401


def _split_list(items, k):
    k = max(1, min(k, len(items)))
    per, extra = divmod(len(items), k)
    out, s = [], 0
    for i in range(k):
        e = s + per + (1 if i < extra else 0)
        out.append(items[s:e])
        s = e
    return out


class DistributedOptimizer(mx.optimizer.Optimizer):
    """Wraps an mx.optimizer.Optimizer: gradients are summed over ranks inside `update`; the averaging is folded into
    the wrapped optimizer's rescale_grad (so no extra pass over the gradient), as the reference does."""

    def __init__(self, optimizer, gradient_predivide_factor=1.0, num_groups=0, process_set=global_process_set):  # noqa: F821
        if gradient_predivide_factor != 1.0 and rocm_built():  # noqa: F821
            raise ValueError('gradient_predivide_factor not supported yet with ROCm')
        self._optimizer = optimizer
        self._optimizer.rescale_grad *= gradient_predivide_factor / process_set.size()
        self._gradient_predivide_factor = gradient_predivide_factor
        self._num_groups = num_groups
        self._process_set = process_set

    def __getattr__(self, item):
        return getattr(self._optimizer, item)

    def create_state(self, index, weight):
        return self._optimizer.create_state(index, weight)

    def create_state_multi_precision(self, index, weight):
        return self._optimizer.create_state_multi_precision(index, weight)

    def _do_allreduce(self, index, grad):
        if self._process_set.size() == 1:
            return
        pre = 1.0 / self._gradient_predivide_factor
        if isinstance(index, (tuple, list)):
            if self._num_groups > 0:
                pairs = list(zip(index, grad))
                for gi, chunk in enumerate(_split_list(pairs, self._num_groups)):
                    grouped_allreduce_([g for _, g in chunk], average=False, name=f'{chunk[0][0]}:{chunk[-1][0]}', priority=-gi,
                                       prescale_factor=pre, process_set=self._process_set)
            else:
                for i, g in zip(index, grad):
                    allreduce_(g, average=False, name=str(i), priority=-i, prescale_factor=pre, process_set=self._process_set)
        else:
            allreduce_(grad, average=False, name=str(index), prescale_factor=pre, process_set=self._process_set)

    def update(self, index, weight, grad, state):
        self._do_allreduce(index, grad)
        self._optimizer.update(index, weight, grad, state)

    def update_multi_precision(self, index, weight, grad, state):
        self._do_allreduce(index, grad)
        self._optimizer.update_multi_precision(index, weight, grad, state)

    def set_learning_rate(self, lr):
        self._optimizer.set_learning_rate(lr)

    def set_lr_mult(self, args_lr_mult):
        self._optimizer.set_lr_mult(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        self._optimizer.set_wd_mult(args_wd_mult)


class DistributedTrainer(mx.gluon.Trainer):
    """gluon.Trainer whose `_allreduce_grads` uses hvd instead of a kvstore; `_scale` carries the 1/size averaging."""

    def __init__(self, params, optimizer, optimizer_params=None, compression=Compression.none, gradient_predivide_factor=1.0,
                 prefix=None, num_groups=0, process_set=global_process_set):  # noqa: F821
        self._compression = compression
        self._process_set = process_set
        if gradient_predivide_factor != 1.0 and rocm_built():  # noqa: F821
            raise ValueError('gradient_predivide_factor not supported yet with ROCm')
        if isinstance(optimizer, DistributedOptimizer):
            optimizer = optimizer._optimizer
            warnings.warn('DistributedTrainer does not take DistributedOptimizer as its optimizer. We have unwrapped it for you.')
        if isinstance(params, dict):
            params = OrderedDict(sorted(params.items()))  # identical reduction order on every rank
        elif isinstance(params, (list, tuple)):
            params = sorted(params, key=lambda p: p.name)
        super().__init__(params, optimizer, optimizer_params=optimizer_params, kvstore=None)
        self._scale *= gradient_predivide_factor / process_set.size()
        self._gradient_predivide_factor = gradient_predivide_factor
        self._prefix = prefix or ''
        self._num_groups = num_groups

    def _allreduce_grads(self):
        if self._process_set.size() == 1:
            return
        pre = 1.0 / self._gradient_predivide_factor
        live = [(i, p) for i, p in enumerate(self._params) if p.grad_req != 'null']
        if self._num_groups > 0:
            for gi, chunk in enumerate(_split_list(live, self._num_groups)):
                grads, ctxs = [], []
                for _, p in chunk:
                    c, ctx = self._compression.compress(p.list_grad()[0])
                    grads.append(c)
                    ctxs.append(ctx)
                grouped_allreduce_(grads, average=False, name=f'{self._prefix}{chunk[0][0]}:{chunk[-1][0]}', priority=-gi,
                                   prescale_factor=pre, process_set=self._process_set)
                for (_, p), g, ctx in zip(chunk, grads, ctxs):
                    _assign(p.list_grad()[0], self._compression.decompress(g, ctx))
        else:
            for i, p in live:
                c, ctx = self._compression.compress(p.list_grad()[0])
                allreduce_(c, average=False, name=self._prefix + str(i), priority=-i, prescale_factor=pre, process_set=self._process_set)
                _assign(p.list_grad()[0], self._compression.decompress(c, ctx))


def _broadcast_after_init(param, root_rank, name):
    """Deferred-initialisation parameters are broadcast right after they materialise."""
    init_impl = getattr(param, '_init_impl')

    def wrapped(self, *args, **kwargs):
        init_impl(*args, **kwargs)
        broadcast_(self.data(), root_rank=root_rank, name=name)
    param._init_impl = types.MethodType(wrapped, param)


def broadcast_parameters(params, root_rank=0, prefix=None):
    """Broadcasts a dict of NDArrays or a gluon ParameterDict from root_rank (sorted by name on every rank)."""
    if size() == 1:  # noqa: F821
        return
    prefix = prefix or ''
    tensors, names = [], []
    try:
        from mxnet.gluon.parameter import ParameterDict
        valid = (dict, ParameterDict)
    except ImportError:
        valid = (dict,)
    if not isinstance(params, valid):
        raise ValueError('invalid params of type: %s' % type(params))
    for name, p in sorted(params.items()):
        if isinstance(p, mx.gluon.parameter.Parameter):
            try:
                tensors.append(p.data())
                names.append(prefix + str(name))
            except mx.gluon.parameter.DeferredInitializationError:
                _broadcast_after_init(p, root_rank, prefix + str(name))
        else:
            tensors.append(p)
            names.append(prefix + str(name))
    for t, nme in zip(tensors, names):
        broadcast_(t, root_rank, name=nme)
    for t in tensors:
        t.wait_to_read()
