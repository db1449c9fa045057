"""`import horovod_b200.numpy as hvd` — the collectives on numpy arrays (host memory, CPU data plane).

Not in the reference (its framework-neutral surface is the C API only); here it is the smallest `TensorBridge` and the
template the TensorFlow / MXNet front ends follow.  Arrays are viewed, not copied, on the way in; results are fresh
arrays.
"""
import numpy as _np
import torch as _torch

from horovod_b200._bridge import BridgedOps as _BridgedOps, TensorBridge as _TensorBridge


class _NumpyBridge(_TensorBridge):
    name = 'numpy'

    def to_torch(self, x):
        a = _np.ascontiguousarray(x)
        if not a.flags.writeable:
            a = a.copy()
        return _torch.from_numpy(a)

    def from_torch(self, t, like=None):
        return t.detach().cpu().numpy()


_bridge_ops = _BridgedOps(_NumpyBridge())
_bridge_ops.export(globals())


def broadcast_(array, root_rank, name=None, process_set=global_process_set):  # noqa: F821
    """In-place broadcast into a writeable, C-contiguous array."""
    if not (isinstance(array, _np.ndarray) and array.flags.c_contiguous and array.flags.writeable):
        raise ValueError('broadcast_ needs a writeable C-contiguous ndarray')
    from horovod_b200.torch import mpi_ops as _ops
    _ops.broadcast_(_torch.from_numpy(array), root_rank, name, process_set)
    return array


def allreduce_(array, average=None, name=None, op=None, prescale_factor=1.0, postscale_factor=1.0,
               process_set=global_process_set):  # noqa: F821
    if not (isinstance(array, _np.ndarray) and array.flags.c_contiguous and array.flags.writeable):
        raise ValueError('allreduce_ needs a writeable C-contiguous ndarray')
    from horovod_b200.torch import mpi_ops as _ops
    _ops.allreduce_(_torch.from_numpy(array), average, name, op, prescale_factor, postscale_factor, process_set)
    return array
