"""NDArray collectives (reference horovod/mxnet/mpi_ops.py: allreduce :75, allreduce_ :124, grouped_allreduce(_) :173/:232,
allgather :290, grouped_allgather :327, broadcast(_) :366/:409, alltoall :449, reducescatter :501, grouped_reducescatter :541).

There is no MXNet-specific native extension (the reference's mxnet/mpi_ops.cc + adapter.cc + tensor_util.cc + cuda_util.cc):
NDArrays reach the runtime through the framework bridge — DLPack for GPU arrays (`to_dlpack_for_write`: the kernels write
MXNet's own allocation, which makes the in-place ops zero-copy), numpy for host arrays.  MXNet's engine is asynchronous;
`wait_to_read()` before the hand-off replaces the reference's engine-callback (`MXEnginePushAsync`) ordering.  `priority`
is accepted for API parity and unused (the runtime orders by negotiation, not by an engine priority).
"""
try:
    import mxnet as mx
except ImportError as _e:  # pragma: no cover
    raise ImportError('horovod_b200.mxnet needs MXNet (not installed in this environment); the PyTorch front end is '
                      'horovod_b200.torch') from _e

import torch as _torch

from horovod_b200._bridge import BridgedOps as _BridgedOps, TensorBridge as _TensorBridge
from horovod_b200.torch import mpi_ops as _ops


class _MXBridge(_TensorBridge):
    name = 'mxnet'

    def to_torch(self, x):
        x.wait_to_read()
        if x.context.device_type == 'gpu':
            return _torch.utils.dlpack.from_dlpack(x.to_dlpack_for_write())
        return _torch.from_numpy(x.asnumpy())

    def from_torch(self, t, like=None):
        if t.is_cuda:
            return mx.nd.from_dlpack(_torch.utils.dlpack.to_dlpack(t.contiguous()))
        out = mx.nd.array(t.detach().cpu().numpy(), dtype=str(t.dtype).replace('torch.', ''))
        return out.as_in_context(like.context) if like is not None and hasattr(like, 'context') else out


_b = _BridgedOps(_MXBridge())
_b.export(globals())
_out_of_place_allreduce = allreduce  # noqa: F821 (exported above)

from horovod_b200.common.util import check_extension, split_list  # noqa: E402,F401


def _assign(dst, src):
    if src is not dst:
        dst[:] = src
    return dst


def allreduce(tensor, average=None, name=None, priority=0, prescale_factor=1.0, postscale_factor=1.0,  # noqa: F811
              process_set=global_process_set, op=None):  # noqa: F821
    return _out_of_place_allreduce(tensor, average, name, op, prescale_factor, postscale_factor, process_set)


def allreduce_(tensor, average=None, name=None, priority=0, prescale_factor=1.0, postscale_factor=1.0,
               process_set=global_process_set, op=None):  # noqa: F821
    """In place.  GPU arrays are reduced directly in MXNet's memory; host arrays go through one staging copy."""
    tensor.wait_to_read()
    if tensor.context.device_type == 'gpu':
        _ops.allreduce_(_torch.utils.dlpack.from_dlpack(tensor.to_dlpack_for_write()), average, name, op, prescale_factor,
                        postscale_factor, process_set)
        return tensor
    return _assign(tensor, allreduce(tensor, average, name, priority, prescale_factor, postscale_factor, process_set, op))


def grouped_allreduce_(tensors, average=None, name=None, priority=0, prescale_factor=1.0, postscale_factor=1.0,
                       process_set=global_process_set, op=None):  # noqa: F821
    outs = grouped_allreduce(tensors, average, name, op, prescale_factor, postscale_factor, process_set)  # noqa: F821
    for t, o in zip(tensors, outs):
        _assign(t, o)
    return tensors


def broadcast_(tensor, root_rank, name=None, priority=0, process_set=global_process_set):  # noqa: F821
    return _assign(tensor, broadcast(tensor, root_rank, name, process_set))  # noqa: F821
