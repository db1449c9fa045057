"""Object helpers of the MXNet front end (reference horovod/mxnet/functions.py: broadcast_object :27, allgather_object :64)."""
from horovod_b200.mxnet import mpi_ops as _mpi_ops


def broadcast_object(obj, root_rank=0, name=None):
    """root_rank's picklable `obj` on every rank (pickled into a byte tensor, size first)."""
    return _mpi_ops._b.broadcast_object(obj, root_rank=root_rank, name=name)


def allgather_object(obj, name=None):
    """List with every rank's picklable `obj`, in rank order."""
    return _mpi_ops._b.allgather_object(obj, name=name)
