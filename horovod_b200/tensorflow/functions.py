"""Variable / object helpers of the TensorFlow front end (reference horovod/tensorflow/functions.py: broadcast_variables
:66, broadcast_object :97, broadcast_object_fn :144, allgather_object :177)."""
import tensorflow as tf

from horovod_b200.tensorflow.mpi_ops import _eager, _normalize_name, _ns, _ops, broadcast



def broadcast_object(obj, root_rank=0, session=None, name=None, process_set=_ops.global_process_set):
    """root_rank's picklable `obj` on every rank.  `session` (TF1) is accepted and unused: the object path never builds
    graph nodes here, it goes through the runtime's byte-tensor broadcast directly."""
    return _ns['broadcast_object'](obj, root_rank=root_rank, name=name, process_set=process_set)


def allgather_object(obj, session=None, name=None, process_set=_ops.global_process_set):
    """List with every rank's picklable `obj`, in rank order."""
    return _ns['allgather_object'](obj, name=name, process_set=process_set)


def broadcast_variables(variables, root_rank, process_set=_ops.global_process_set, inplace=False):
    """Assigns root_rank's value to every variable on every rank (reference functions.py:66-95)."""
    variables = list(variables)
    for i, v in enumerate(variables):
        v.assign(broadcast(v, root_rank, name=f'bcast_var_{i}_{_normalize_name(getattr(v, "name", "") or str(i))}', process_set=process_set))
    return variables


def broadcast_global_variables(root_rank):
    """TF1-style helper; under TF2 eager there is no global collection, so the v1 collection is used if present."""
    if _eager():
        raise RuntimeError('hvd.broadcast_global_variables() does not support eager execution. Use hvd.broadcast_variables(<model/optimizer variables>) instead.')
    return broadcast_variables(tf.compat.v1.global_variables(), root_rank)


def broadcast_object_fn(root_rank=0, session=None, name=None, process_set=_ops.global_process_set):
    """Returns fn(obj) -> root_rank's obj (the reference builds a reusable graph for TF1 sessions; here the object path is
    eager in both modes, so this is a closure over `broadcast_object`)."""
    def _bcast(obj):
        return broadcast_object(obj, root_rank=root_rank, session=session, name=name, process_set=process_set)
    return _bcast
