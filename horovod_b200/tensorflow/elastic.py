"""Elastic state for TensorFlow / Keras (parity: horovod/tensorflow/elastic.py:31-230)."""
import tensorflow as tf

from horovod_b200.common.elastic import ObjectState, run_fn
from horovod_b200.torch import mpi_ops as _ops


def _reset():
    _ops.shutdown()
    _ops.init()


def run(func):
    """Decorator: retries `func(state, ...)` after HorovodInternalError / HostsUpdatedInterrupt (TF2 eager only)."""
    return run_fn(func, _reset)


def _bcast_object(obj, root_rank=0, name=None):
    from horovod_b200.torch.functions import broadcast_object
    return broadcast_object(obj, root_rank, name)


class TensorFlowKerasState(ObjectState):
    """Tracks a Keras model (+ optimizer) and arbitrary picklable attributes."""

    def __init__(self, model, optimizer=None, backend=None, **kwargs):
        self.model = model
        self.optimizer = optimizer if optimizer is not None else getattr(model, 'optimizer', None)
        self._saved_model = None
        self._saved_opt = None
        super().__init__(bcast_object=_bcast_object, get_rank=_ops.rank, **kwargs)

    def _opt_vars(self):
        if self.optimizer is None:
            return []
        v = self.optimizer.variables
        return list(v() if callable(v) else v)

    def save(self):
        self._saved_model = [w.copy() for w in self.model.get_weights()]
        self._saved_opt = [v.numpy().copy() for v in self._opt_vars()]
        super().save()

    def restore(self):
        if self._saved_model is not None:
            self.model.set_weights(self._saved_model)
            for v, s in zip(self._opt_vars(), self._saved_opt or []):
                v.assign(s)
        super().restore()

    def sync(self):
        import horovod_b200.tensorflow as hvd
        hvd.broadcast_variables(list(self.model.variables) + self._opt_vars(), root_rank=0)
        self.save()
        super().sync()


class TensorFlowState(ObjectState):
    """Tracks an explicit list of tf.Variables (TF2 eager)."""

    def __init__(self, variables=None, session=None, **kwargs):
        if session is not None:
            raise ValueError('TF1 sessions are not supported by this front end; run eagerly')
        self.variables = list(variables or [])
        self._values = None
        super().__init__(bcast_object=_bcast_object, get_rank=_ops.rank, **kwargs)

    def save(self):
        self._values = [v.numpy().copy() for v in self.variables]
        super().save()

    def restore(self):
        for v, s in zip(self.variables, self._values or []):
            v.assign(s)
        super().restore()

    def sync(self):
        import horovod_b200.tensorflow as hvd
        hvd.broadcast_variables(self.variables, root_rank=0)
        self.save()
        super().sync()
