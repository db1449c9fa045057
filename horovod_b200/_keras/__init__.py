"""Keras integration shared by `horovod_b200.keras` and `horovod_b200.tensorflow.keras`.

Parity: horovod/_keras/__init__.py (create_distributed_optimizer :30-256, allreduce/allgather/broadcast/reducescatter
value helpers :277-294, load_model :296-320).  The optimizer wrapper subclasses the user's optimizer class under the
SAME class name (so a saved model reloads without the wrapper), and reduces gradients at the three hook points Keras
has used over its versions: `_compute_gradients`/`get_gradients` (≤2.3), `_aggregate_gradients` (2.4–2.10) and
`apply_gradients` (Keras 3 / "experimental" optimizers, where aggregation hooks are gone).
"""
import tensorflow as tf

import horovod_b200.tensorflow as hvd


def create_distributed_optimizer(keras, optimizer, name, device_dense, device_sparse, compression, sparse_as_dense,
                                 gradient_predivide_factor, op, backward_passes_per_step=1,
                                 average_aggregated_gradients=False, groups=None, process_set=hvd.global_process_set,
                                 scale_local_gradients=True):
    base = optimizer.__class__
    reduce_fn = hvd._make_allreduce_grads_fn(name or 'Distributed' + base.__name__, device_dense, device_sparse, compression,
                                             sparse_as_dense, op, gradient_predivide_factor, groups, process_set)

    class _Distributed(base):
        _HAS_AGGREGATE_GRAD = True

        def __init__(self, **kwargs):
            super().__init__(**kwargs)
            self._hvd_reduced = False
            self._hvd_local = set()
            # local (unsynchronised) variables are filtered in _hvd_allreduce, the helper only keeps the window
            self._hvd_agg = hvd.LocalGradientAggregationHelperEager(backward_passes_per_step, self._hvd_allreduce,
                                                                    sparse_as_dense=True,
                                                                    average_aggregated_gradients=average_aggregated_gradients,
                                                                    process_set=process_set, scale_local_gradients=False)

        # -- local (non-synchronised) variables -----------------------------------------------------------------
        def register_local_var(self, var):
            self._hvd_local.add(var.ref() if hasattr(var, 'ref') else id(var))

        def _hvd_allreduce(self, grads, variables):
            key = (lambda v: v.ref()) if variables and hasattr(variables[0], 'ref') else id
            shared = [i for i, v in enumerate(variables) if key(v) not in self._hvd_local]
            out = list(grads)
            red = reduce_fn([grads[i] for i in shared], [variables[i] for i in shared])
            for i, r in zip(shared, red):
                out[i] = r
            if scale_local_gradients and len(shared) != len(out):
                n = process_set.size()
                out = [g if (i in shared or g is None) else g / n for i, g in enumerate(out)]
            return out

        def _hvd_reduce_pairs(self, grads_and_vars):
            pairs = list(grads_and_vars)
            grads, variables = [g for g, _ in pairs], [v for _, v in pairs]
            red = self._hvd_agg.compute_gradients(grads, variables)
            if not self._hvd_agg.synced:      # still inside an aggregation window: nothing to apply on this pass
                return None
            return list(zip(red, variables))

        # -- Keras <= 2.3 --------------------------------------------------------------------------------------
        def get_gradients(self, loss, params):
            grads = super().get_gradients(loss, params)
            self._hvd_reduced = True
            return self._hvd_allreduce(grads, params)

        # -- Keras 2.4 .. 2.10 -----------------------------------------------------------------------------------
        def _aggregate_gradients(self, grads_and_vars):
            pairs = self._hvd_reduce_pairs(grads_and_vars)
            self._hvd_reduced = True
            if pairs is None:  # accumulating: hand back zeros so that apply_gradients is a no-op step
                return [tf.zeros_like(g) if g is not None else None for g, _ in grads_and_vars]
            return [g for g, _ in pairs]

        # -- every version: last line of defence --------------------------------------------------------------
        def apply_gradients(self, grads_and_vars, *args, **kwargs):
            if self._hvd_reduced:
                self._hvd_reduced = False
                return super().apply_gradients(grads_and_vars, *args, **kwargs)
            pairs = self._hvd_reduce_pairs(grads_and_vars)
            if pairs is None:
                return None
            try:
                self._hvd_reduced = True  # _aggregate_gradients (if the base still calls it) must not reduce again
                return super().apply_gradients(pairs, *args, **kwargs)
            finally:
                self._hvd_reduced = False

    # the wrapper class carries the wrapped optimizer's name so that a saved model reloads without this package
    _Distributed.__name__ = base.__name__
    _Distributed.__qualname__ = base.__qualname__
    cfg = optimizer.get_config()
    try:
        return _Distributed.from_config(cfg)
    except Exception:  # noqa: BLE001 - optimizers whose from_config needs more than the config
        return _Distributed(**cfg)


def _value(x):
    return tf.convert_to_tensor(x) if not tf.is_tensor(x) else x


def allreduce(backend, value, name, average, prescale_factor, postscale_factor, op, compression):
    return hvd.allreduce(_value(value), average=average, name=name, op=op, prescale_factor=prescale_factor,
                         postscale_factor=postscale_factor, compression=compression).numpy()


def allgather(backend, value, name):
    return hvd.allgather(_value(value), name=name).numpy()


def broadcast(backend, value, root_rank, name):
    return hvd.broadcast(_value(value), root_rank, name=name).numpy()


def reducescatter(backend, value, name, op):
    return hvd.reducescatter(_value(value), name=name, op=op).numpy()


def broadcast_global_variables(backend, root_rank):
    return hvd.broadcast_global_variables(root_rank)


def load_model(keras, wrap_optimizer, filepath, custom_optimizers, custom_objects, legacy_opts=False):
    """keras.models.load_model with every known optimizer class swapped for its distributed wrapper."""
    opt_mod = keras.optimizers.legacy if legacy_opts and hasattr(keras.optimizers, 'legacy') else keras.optimizers
    classes = [c for c in vars(opt_mod).values() if isinstance(c, type) and issubclass(c, opt_mod.Optimizer) and c is not opt_mod.Optimizer]
    objs = {c.__name__: wrap_optimizer(c) for c in classes}
    for c in custom_optimizers or []:
        objs[c.__name__] = wrap_optimizer(c)
    objs.update(custom_objects or {})
    return keras.models.load_model(filepath, custom_objects=objs)
