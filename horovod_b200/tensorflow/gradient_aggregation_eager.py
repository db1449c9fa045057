"""Local gradient aggregation for eager execution: gradients are summed on this rank for `backward_passes_per_step`
calls and cross the wire once per window (reference horovod/tensorflow/gradient_aggregation_eager.py:12-210).

The running sums live in tf.Variables (GPU-resident for GPU gradients), so a window costs one `assign_add` per gradient
per pass and ONE fused allreduce launch per window instead of one per pass.
"""
import tensorflow as tf

from horovod_b200.tensorflow.mpi_ops import _ops


def _key(var):
    return var.ref() if hasattr(var, 'ref') else id(var)


class LocalGradientAggregationHelperEager:
    def __init__(self, backward_passes_per_step, allreduce_func, sparse_as_dense=False, average_aggregated_gradients=False,
                 process_set=_ops.global_process_set, scale_local_gradients=True):
        if int(backward_passes_per_step) <= 0:
            raise ValueError('backward_passes_per_step must be > 0')
        self.backward_passes_per_step = int(backward_passes_per_step)
        self.allreduce_grads = allreduce_func
        self.sparse_as_dense = sparse_as_dense
        self.average_aggregated_gradients = average_aggregated_gradients
        self.process_set = process_set
        self.scale_local_gradients = scale_local_gradients
        self.locally_aggregated_grads = {}       # gradient index -> tf.Variable holding the running sum of the window
        self.counter = 0                         # passes seen in the current window
        self.synced = False                      # did the last compute_gradients() close a window (= reduce)?
        self._local_vars = set()

    def register_local_var(self, var):
        """Gradients of `var` never leave this rank (they are divided by the set size when `scale_local_gradients`)."""
        self._local_vars.add(_key(var))

    # -- one backward pass ----------------------------------------------------------------------------------------------
    def _densify(self, grad):
        if isinstance(grad, tf.IndexedSlices):
            if not self.sparse_as_dense:
                raise ValueError('IndexedSlices are not supported when `backward_passes_per_step` > 1 and `sparse_as_dense` is False.')
            return tf.convert_to_tensor(grad)
        return grad

    def compute_gradients(self, grads, vars):
        """Returns the reduced gradients on the pass that closes a window (`self.synced` is True then) and the running
        local sums on the passes before it."""
        grads = list(grads)
        if self.backward_passes_per_step == 1:
            self.synced = True
            return self._reduce(grads, vars)
        sums = []
        for idx, grad in enumerate(grads):
            grad = self._densify(grad)
            if grad is None:
                sums.append(None)
                continue
            acc = self.locally_aggregated_grads.get(idx)
            if acc is None:
                acc = self.locally_aggregated_grads[idx] = tf.Variable(tf.zeros_like(grad), trainable=False)
            acc.assign_add(grad)
            sums.append(acc.read_value() if hasattr(acc, 'read_value') else acc.value())
        self.counter += 1
        self.synced = self.counter == self.backward_passes_per_step
        if not self.synced:
            return sums
        self.counter = 0
        reduced = self._reduce(sums, vars)
        for acc in self.locally_aggregated_grads.values():
            acc.assign(tf.zeros_like(acc))
        if self.average_aggregated_gradients:
            reduced = [g if g is None else g / self.backward_passes_per_step for g in reduced]
        return reduced

    def _reduce(self, grads, vars):
        vars = list(vars) if vars is not None else [None] * len(grads)
        shared = [i for i, v in enumerate(vars) if v is None or _key(v) not in self._local_vars]
        out = list(grads)
        red = self.allreduce_grads([grads[i] for i in shared], [vars[i] for i in shared])
        for i, r in zip(shared, red):
            out[i] = r
        if self.scale_local_gradients and len(shared) != len(out):
            n = float(self.process_set.size())
            keep = set(shared)
            out = [g if (i in keep or g is None) else g / n for i, g in enumerate(out)]
        return out

    # -- applying ---------------------------------------------------------------------------------------------------------
    def apply_gradients(self, apply_grads_closure, optimizer, *args, **kwargs):
        """Runs `apply_grads_closure()` after a pass that closed a window; on the other passes only the optimizer's
        iteration counter advances (what the reference does, so learning-rate schedules keyed on it see every pass)."""
        if self.synced:
            return apply_grads_closure()
        it = getattr(optimizer, 'iterations', None)
        if it is not None and hasattr(it, 'assign_add'):
            it.assign_add(1)
        return None
