"""Local gradient aggregation that also works inside `tf.function` graphs (reference
horovod/tensorflow/gradient_aggregation.py:23-310: counter and sums are graph variables, the reduce-or-skip decision is a
`tf.cond`).

Collectives are `tf.py_function` nodes in this runtime, so the graph variant only differs from the eager one in where the
window counter lives: a tf.Variable read by a `tf.cond`, instead of a Python integer that a trace would freeze.
"""
import tensorflow as tf

from horovod_b200.tensorflow.gradient_aggregation_eager import LocalGradientAggregationHelperEager
from horovod_b200.tensorflow.mpi_ops import _ops
from horovod_b200.tensorflow.util import _executing_eagerly


def apply_op_to_not_none_tensors(tensor_op, tensors, *args):
    """[tensor_op(t, *args) for the entries that are not None], None kept in place."""
    return [tensor_op(t, *args) if t is not None else t for t in tensors]


def get_not_none_from_list(tensor_list):
    return [x for x in tensor_list if x is not None]


class LocalGradientAggregationHelper(LocalGradientAggregationHelperEager):
    _OPTIMIZER_TYPE_KERAS = 'optimizer_type_keras'
    _OPTIMIZER_TYPE_LEGACY = 'optimizer_type_legacy'

    def __init__(self, backward_passes_per_step, allreduce_func, sparse_as_dense=False, average_aggregated_gradients=False,
                 rank=0, optimizer_type=_OPTIMIZER_TYPE_KERAS, process_set=_ops.global_process_set, scale_local_gradients=True):
        super().__init__(backward_passes_per_step, allreduce_func, sparse_as_dense, average_aggregated_gradients,
                         process_set, scale_local_gradients)
        self.rank = rank
        self.optimizer_type = optimizer_type
        self._graph_counter = None

    def compute_gradients(self, grads, vars):
        if _executing_eagerly() or self.backward_passes_per_step == 1:
            return super().compute_gradients(grads, vars)
        # ---- traced: every pass adds to the sums; the pass that fills the window reduces and clears them ----
        grads = [self._densify(g) for g in grads]
        if self._graph_counter is None:
            self._graph_counter = tf.Variable(0, trainable=False, dtype=tf.int32)
        for idx, g in enumerate(grads):
            if g is not None and idx not in self.locally_aggregated_grads:
                self.locally_aggregated_grads[idx] = tf.Variable(tf.zeros_like(g), trainable=False)
        live = [i for i, g in enumerate(grads) if g is not None]
        sums = [self.locally_aggregated_grads[i].assign_add(grads[i]) for i in live]
        count = self._graph_counter.assign_add(1)

        def reduce_and_clear():
            full = [None] * len(grads)
            for i, s in zip(live, sums):
                full[i] = s
            red = self._reduce(full, vars)
            if self.average_aggregated_gradients:
                red = apply_op_to_not_none_tensors(lambda t: t / self.backward_passes_per_step, red)
            with tf.control_dependencies(get_not_none_from_list(red)):
                clears = [self.locally_aggregated_grads[i].assign(tf.zeros_like(self.locally_aggregated_grads[i])) for i in live]
                clears.append(self._graph_counter.assign(0))
            with tf.control_dependencies(clears):
                return [tf.identity(red[i]) for i in live]

        out_live = tf.cond(tf.equal(count, self.backward_passes_per_step), reduce_and_clear, lambda: [tf.identity(s) for s in sums])
        out = [None] * len(grads)
        for i, o in zip(live, out_live):
            out[i] = o
        self.synced = None                        # decided at run time: see apply_gradients
        return out

    def apply_gradients(self, apply_grads_closure, optimizer, *args, **kwargs):
        if self.synced is not None:
            return super().apply_gradients(apply_grads_closure, optimizer, *args, **kwargs)
        # traced: the counter was reset to 0 by the pass that reduced

        def skip():
            it = getattr(optimizer, 'iterations', None)
            return it.assign_add(1) if it is not None else tf.no_op()
        return tf.cond(tf.equal(self._graph_counter, 0), apply_grads_closure, skip)
