"""Cross-rank batch normalisation layer (parity: horovod/tensorflow/sync_batch_norm.py:22-70): the batch mean and the
mean of squares are averaged over ranks in ONE grouped allreduce; variance is derived from them."""
import tensorflow as tf


class SyncBatchNormalization(tf.keras.layers.BatchNormalization):
    def __init__(self, fused=False, **kwargs):
        if fused in (True, None):
            raise ValueError('SyncBatchNormalization does not support fused=True.')
        if not kwargs.get('name'):
            kwargs['name'] = 'sync_batch_normalization'
        try:
            super().__init__(fused=fused, **kwargs)
        except TypeError:  # Keras 3 dropped the `fused` argument
            super().__init__(**kwargs)

    def _moments(self, inputs, reduction_axes, keep_dims=False, *a, **k):
        import horovod_b200.tensorflow as hvd
        mean = tf.reduce_mean(inputs, axis=reduction_axes, keepdims=keep_dims)
        sq = tf.reduce_mean(tf.square(inputs), axis=reduction_axes, keepdims=keep_dims)
        if hvd.size() > 1:
            mean, sq = hvd.grouped_allreduce([mean, sq], op=hvd.Average, name=self.name + '_moments')
        return mean, sq - tf.square(mean)

    _calculate_mean_and_var = _moments  # TF <= 2.15 hook name
