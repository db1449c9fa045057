"""KerasEstimator / KerasModel.

Parity: horovod/spark/keras/estimator.py (`KerasEstimator` :105-395 — params model/optimizer/loss/loss_weights/metrics/
custom_objects/callbacks/checkpoint_callback; `_load_model_from_checkpoint`; `KerasModel.transform` :397-539) and
spark/keras/remote.py (per-rank function: deserialize, scale / wrap the optimizer, compile, BroadcastGlobalVariables +
MetricAverage callbacks, rank-0 checkpoint, `model.fit` on the rank's shard).

Same Store / Parquet / shard-reader core as the torch estimators (`spark/common`, `spark/data_loaders`); the model's
`fit` is fed by a Python generator of numpy batches, so no tf.data / Petastorm dependency.
"""
import numpy as np

from horovod_b200.spark.common.estimator import HorovodEstimator, HorovodModel
from horovod_b200.spark.common.params import P
from horovod_b200.spark.keras import util as kutil
from horovod_b200.spark.keras.datamodule import _NumpyShardBatches  # noqa: F401
from horovod_b200.spark.keras.remote import RemoteTrainer, _train_fn  # noqa: F401


class KerasEstimator(HorovodEstimator):
    """fit(df) -> KerasModel.  `model` is an (uncompiled or compiled) Keras model, `optimizer` a Keras optimizer instance,
    `loss` / `loss_weights` / `metrics` as for `model.compile`."""
    PARAMS = (
        P('custom_objects', None, None, 'custom layers / losses needed to rebuild the model on the workers'),
        P('checkpoint_callback', None, None, 'accepted for compatibility: rank 0 checkpoints into the store after every epoch'),
        P('backend_env', None, None, 'environment variables set on every training process before Keras / TensorFlow initialise'),
    )
    REQUIRED = ('model', 'optimizer', 'loss', 'feature_cols', 'label_cols', 'store')

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._check_params()

    def _check_framework_params(self):
        opt = self._get('optimizer')
        if not (hasattr(opt, 'get_config') and hasattr(type(opt), 'from_config')):
            raise ValueError('optimizer must be a Keras optimizer instance (get_config / from_config)')
        model = self._get('model')
        if not (hasattr(model, 'get_weights') and hasattr(model, 'fit')):
            raise ValueError('model must be a Keras model')

    def _fit_on_prepared_data(self, backend, dataset):
        g = self._get
        store, run_id = g('store'), self._new_run_id()
        columns = list(g('feature_cols')) + list(g('label_cols')) + ([g('sample_weight_col')] if g('sample_weight_col') else [])
        spec = dict(model=kutil.serialize_model(g('model')), optimizer=kutil.serialize_optimizer(g('optimizer')), loss=g('loss'),
                    loss_weights=g('loss_weights'), metrics=g('metrics'), custom_objects=g('custom_objects'),
                    callbacks=list(g('callbacks') or []), columns=columns, feature_cols=list(g('feature_cols')),
                    label_cols=list(g('label_cols')), sample_weight_col=g('sample_weight_col'), store=store,
                    train_path=dataset.train_path, val_path=dataset.val_path, ckpt_path=store.get_checkpoint_path(run_id),
                    resume=self._read_checkpoint(run_id), batch_size=g('batch_size'), val_batch_size=g('val_batch_size') or g('batch_size'),
                    epochs=g('epochs'), shuffle=g('shuffle'), seed=g('random_seed') or 0, train_steps=g('train_steps_per_epoch'),
                    val_steps=g('validation_steps_per_epoch'), verbose=g('verbose'), transformation_fn=g('transformation_fn'),
                    row_shapes=self._row_shapes(), compression=g('gradient_compression'),
                    backward_passes_per_step=g('backward_passes_per_step'), backend_env=g('backend_env'),
                    data_module=g('data_module'), transformation_removed_fields=g('transformation_removed_fields'))
        rank0 = backend.run(_train_fn, args=(spec,))[0]
        model = g('model')
        model.set_weights(kutil.weights_from_bytes(rank0['weights']))
        return KerasModel(model=model, feature_columns=list(g('feature_cols')), label_columns=list(g('label_cols')),
                          history=rank0['history'], run_id=run_id, metadata=dataset.metadata, custom_objects=g('custom_objects'),
                          input_shapes=g('input_shapes'))


class KerasModel(HorovodModel):
    PARAMS = (
        P('custom_objects', None, None, 'custom layers / losses'),
        P('input_shapes', None, None, 'one shape per feature column'),
    )

    def getHistory(self):
        return self._get('history') or {}

    def _predict(self, columns):
        shapes = self._get('input_shapes') or [None] * len(columns)
        feats = []
        for (name, arr), shape in zip(columns.items(), shapes):
            arr = np.asarray(arr, dtype=np.float32) if np.asarray(arr).dtype.kind == 'f' else np.asarray(arr)
            feats.append(arr.reshape([len(arr)] + [d for d in shape if d != -1]) if shape else arr)
        out = self._get('model').predict(feats[0] if len(feats) == 1 else feats)
        return list(out) if isinstance(out, (list, tuple)) else [np.asarray(out)]
