"""TorchEstimator / TorchModel: fit a torch model on a DataFrame with data-parallel training, get back a transformer.

Parity: horovod/spark/torch/estimator.py (`TorchEstimator` :94-353 — params model/optimizer/loss/loss_weights/metrics/
feature_cols/label_cols/input_shapes/batch_size/epochs/validation/sample_weight_col/store/backend/num_proc/shuffle/
train_steps_per_epoch/transformation_fn/verbose; `TorchModel.transform` :355-500) and spark/torch/remote.py (the per-rank
training function: shard the Parquet files by rank, wrap the optimizer in DistributedOptimizer, broadcast the initial
state, average the epoch metrics, checkpoint on rank 0, resume from the run's checkpoint).

The reference materialises the DataFrame to Parquet in the Store and reads it back through Petastorm.  Here the
intermediate format is the same (Parquet in the Store) but the reader is pyarrow.dataset
(`horovod_b200.spark.data_loaders`) feeding a pinned side-stream `DevicePrefetcher`, and the input may be a Spark
DataFrame (written by Spark itself) or a pandas DataFrame (written by pyarrow) — so the estimator also works on a single
multi-GPU box without Spark (`LocalBackend`).
"""
import numpy as np
import torch

from horovod_b200.spark.common.estimator import HorovodEstimator, HorovodModel
from horovod_b200.spark.common.params import P


from horovod_b200.spark.torch.remote import RemoteTrainer, _BatchLoss, _as_list, _train_fn  # noqa: F401
from horovod_b200.spark.torch.util import _deserialize, _serialize  # noqa: F401


class TorchEstimator(HorovodEstimator):
    """fit(df) -> TorchModel.

    `model` (nn.Module), `optimizer` (a torch optimizer INSTANCE built on the model — its class and defaults are re-created on
    every rank), `loss` (callable(output, label) or one per label column; reduction='none' when sample_weight_col is used),
    plus every knob of `EstimatorParams` (also reachable as setX/getX and through `fit(df, params={...})`).
    """
    PARAMS = (
        P('train_minibatch_fn', None, None, 'accepted for compatibility; the loop lives in spark/torch/remote.py:_train_fn'),
        P('loss_constructors', None, None, 'callables that build the loss function(s) on the training processes (instead of `loss`)'),
    )
    REQUIRED = ('model', 'optimizer', 'feature_cols', 'label_cols', 'store')

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._check_params()

    def _check_framework_params(self):
        if not isinstance(self._get('optimizer'), torch.optim.Optimizer):
            raise ValueError('optimizer must be a torch.optim.Optimizer instance')
        n = len(self._get('label_cols'))
        if not self._get('loss') and not self._get('loss_constructors'):
            raise ValueError('TorchEstimator: required parameter(s) missing: loss (or loss_constructors)')
        if self._get('loss_constructors'):
            if not all(callable(c) for c in _as_list(self._get('loss_constructors'), n)):
                raise ValueError('loss_constructors must be callables that return a loss function')
        _as_list(self._get('loss'), n)
        _as_list(self._get('loss_weights'), n)

    def _fit_on_prepared_data(self, backend, dataset):
        g = self._get
        store, run_id = g('store'), self._new_run_id()
        optimizer = g('optimizer')
        skip = ('differentiable', 'foreach', 'fused', 'capturable', 'maximize')
        spec = dict(
            model=_serialize(g('model')), optimizer_cls=type(optimizer),
            optimizer_defaults={k: v for k, v in optimizer.defaults.items() if k not in skip or v},
            loss=g('loss'), loss_weights=g('loss_weights'), metrics=g('metrics'), callbacks=list(g('callbacks') or []),
            feature_cols=list(g('feature_cols')), label_cols=list(g('label_cols')), sample_weight_col=g('sample_weight_col'),
            batch_size=g('batch_size'), val_batch_size=g('val_batch_size') or g('batch_size'), epochs=g('epochs'),
            store=store, train_path=dataset.train_path, val_path=dataset.val_path, ckpt_path=store.get_checkpoint_path(run_id),
            resume=self._read_checkpoint(run_id), shuffle=g('shuffle'), seed=g('random_seed') or 0,
            train_steps=g('train_steps_per_epoch'), val_steps=g('validation_steps_per_epoch'), use_gpu=g('use_gpu'),
            verbose=g('verbose'), transformation_fn=g('transformation_fn'), row_shapes=self._row_shapes(),
            inmemory_cache_all=g('inmemory_cache_all'), compression=g('gradient_compression'),
            backward_passes_per_step=g('backward_passes_per_step'), data_module=g('data_module'),
            loss_constructors=g('loss_constructors'), train_reader_num_workers=g('train_reader_num_workers'),
            val_reader_num_workers=g('val_reader_num_workers'), transformation_removed_fields=g('transformation_removed_fields'),
            categorical_cols=g('categorical_cols'), continuous_cols=g('continuous_cols'))
        results = backend.run(_train_fn, args=(spec,))
        rank0 = results[0]
        model = g('model')
        model.load_state_dict(rank0['state_dict'])
        return TorchModel(model=model, feature_columns=spec['feature_cols'], label_columns=spec['label_cols'], history=rank0['history'],
                          run_id=run_id, metadata=dataset.metadata, input_shapes=g('input_shapes'), optimizer=optimizer,
                          loss=g('loss'), loss_constructors=g('loss_constructors'))


class TorchModel(HorovodModel):
    """Transformer returned by fit(): appends `<label>__output` prediction columns."""
    PARAMS = (
        P('input_shapes', None, None, 'one shape per feature column'),
        P('optimizer', None, None, 'the optimizer the model was trained with'),
        P('loss', None, None, 'the loss function(s) the model was trained with'),
        P('loss_constructors', None, None, 'callables that build the loss function(s)'),
    )

    def __init__(self, model=None, feature_cols=None, label_cols=None, **kwargs):
        if feature_cols is not None:
            kwargs.setdefault('feature_columns', list(feature_cols))
        if label_cols is not None:
            kwargs.setdefault('label_columns', list(label_cols))
        super().__init__(model=model, **kwargs)

    def _predict(self, columns):
        model = self._get('model').cpu().eval()
        shapes = self._get('input_shapes') or [None] * len(columns)
        feats = []
        for (name, arr), shape in zip(columns.items(), shapes):
            t = torch.as_tensor(np.ascontiguousarray(arr))
            t = t.float() if t.dtype.is_floating_point else t
            feats.append(t.reshape([len(t)] + [d for d in shape if d != -1]) if shape else t)
        with torch.no_grad():
            out = model(*feats)
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        return [o.numpy() for o in outs]
