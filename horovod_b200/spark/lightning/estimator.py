"""LightningEstimator / LightningModel (exported as TorchEstimator / TorchModel too, as the reference does).

Parity: horovod/spark/lightning/estimator.py (`TorchEstimator` :96-560: `model` is a LightningModule; plain nn.Module +
`optimizer` + `loss` are wrapped by `to_lightning_module`; params incl. `gradient_clip_val`, `callbacks`, `checkpoint_callback`
semantics via the Store) and spark/lightning/remote.py (per-rank function).  Training runs on
`spark/lightning/trainer.py:ModuleProtocolTrainer` — no dependency on the pytorch_lightning package.
"""
import numpy as np
import torch

from horovod_b200.spark.common.estimator import HorovodEstimator, HorovodModel
from horovod_b200.spark.common.params import P
from horovod_b200.spark.lightning.remote import RemoteTrainer, _train_fn, translate_trainer_args  # noqa: F401
from horovod_b200.spark.lightning.util import _deserialize, _serialize  # noqa: F401


def _is_protocol_module(m):
    return callable(getattr(m, 'training_step', None)) and callable(getattr(m, 'configure_optimizers', None))


class LightningEstimator(HorovodEstimator):
    """fit(df) -> LightningModel.  `model` follows the LightningModule protocol (its training_step receives a dict
    column name -> tensor); or pass a plain nn.Module together with `optimizer` and `loss`."""
    PARAMS = (
        P('gradient_clip_val', None, None, 'clip the global gradient norm after the allreduce'),
        P('num_gpus', None, None, 'accepted for compatibility: one GPU per process', camel='NumGPUs'),
        P('logger', None, None, 'object with log_metrics(dict, step=): gets the step loss every log_every_n_steps steps and every epoch record (rank 0)'),
        P('log_every_n_steps', 50, None, 'training steps between two logger calls'),
        P('loader_num_epochs', None, None, 'accepted for compatibility'),
        P('terminate_on_nan', False, None, 'raise as soon as a training loss is NaN or infinite (one device -> host read per step)'),
        P('profiler', None, None, 'accepted for compatibility'),
        P('checkpoint_callback', None, None, 'accepted for compatibility: rank 0 checkpoints into the store after every epoch'),
        P('trainer_args', None, None, 'pytorch_lightning.Trainer keyword arguments; max_epochs / gradient_clip_val / accumulate_grad_batches are honoured'),
        P('loss_constructors', None, None, 'callables that build the loss function(s) for a plain nn.Module (instead of `loss`)'),
        P('train_minibatch_fn', None, None, 'accepted for compatibility: the step is the module\'s training_step'),
        P('train_async_data_loader_queue_size', 64, None, 'batches the training reader thread may run ahead (with train_reader_num_workers >= 1)'),
        P('val_async_data_loader_queue_size', 64, None, 'same for the validation reader'),
        P('debug_data_loader', False, None, 'print per-batch timing of the async readers'),
    )

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._check_params()

    def _check_framework_params(self):
        model = self._get('model')
        if not _is_protocol_module(model):
            if not isinstance(model, torch.nn.Module):
                raise ValueError('model must follow the LightningModule protocol or be a torch.nn.Module')
            if self._get('optimizer') is None or (self._get('loss') is None and not self._get('loss_constructors')):
                raise ValueError('a plain torch.nn.Module needs `optimizer` and `loss` (or implement training_step / configure_optimizers)')

    def _module(self):
        model = self._get('model')
        if _is_protocol_module(model):
            return model
        from horovod_b200.spark.lightning.legacy import to_lightning_module
        loss = [make() for make in self._get('loss_constructors')] if self._get('loss_constructors') else self._get('loss')
        return to_lightning_module(model, self._get('optimizer'), loss, self._get('loss_weights'),
                                   self._get('feature_cols'), self._get('label_cols'), self._get('sample_weight_col'))

    def _fit_on_prepared_data(self, backend, dataset):
        g = self._get
        store, run_id = g('store'), self._new_run_id()
        module = self._module()
        columns = list(g('feature_cols')) + list(g('label_cols')) + ([g('sample_weight_col')] if g('sample_weight_col') else [])
        spec = dict(module=_serialize(module), columns=columns, store=store, train_path=dataset.train_path, val_path=dataset.val_path,
                    ckpt_path=store.get_checkpoint_path(run_id), resume=self._read_checkpoint(run_id), batch_size=g('batch_size'),
                    val_batch_size=g('val_batch_size') or g('batch_size'), epochs=g('epochs'), shuffle=g('shuffle'),
                    seed=g('random_seed') or 0, train_steps=g('train_steps_per_epoch'), val_steps=g('validation_steps_per_epoch'),
                    use_gpu=g('use_gpu'), verbose=g('verbose'), transformation_fn=g('transformation_fn'), row_shapes=self._row_shapes(),
                    inmemory_cache_all=g('inmemory_cache_all'), compression=g('gradient_compression'),
                    backward_passes_per_step=g('backward_passes_per_step'), gradient_clip_val=g('gradient_clip_val'),
                    callbacks=list(g('callbacks') or []), trainer_args=g('trainer_args'), data_module=g('data_module'),
                    train_reader_num_workers=g('train_reader_num_workers'), val_reader_num_workers=g('val_reader_num_workers'),
                    train_async_data_loader_queue_size=g('train_async_data_loader_queue_size'),
                    val_async_data_loader_queue_size=g('val_async_data_loader_queue_size'), debug_data_loader=g('debug_data_loader'),
                    transformation_removed_fields=g('transformation_removed_fields'), logger=g('logger'),
                    log_every_n_steps=g('log_every_n_steps'), terminate_on_nan=g('terminate_on_nan'))
        translate_trainer_args(g('trainer_args'))          # fail on the driver, not inside the job
        rank0 = backend.run(_train_fn, args=(spec,))[0]
        module.load_state_dict(rank0['state_dict'])
        return LightningModel(model=module, feature_columns=list(g('feature_cols')), label_columns=list(g('label_cols')),
                              history=rank0['history'], run_id=run_id, metadata=dataset.metadata, input_shapes=g('input_shapes'),
                              optimizer=g('optimizer'), loss=g('loss'), loss_constructors=g('loss_constructors'))


class LightningModel(HorovodModel):
    PARAMS = (
        P('input_shapes', None, None, 'one shape per feature column'),
        P('optimizer', None, None, 'the optimizer the model was trained with'),
        P('loss', None, None, 'the loss function(s) the model was trained with'),
        P('loss_constructors', None, None, 'callables that build the loss function(s)'),
    )

    def _predict(self, columns):
        module = self._get('model').cpu().eval()
        shapes = self._get('input_shapes') or [None] * len(columns)
        feats = []
        for (name, arr), shape in zip(columns.items(), shapes):
            t = torch.as_tensor(np.ascontiguousarray(arr))
            t = t.float() if t.dtype.is_floating_point else t
            feats.append(t.reshape([len(t)] + [d for d in shape if d != -1]) if shape else t)
        with torch.no_grad():
            out = module(*feats)
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        return [o.numpy() for o in outs]


# the reference exports its Lightning estimator under the torch names as well
TorchEstimator, TorchModel = LightningEstimator, LightningModel
