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
from horovod_b200.spark.torch.estimator import _deserialize, _serialize


def _is_protocol_module(m):
    return callable(getattr(m, 'training_step', None)) and callable(getattr(m, 'configure_optimizers', None))


def _train_fn(spec):
    import horovod_b200.torch as hvd
    from horovod_b200.data import DevicePrefetcher
    from horovod_b200.spark.data_loaders import ParquetShard, PytorchDataLoader, PytorchInmemDataLoader
    from horovod_b200.spark.lightning.trainer import ModuleProtocolTrainer
    hvd.init()
    dev = torch.device('cuda', hvd.local_rank()) if spec['use_gpu'] and torch.cuda.is_available() else torch.device('cpu')
    if dev.type == 'cuda':
        torch.cuda.set_device(dev)
    store = spec['store']
    module = _deserialize(spec['module'])
    first_epoch, opt_state = 0, None
    if spec['resume'] is not None and hvd.rank() == 0:
        ck = _deserialize(spec['resume'])
        module.load_state_dict(ck['model'])
        first_epoch, opt_state = ck['epoch'] + 1, ck['optimizer']
    first_epoch = hvd.broadcast_object(first_epoch, root_rank=0, name='pl.first_epoch')
    cols = spec['columns']
    loader_cls = PytorchInmemDataLoader if spec['inmemory_cache_all'] else PytorchDataLoader

    def loader(path, batch_size, shuffle, steps):
        shard = ParquetShard(store, path, cols, hvd.rank(), hvd.size(), spec['row_shapes'])
        return loader_cls(shard, batch_size=batch_size, shuffle=shuffle, seed=spec['seed'], steps=steps,
                          transformation_fn=spec['transformation_fn'], pin_memory=dev.type == 'cuda')

    def checkpoint(mod, opt, epoch):
        if spec['ckpt_path']:
            store.write(spec['ckpt_path'], _serialize({'model': mod.state_dict(), 'optimizer': opt.state_dict(), 'epoch': epoch}))
    trainer = ModuleProtocolTrainer(hvd, dev, epochs=spec['epochs'], first_epoch=first_epoch, compression=spec['compression'],
                                    backward_passes_per_step=spec['backward_passes_per_step'], gradient_clip_val=spec['gradient_clip_val'],
                                    callbacks=spec['callbacks'], checkpoint=checkpoint, verbose=spec['verbose'],
                                    prefetcher=lambda l: DevicePrefetcher(l, device=dev))
    trainer.setup(module, optimizer_state=opt_state)
    train = loader(spec['train_path'], spec['batch_size'], spec['shuffle'], spec['train_steps'])
    val = loader(spec['val_path'], spec['val_batch_size'], False, spec['val_steps']) if spec['val_path'] else None
    history = trainer.fit(module, train, val)
    state = {k: v.cpu() for k, v in module.state_dict().items()} if hvd.rank() == 0 else None
    hvd.barrier()
    hvd.shutdown()
    return {'history': history, 'state_dict': state}


class LightningEstimator(HorovodEstimator):
    """fit(df) -> LightningModel.  `model` follows the LightningModule protocol (its training_step receives a dict
    column name -> tensor); or pass a plain nn.Module together with `optimizer` and `loss`."""
    PARAMS = (
        P('gradient_clip_val', None, None, 'clip the global gradient norm after the allreduce'),
        P('num_gpus', None, None, 'accepted for compatibility: one GPU per process'),
        P('logger', None, None, 'accepted for compatibility: the history is returned with the model'),
        P('log_every_n_steps', 50, None, 'accepted for compatibility'),
        P('data_module', None, None, 'accepted for compatibility: shards are read by horovod_b200.spark.data_loaders'),
        P('loader_num_epochs', None, None, 'accepted for compatibility'),
        P('terminate_on_nan', False, None, 'accepted for compatibility'),
        P('profiler', None, None, 'accepted for compatibility'),
        P('checkpoint_callback', None, None, 'accepted for compatibility: rank 0 checkpoints into the store after every epoch'),
    )

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._check_params()

    def _check_framework_params(self):
        model = self._get('model')
        if not _is_protocol_module(model):
            if not isinstance(model, torch.nn.Module):
                raise ValueError('model must follow the LightningModule protocol or be a torch.nn.Module')
            if self._get('optimizer') is None or self._get('loss') is None:
                raise ValueError('a plain torch.nn.Module needs `optimizer` and `loss` (or implement training_step / configure_optimizers)')

    def _module(self):
        model = self._get('model')
        if _is_protocol_module(model):
            return model
        from horovod_b200.spark.lightning.trainer import to_lightning_module
        return to_lightning_module(model, self._get('optimizer'), self._get('loss'), self._get('loss_weights'),
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
                    callbacks=list(g('callbacks') or []))
        rank0 = backend.run(_train_fn, args=(spec,))[0]
        module.load_state_dict(rank0['state_dict'])
        return LightningModel(model=module, feature_columns=list(g('feature_cols')), label_columns=list(g('label_cols')),
                              history=rank0['history'], run_id=run_id, metadata=dataset.metadata, input_shapes=g('input_shapes'))


class LightningModel(HorovodModel):
    PARAMS = (
        P('input_shapes', None, None, 'one shape per feature column'),
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
