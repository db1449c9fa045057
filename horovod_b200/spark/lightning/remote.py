"""The per-rank training function of the Lightning estimator (reference horovod/spark/lightning/remote.py `RemoteTrainer`
:37-330).  Training runs on `spark/lightning/trainer.py:ModuleProtocolTrainer` — no dependency on the pytorch_lightning package."""
import torch

from horovod_b200.spark.lightning.util import _deserialize, _serialize

# pytorch_lightning.Trainer argument -> ModuleProtocolTrainer argument; anything else in `trainer_args` is rejected by name
_TRAINER_ARGS = {'max_epochs': 'epochs', 'gradient_clip_val': 'gradient_clip_val', 'accumulate_grad_batches': 'backward_passes_per_step'}
# accepted and meaningless here (one process per GPU, no logger / progress bar objects)
_IGNORED_TRAINER_ARGS = ('gpus', 'devices', 'accelerator', 'strategy', 'logger', 'enable_progress_bar', 'progress_bar_refresh_rate',
                         'log_every_n_steps', 'num_sanity_val_steps', 'enable_checkpointing', 'checkpoint_callback',
                         'replace_sampler_ddp', 'use_distributed_sampler', 'profiler', 'terminate_on_nan', 'detect_anomaly')


def translate_trainer_args(trainer_args):
    """`trainer_args` of the estimator (keyword arguments a pytorch_lightning.Trainer would get) -> ModuleProtocolTrainer kwargs."""
    out = {}
    for k, v in (trainer_args or {}).items():
        if k in _TRAINER_ARGS:
            out[_TRAINER_ARGS[k]] = v
        elif k not in _IGNORED_TRAINER_ARGS:
            raise ValueError('trainer_args: %r is not supported by the protocol trainer (supported: %s)' % (k, ', '.join(sorted(_TRAINER_ARGS))))
    return out


def _train_fn(spec):
    import horovod_b200.torch as hvd
    from horovod_b200.data import DevicePrefetcher
    from horovod_b200.spark.common.util import make_transform
    from horovod_b200.spark.lightning.datamodule import ParquetDataModule
    from horovod_b200.spark.lightning.trainer import ModuleProtocolTrainer
    hvd.init()
    from horovod_b200.spark.common.util import gpu_index_for
    dev = torch.device('cuda', gpu_index_for(hvd.local_rank())) if spec['use_gpu'] and torch.cuda.is_available() else torch.device('cpu')
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
    module_cls = spec.get('data_module') or ParquetDataModule
    data_module = module_cls(train_dir=spec['train_path'], val_dir=spec['val_path'], num_train_epochs=spec['epochs'],
                             has_val=bool(spec['val_path']), train_batch_size=spec['batch_size'], val_batch_size=spec['val_batch_size'],
                             shuffle=spec['shuffle'],
                             transform_fn=make_transform(spec['transformation_fn'], spec.get('transformation_removed_fields')),
                             inmemory_cache_all=spec['inmemory_cache_all'], cur_shard=hvd.rank(), shard_count=hvd.size(),
                             schema_fields=cols, steps_per_epoch_train=spec['train_steps'], steps_per_epoch_val=spec['val_steps'],
                             verbose=spec['verbose'], store=store, row_shapes=spec['row_shapes'], seed=spec['seed'],
                             pin_memory=dev.type == 'cuda', train_reader_num_workers=spec.get('train_reader_num_workers'),
                             val_reader_num_workers=spec.get('val_reader_num_workers'),
                             train_async_data_loader_queue_size=spec.get('train_async_data_loader_queue_size') or 64,
                             val_async_data_loader_queue_size=spec.get('val_async_data_loader_queue_size') or 64,
                             debug_data_loader=bool(spec.get('debug_data_loader')))
    data_module.__enter__()

    def checkpoint(mod, opt, epoch):
        if spec['ckpt_path']:
            store.write(spec['ckpt_path'], _serialize({'model': mod.state_dict(), 'optimizer': opt.state_dict(), 'epoch': epoch}))
    trainer_kwargs = dict(epochs=spec['epochs'], first_epoch=first_epoch, compression=spec['compression'],
                          backward_passes_per_step=spec['backward_passes_per_step'], gradient_clip_val=spec['gradient_clip_val'],
                          callbacks=spec['callbacks'], checkpoint=checkpoint, verbose=spec['verbose'],
                          prefetcher=lambda l: DevicePrefetcher(l, device=dev), logger=spec.get('logger'),
                          log_every_n_steps=spec.get('log_every_n_steps') or 50, terminate_on_nan=bool(spec.get('terminate_on_nan')))
    trainer_kwargs.update(translate_trainer_args(spec.get('trainer_args')))
    trainer = ModuleProtocolTrainer(hvd, dev, **trainer_kwargs)
    trainer.setup(module, optimizer_state=opt_state)
    train = data_module.train_data()
    val = data_module.val_data() if spec['val_path'] else None
    history = trainer.fit(module, train, val)
    data_module.__exit__(None, None, None)
    state = {k: v.cpu() for k, v in module.state_dict().items()} if hvd.rank() == 0 else None
    hvd.barrier()
    hvd.shutdown()
    return {'history': history, 'state_dict': state}


def RemoteTrainer(spec):
    """-> fn() for the backend to run on every rank."""
    def train():
        return _train_fn(spec)
    return train
