"""The per-rank training function of the Keras estimator (reference horovod/spark/keras/remote.py `RemoteTrainer` :33-330:
deserialize, wrap the optimizer, compile, BroadcastGlobalVariables + MetricAverage callbacks, rank-0 checkpoint, `model.fit`
on the rank's shard)."""
from horovod_b200.spark.keras import util as kutil


def _train_fn(spec):
    import os
    os.environ.update({k: str(v) for k, v in (spec.get('backend_env') or {}).items()})   # before Keras / TF initialise
    import horovod_b200.tensorflow.keras as hvd
    from horovod_b200.spark.common.util import make_transform
    from horovod_b200.spark.keras.datamodule import ParquetDataModule
    hvd.init()
    keras = kutil.keras_module()
    store = spec['store']
    model = kutil.deserialize_model(spec['model'], spec['custom_objects'])
    first_epoch = 0
    if spec['resume'] is not None:
        ck = __import__('cloudpickle').loads(spec['resume'])
        model.set_weights(kutil.weights_from_bytes(ck['weights']))
        first_epoch = ck['epoch'] + 1
    optimizer = hvd.DistributedOptimizer(kutil.deserialize_optimizer(spec['optimizer']),
                                         compression=spec['compression'] or hvd.Compression.none,
                                         backward_passes_per_step=spec['backward_passes_per_step'])
    model.compile(optimizer=optimizer, loss=spec['loss'], loss_weights=spec['loss_weights'], metrics=spec['metrics'])

    cols = spec['columns']

    module_cls = spec.get('data_module') or ParquetDataModule
    data_module = module_cls(train_dir=spec['train_path'], val_dir=spec['val_path'], num_train_epochs=spec['epochs'],
                             has_val=bool(spec['val_path']), train_batch_size=spec['batch_size'], val_batch_size=spec['val_batch_size'],
                             shuffle=spec['shuffle'],
                             transform_fn=make_transform(spec['transformation_fn'], spec.get('transformation_removed_fields')),
                             cur_shard=hvd.rank(), shard_count=hvd.size(), schema_fields=cols,
                             steps_per_epoch_train=spec['train_steps'], steps_per_epoch_val=spec['val_steps'],
                             verbose=spec['verbose'], store=store, row_shapes=spec['row_shapes'], seed=spec['seed'])
    data_module.__enter__()
    train = data_module.train_data()
    val = data_module.val_data() if spec['val_path'] else None

    class _StoreCheckpoint(keras.callbacks.Callback):
        def on_epoch_end(self, epoch, logs=None):
            import cloudpickle
            store.write(spec['ckpt_path'], cloudpickle.dumps({'weights': kutil.weights_to_bytes(self.model.get_weights()), 'epoch': epoch}))

    callbacks = [hvd.callbacks.BroadcastGlobalVariablesCallback(0), hvd.callbacks.MetricAverageCallback()]
    callbacks += list(spec['callbacks'])
    if hvd.rank() == 0 and spec['ckpt_path']:
        callbacks.append(_StoreCheckpoint())
    fit_kwargs = dict(steps_per_epoch=train.steps, epochs=spec['epochs'], initial_epoch=first_epoch, callbacks=callbacks,
                      verbose=spec['verbose'] if hvd.rank() == 0 else 0)
    if val is not None:
        fit_kwargs.update(validation_data=kutil.batch_generator(val, spec['feature_cols'], spec['label_cols'], spec['sample_weight_col']),
                          validation_steps=val.steps)
    history = model.fit(kutil.batch_generator(train, spec['feature_cols'], spec['label_cols'], spec['sample_weight_col']), **fit_kwargs)
    hist = {k: [float(x) for x in v] for k, v in getattr(history, 'history', {}).items()}
    data_module.__exit__(None, None, None)
    weights = kutil.weights_to_bytes(model.get_weights()) if hvd.rank() == 0 else None
    hvd.barrier()
    hvd.shutdown()
    return {'history': hist, 'weights': weights}


def RemoteTrainer(spec):
    """-> fn() for the backend to run on every rank."""
    def train():
        return _train_fn(spec)
    return train
