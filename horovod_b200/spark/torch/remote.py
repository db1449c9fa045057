"""The per-rank training function of the torch estimator (reference horovod/spark/torch/remote.py `RemoteTrainer` :36-460:
shard the Parquet files by rank, wrap the optimizer in DistributedOptimizer, broadcast the initial state, average the epoch
metrics, checkpoint on rank 0, resume from the run's checkpoint)."""
import torch

from horovod_b200.spark.torch.util import _deserialize, _serialize


def _as_list(x, n):
    if x is None:
        return [None] * n
    if isinstance(x, (list, tuple)):
        if len(x) != n:
            raise ValueError('expected %d entries, got %d' % (n, len(x)))
        return list(x)
    return [x] * n


class _BatchLoss:
    """sum_i weight_i * loss_i(output_i, label_i), with optional per-row sample weights; also evaluates the metrics."""

    def __init__(self, model, loss, loss_weights, metrics, feature_cols, label_cols, sample_weight_col):
        n = len(label_cols)
        self.model, self.feature_cols, self.label_cols, self.sample_weight_col = model, feature_cols, label_cols, sample_weight_col
        self.losses = _as_list(loss, n)
        self.weights = [1.0 if w is None else float(w) for w in _as_list(loss_weights, n)]
        self.metrics = list(metrics or [])

    def _pairs(self, batch):
        out = self.model(*[batch[c].float() if batch[c].dtype.is_floating_point else batch[c] for c in self.feature_cols])
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        for o, col in zip(outs, self.label_cols):
            y = batch[col]
            if o.dtype.is_floating_point and y.dtype.is_floating_point:
                y = y.to(o.dtype)
            if o.dim() == y.dim() + 1 and o.shape[-1] == 1:
                o = o.squeeze(-1)
            yield o, y

    def __call__(self, batch, with_metrics=False):
        total, extra = 0.0, {}
        for i, (o, y) in enumerate(self._pairs(batch)):
            l = self.losses[i](o, y)
            if l.dim() > 0:
                if self.sample_weight_col:
                    w = batch[self.sample_weight_col].to(l.dtype)
                    l = l * w.reshape([-1] + [1] * (l.dim() - 1))
                l = l.mean()
            total = total + self.weights[i] * l
            if with_metrics:
                for m in self.metrics:
                    name = getattr(m, '__name__', type(m).__name__) + ('' if len(self.label_cols) == 1 else '_%d' % i)
                    extra[name] = torch.as_tensor(m(o.detach(), y), dtype=torch.float32, device=o.device).mean()
        return (total, extra) if with_metrics else total


def _train_fn(spec):
    """Runs on every rank; `spec` is the plain dict built by TorchEstimator._fit_on_prepared_data."""
    import horovod_b200.torch as hvd
    from horovod_b200.data import DevicePrefetcher
    from horovod_b200.spark.common.util import make_transform
    from horovod_b200.spark.torch.datamodule import ParquetDataModule
    hvd.init()
    from horovod_b200.spark.common.util import gpu_index_for
    dev = torch.device('cuda', gpu_index_for(hvd.local_rank())) if spec['use_gpu'] and torch.cuda.is_available() else torch.device('cpu')
    if dev.type == 'cuda':
        torch.cuda.set_device(dev)
    store = spec['store']
    model = _deserialize(spec['model']).to(dev)
    opt = spec['optimizer_cls'](model.parameters(), **spec['optimizer_defaults'])
    first_epoch = 0
    resume = spec['resume']
    if resume is not None and hvd.rank() == 0:            # rank 0 loads, everybody receives by broadcast
        ck = _deserialize(resume)
        model.load_state_dict(ck['model'])
        opt.load_state_dict(ck['optimizer'])
        first_epoch = ck['epoch'] + 1
    first_epoch = hvd.broadcast_object(first_epoch, root_rank=0, name='est.first_epoch')
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters(),
                                   compression=spec['compression'] or hvd.Compression.none,
                                   backward_passes_per_step=spec['backward_passes_per_step'])
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    hvd.broadcast_optimizer_state(opt, root_rank=0)

    cols = list(spec['feature_cols']) + list(spec['label_cols']) + ([spec['sample_weight_col']] if spec['sample_weight_col'] else [])
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
                             categorical_cols=spec.get('categorical_cols'), continuous_cols=spec.get('continuous_cols'))
    data_module.__enter__()
    train = data_module.train_data()
    val = data_module.val_data() if spec['val_path'] else None
    losses = [make() for make in spec['loss_constructors']] if spec.get('loss_constructors') else spec['loss']
    batch_loss = _BatchLoss(model, losses, spec['loss_weights'], spec['metrics'], spec['feature_cols'], spec['label_cols'],
                            spec['sample_weight_col'])
    accumulate = spec['backward_passes_per_step']

    def averaged(sums, count, prefix):
        names = sorted(sums)
        if not names:
            return {}
        vec = torch.stack([sums[n] for n in names]) / max(count, 1)
        vec = hvd.allreduce(vec, name='est.%smetrics' % prefix)
        return {prefix + n: v for n, v in zip(names, vec.tolist())}

    history = []
    for epoch in range(first_epoch, spec['epochs']):
        model.train()
        sums, count = {'loss': torch.zeros((), device=dev)}, 0
        opt.zero_grad()
        for step, batch in enumerate(DevicePrefetcher(train, device=dev)):
            loss = batch_loss(batch)
            (loss / accumulate).backward()
            if (step + 1) % accumulate == 0:
                opt.step()
                opt.zero_grad()
            sums['loss'] += loss.detach()
            count += 1
        record = {'epoch': epoch}
        record.update(averaged(sums, count, ''))
        if val is not None:
            model.eval()
            vsums, vcount = {'loss': torch.zeros((), device=dev)}, 0
            with torch.no_grad():
                for batch in DevicePrefetcher(val, device=dev):
                    loss, extra = batch_loss(batch, with_metrics=True)
                    vsums['loss'] += loss
                    for k, v in extra.items():
                        vsums[k] = vsums.get(k, torch.zeros((), device=dev)) + v
                    vcount += 1
            record.update(averaged(vsums, vcount, 'val_'))
        history.append(record)
        for cb in spec['callbacks']:
            cb(epoch, record) if callable(cb) else cb.on_epoch_end(epoch, record)
        if spec['verbose'] and hvd.rank() == 0:
            print('epoch %d: %s' % (epoch, record), flush=True)
        if spec['ckpt_path'] and hvd.rank() == 0:
            store.write(spec['ckpt_path'], _serialize({'model': model.state_dict(), 'optimizer': opt.state_dict(), 'epoch': epoch}))
    data_module.__exit__(None, None, None)
    state = {k: v.cpu() for k, v in model.state_dict().items()} if hvd.rank() == 0 else None
    hvd.barrier()  # shutdown is job-wide: nobody leaves while a peer still talks to the runtime
    hvd.shutdown()
    return {'history': history, 'state_dict': state}


def RemoteTrainer(spec):
    """-> fn() run on every rank by the backend (the reference builds the closure from the estimator, its metadata and the
    dataset properties; here all of that is the plain dict `TorchEstimator._fit_on_prepared_data` assembles)."""
    def train():
        return _train_fn(spec)
    return train
