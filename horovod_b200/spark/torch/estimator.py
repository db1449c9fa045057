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
import io

import numpy as np
import torch

from horovod_b200.spark.common.estimator import HorovodEstimator, HorovodModel
from horovod_b200.spark.common.params import P


def _serialize(obj):
    """torch.save through cloudpickle: a model class defined in a script / notebook / test module that the workers cannot
    import travels by value."""
    import cloudpickle
    if isinstance(obj, torch.nn.Module) and getattr(__import__('sys').modules.get(type(obj).__module__), '__file__', None):
        from horovod_b200.runner import _pickle_by_value_if_not_importable
        _pickle_by_value_if_not_importable(type(obj))
    buf = io.BytesIO()
    torch.save(obj, buf, pickle_module=cloudpickle)
    return buf.getvalue()


def _deserialize(data):
    return torch.load(io.BytesIO(data), weights_only=False)


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
    from horovod_b200.spark.data_loaders import ParquetShard, PytorchDataLoader, PytorchInmemDataLoader
    hvd.init()
    dev = torch.device('cuda', hvd.local_rank()) if spec['use_gpu'] and torch.cuda.is_available() else torch.device('cpu')
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
    loader_cls = PytorchInmemDataLoader if spec['inmemory_cache_all'] else PytorchDataLoader

    def loader(path, batch_size, shuffle, steps):
        shard = ParquetShard(store, path, cols, hvd.rank(), hvd.size(), spec['row_shapes'])
        return loader_cls(shard, batch_size=batch_size, shuffle=shuffle, seed=spec['seed'], steps=steps,
                          transformation_fn=spec['transformation_fn'], pin_memory=dev.type == 'cuda')
    train = loader(spec['train_path'], spec['batch_size'], spec['shuffle'], spec['train_steps'])
    val = loader(spec['val_path'], spec['val_batch_size'], False, spec['val_steps']) if spec['val_path'] else None
    batch_loss = _BatchLoss(model, spec['loss'], spec['loss_weights'], spec['metrics'], spec['feature_cols'], spec['label_cols'],
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
    state = {k: v.cpu() for k, v in model.state_dict().items()} if hvd.rank() == 0 else None
    hvd.barrier()  # shutdown is job-wide: nobody leaves while a peer still talks to the runtime
    hvd.shutdown()
    return {'history': history, 'state_dict': state}


class TorchEstimator(HorovodEstimator):
    """fit(df) -> TorchModel.

    `model` (nn.Module), `optimizer` (a torch optimizer INSTANCE built on the model — its class and defaults are re-created on
    every rank), `loss` (callable(output, label) or one per label column; reduction='none' when sample_weight_col is used),
    plus every knob of `EstimatorParams` (also reachable as setX/getX and through `fit(df, params={...})`).
    """
    PARAMS = (
        P('train_minibatch_fn', None, None, 'accepted for compatibility; the loop lives in spark/torch/estimator.py:_train_fn'),
    )
    REQUIRED = ('model', 'optimizer', 'loss', 'feature_cols', 'label_cols', 'store')

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._check_params()

    def _check_framework_params(self):
        if not isinstance(self._get('optimizer'), torch.optim.Optimizer):
            raise ValueError('optimizer must be a torch.optim.Optimizer instance')
        n = len(self._get('label_cols'))
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
            backward_passes_per_step=g('backward_passes_per_step'))
        results = backend.run(_train_fn, args=(spec,))
        rank0 = results[0]
        model = g('model')
        model.load_state_dict(rank0['state_dict'])
        return TorchModel(model=model, feature_columns=spec['feature_cols'], label_columns=spec['label_cols'], history=rank0['history'],
                          run_id=run_id, metadata=dataset.metadata, input_shapes=g('input_shapes'))


class TorchModel(HorovodModel):
    """Transformer returned by fit(): appends `<label>__output` prediction columns."""
    PARAMS = (
        P('input_shapes', None, None, 'one shape per feature column'),
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
