"""TorchEstimator / TorchModel: fit a torch model on a DataFrame with data-parallel training, get back a transformer.

Parity: horovod/spark/torch/estimator.py (`TorchEstimator` :94-353 — params model/optimizer/loss/feature_cols/label_cols/
batch_size/epochs/validation/sample_weight_col/store/backend/num_proc/shuffle/train_steps_per_epoch/verbose;
`TorchModel.transform` :355-500) and spark/torch/remote.py (the per-rank training function: shard the parquet files by
rank, wrap the optimizer in DistributedOptimizer, broadcast the initial state, average the epoch metrics, checkpoint on
rank 0).

The reference materialises the DataFrame to Parquet in the Store and reads it back through Petastorm.  Here the
intermediate format is the same (Parquet in the Store) but the reader is pyarrow.dataset + a pinned side-stream
`DevicePrefetcher`, and the input may be a Spark DataFrame (written by Spark itself) or a pandas DataFrame (written by
pyarrow) — so the estimator also works on a single multi-GPU box without Spark (`LocalBackend`).
"""
import io
import time
import uuid

import torch

from horovod_b200.spark.common.backend import LocalBackend, SparkBackend
from horovod_b200.spark.common.store import Store


def _is_spark_df(df):
    return type(df).__module__.startswith('pyspark.')


def _write_parquet(df, path, store, num_files):
    store.delete(path)
    if _is_spark_df(df):
        df.repartition(num_files).write.mode('overwrite').parquet(path)
        return df.count()
    import os
    import pyarrow as pa
    import pyarrow.parquet as pq
    table = pa.Table.from_pandas(df, preserve_index=False)
    local = store._local(path)
    store.fs.create_dir(local, recursive=True)
    n = len(df)
    per = -(-n // num_files)
    for i in range(num_files):
        if i * per < n:
            pq.write_table(table.slice(i * per, per), os.path.join(local, f'part-{i:05d}.parquet'), filesystem=store.fs)
    return n


def _to_tensor(col_values):
    import numpy as np
    first = col_values[0] if len(col_values) else 0.0
    if isinstance(first, (list, tuple, np.ndarray)):
        return torch.as_tensor(np.stack([np.asarray(v) for v in col_values]))
    return torch.as_tensor(np.asarray(col_values))


class _ParquetShardLoader:
    """Batches from the row groups of this rank's share of a Parquet dataset (files are dealt round-robin by rank; every
    rank sees the same number of batches so that no rank runs out of collectives early)."""

    def __init__(self, store, path, columns, batch_size, rank, size, shuffle, seed, steps=None):
        import pyarrow.dataset as ds
        self.dataset = ds.dataset(store._local(path), format='parquet', filesystem=store.fs)
        frags = sorted(self.dataset.get_fragments(), key=lambda f: f.path)
        self.frags = [f for i, f in enumerate(frags) if i % size == rank] or frags[rank % len(frags):rank % len(frags) + 1]
        self.columns, self.batch_size, self.shuffle, self.seed = columns, batch_size, shuffle, seed
        rows = [sum(f.count_rows() for i, f in enumerate(frags) if i % size == r) for r in range(size)]
        self.steps = steps or max(1, min(r for r in rows if r > 0) // batch_size) if any(rows) else 0
        self.epoch = 0

    def __len__(self):
        return self.steps

    def __iter__(self):
        import pyarrow as pa
        table = pa.concat_tables([f.to_table(columns=self.columns) for f in self.frags])
        cols = {c: _to_tensor(table.column(c).to_pylist()) for c in self.columns}
        n = table.num_rows
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        self.epoch += 1
        order = torch.randperm(n, generator=g) if self.shuffle else torch.arange(n)
        for s in range(self.steps):
            idx = order[(torch.arange(self.batch_size) + s * self.batch_size) % n]
            yield {c: v[idx] for c, v in cols.items()}


def _serialize(obj):
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def _train_fn(model_bytes, opt_cls, opt_defaults, loss_fn, feature_cols, label_cols, sample_weight_col, batch_size, epochs,
              store, train_path, val_path, ckpt_path, shuffle, seed, steps, use_gpu, verbose):
    """Runs on every rank."""
    import horovod_b200.torch as hvd
    from horovod_b200.data import DevicePrefetcher
    hvd.init()
    dev = torch.device('cuda', hvd.local_rank()) if use_gpu and torch.cuda.is_available() else torch.device('cpu')
    if dev.type == 'cuda':
        torch.cuda.set_device(dev)
    model = torch.load(io.BytesIO(model_bytes), weights_only=False).to(dev)
    opt = opt_cls(model.parameters(), **opt_defaults)
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    hvd.broadcast_optimizer_state(opt, root_rank=0)
    cols = list(feature_cols) + list(label_cols) + ([sample_weight_col] if sample_weight_col else [])
    train = _ParquetShardLoader(store, train_path, cols, batch_size, hvd.rank(), hvd.size(), shuffle, seed, steps)
    val = _ParquetShardLoader(store, val_path, cols, batch_size, hvd.rank(), hvd.size(), False, seed) if val_path else None

    def batch_loss(b):
        feats = [b[c].float() for c in feature_cols]
        out = model(*feats)
        outs = out if isinstance(out, (tuple, list)) else [out]
        losses = []
        for o, c in zip(outs, label_cols):
            y = b[c]
            y = y.float() if o.dtype.is_floating_point and y.dtype.is_floating_point else y
            if o.dim() == y.dim() + 1 and o.shape[-1] == 1:
                o = o.squeeze(-1)
            l = loss_fn(o, y)
            if sample_weight_col and l.dim() > 0:
                l = (l * b[sample_weight_col].float()).mean()
            losses.append(l.mean() if l.dim() > 0 else l)
        return sum(losses)

    history = []
    for epoch in range(epochs):
        model.train()
        tot, cnt = torch.zeros((), device=dev), 0
        for b in DevicePrefetcher(train, device=dev):
            opt.zero_grad()
            loss = batch_loss(b)
            loss.backward()
            opt.step()
            tot += loss.detach()
            cnt += 1
        rec = {'epoch': epoch, 'loss': hvd.allreduce(tot / max(cnt, 1), name='est.train_loss').item()}
        if val is not None:
            model.eval()
            vt, vc = torch.zeros((), device=dev), 0
            with torch.no_grad():
                for b in DevicePrefetcher(val, device=dev):
                    vt += batch_loss(b)
                    vc += 1
            rec['val_loss'] = hvd.allreduce(vt / max(vc, 1), name='est.val_loss').item()
        history.append(rec)
        if verbose and hvd.rank() == 0:
            print('epoch %d: %s' % (epoch, rec), flush=True)
        if ckpt_path and hvd.rank() == 0:
            store.write(ckpt_path, _serialize({'model': model.state_dict(), 'optimizer': opt.state_dict(), 'epoch': epoch}))
    state = {k: v.cpu() for k, v in model.state_dict().items()} if hvd.rank() == 0 else None
    hvd.barrier()  # shutdown is job-wide: nobody leaves while a peer still talks to the runtime
    hvd.shutdown()
    return {'history': history, 'state_dict': state}


class TorchEstimator:
    """fit(df) -> TorchModel.

    Args mirror the reference's Params: `model` (nn.Module), `optimizer` (a torch optimizer INSTANCE built on the model —
    its class and defaults are re-created on every rank), `loss` (callable(output, label); reduction='none' when
    sample_weight_col is used), `feature_cols`, `label_cols`, `batch_size`, `epochs`, `validation` (None | float fraction |
    column name marking validation rows), `store`, `backend` or `num_proc`, `shuffle`, `train_steps_per_epoch`, `verbose`.
    """

    def __init__(self, model=None, optimizer=None, loss=None, feature_cols=None, label_cols=None, batch_size=32, epochs=1,
                 validation=None, sample_weight_col=None, store=None, backend=None, num_proc=None, shuffle=True, random_seed=0,
                 train_steps_per_epoch=None, use_gpu=True, verbose=1, run_id=None):
        if model is None or optimizer is None or loss is None:
            raise ValueError('model, optimizer and loss are required')
        if not feature_cols or not label_cols:
            raise ValueError('feature_cols and label_cols are required')
        if backend is not None and num_proc is not None:
            raise ValueError('At most one of parameters "backend" and "num_proc" may be specified')
        self.model, self.optimizer, self.loss = model, optimizer, loss
        self.feature_cols, self.label_cols = list(feature_cols), list(label_cols)
        self.batch_size, self.epochs, self.validation = batch_size, epochs, validation
        self.sample_weight_col, self.shuffle, self.random_seed = sample_weight_col, shuffle, random_seed
        self.train_steps_per_epoch, self.use_gpu, self.verbose, self.run_id = train_steps_per_epoch, use_gpu, verbose, run_id
        self.store = Store.create(store) if isinstance(store, str) else store
        if self.store is None:
            raise ValueError('store is required (a Store or a path prefix)')
        self.backend, self.num_proc = backend, num_proc

    def _get_backend(self, df):
        if self.backend is not None:
            return self.backend
        return SparkBackend(self.num_proc) if _is_spark_df(df) else LocalBackend(self.num_proc or 1)

    def _split(self, df):
        if self.validation is None:
            return df, None
        if isinstance(self.validation, str):
            if _is_spark_df(df):
                return df.filter(~df[self.validation].cast('boolean')), df.filter(df[self.validation].cast('boolean'))
            m = df[self.validation].astype(bool)
            return df[~m].drop(columns=[self.validation]), df[m].drop(columns=[self.validation])
        frac = float(self.validation)
        if not 0 < frac < 1:
            raise ValueError('validation must be a column name or a fraction in (0, 1)')
        if _is_spark_df(df):
            tr, va = df.randomSplit([1 - frac, frac], seed=self.random_seed)
            return tr, va
        va = df.sample(frac=frac, random_state=self.random_seed)
        return df.drop(va.index), va

    def fit(self, df):
        backend = self._get_backend(df)
        n = backend.num_processes()
        run_id = self.run_id or 'run_' + time.strftime('%Y%m%d_%H%M%S') + '_' + uuid.uuid4().hex[:6]
        train_df, val_df = self._split(df)
        idx = uuid.uuid4().hex[:8]
        train_path, val_path = self.store.get_train_data_path(idx), self.store.get_val_data_path(idx)
        rows = _write_parquet(train_df, train_path, self.store, n)
        if rows < n:
            raise ValueError(f'{rows} training rows cannot be spread over {n} processes')
        if val_df is not None:
            _write_parquet(val_df, val_path, self.store, n)
        ckpt = self.store.get_checkpoint_path(run_id)
        opt_defaults = {k: v for k, v in self.optimizer.defaults.items() if k not in ('differentiable', 'foreach', 'fused', 'capturable', 'maximize') or v}
        results = backend.run(_train_fn, args=(_serialize(self.model), type(self.optimizer), opt_defaults, self.loss, self.feature_cols,
                                               self.label_cols, self.sample_weight_col, self.batch_size, self.epochs, self.store,
                                               train_path, val_path if val_df is not None else None, ckpt, self.shuffle,
                                               self.random_seed, self.train_steps_per_epoch, self.use_gpu, self.verbose))
        r0 = results[0]
        self.model.load_state_dict(r0['state_dict'])
        self.store.delete(train_path)
        if val_df is not None:
            self.store.delete(val_path)
        return TorchModel(self.model, self.feature_cols, self.label_cols, history=r0['history'], run_id=run_id)


class TorchModel:
    """Transformer returned by fit(): appends `<label>__output` prediction columns."""

    def __init__(self, model, feature_cols, label_cols, history=None, run_id=None, output_cols=None):
        self.model, self.feature_cols, self.label_cols = model, list(feature_cols), list(label_cols)
        self.history, self.run_id = history or [], run_id
        self.output_cols = output_cols or [c + '__output' for c in self.label_cols]

    def getModel(self):
        return self.model

    def getHistory(self):
        return self.history

    def _predict_pandas(self, pdf):
        self.model.eval()
        with torch.no_grad():
            feats = [_to_tensor(pdf[c].tolist()).float() for c in self.feature_cols]
            out = self.model.cpu()(*feats)
        outs = out if isinstance(out, (tuple, list)) else [out]
        pdf = pdf.copy()
        for name, o in zip(self.output_cols, outs):
            o = o.squeeze(-1) if o.dim() > 1 and o.shape[-1] == 1 else o
            pdf[name] = o.numpy().tolist()
        return pdf

    def transform(self, df):
        if not _is_spark_df(df):
            return self._predict_pandas(df)
        import pandas as pd  # noqa: F401
        model = self

        def fn(iterator):
            for pdf in iterator:
                yield model._predict_pandas(pdf)
        sample = self._predict_pandas(df.limit(1).toPandas())
        schema = df.sparkSession.createDataFrame(sample).schema
        return df.mapInPandas(fn, schema=schema)
