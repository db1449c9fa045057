"""Torch batch loaders over a `ParquetShard`.

Role parity: horovod/spark/data_loaders/pytorch_data_loaders.py (`PytorchDataLoader`, `PytorchInfiniteDataLoader`,
`PytorchInmemDataLoader` and their `Async` variants built with `AsyncDataLoaderMixin`).  The reference wraps Petastorm's
`BatchedDataLoader` / `InMemBatchedDataLoader`; these loaders slice index permutations out of the decoded numpy columns
and hand out dicts of CPU tensors — pinned when `pin_memory` — so the trainer's `DevicePrefetcher` can issue non-blocking
H2D copies on its side stream.

finite   : one epoch = `len(loader)` batches, new permutation every epoch
infinite : an endless stream with per-pass reshuffling; the trainer takes `len(loader)` batches per epoch and the
           position carries over (no short last batch, no re-decode)
in-memory: the decoded shard is kept for the life of the loader (finite/infinite loaders drop it after every epoch unless
           `inmemory_cache_all`)
"""
import numpy as np
import torch

from horovod_b200.data import AsyncDataLoaderMixin, BaseDataLoader


def _to_tensor(arr, pin):
    if isinstance(arr, np.ndarray) and arr.dtype == object:
        arr = np.stack([np.asarray(v) for v in arr])
    t = torch.as_tensor(arr)
    return t.pin_memory() if pin and torch.cuda.is_available() else t


class PytorchDataLoader(BaseDataLoader):
    def __init__(self, shard, batch_size=32, shuffle=True, seed=0, steps=None, transformation_fn=None, pin_memory=False,
                 inmemory_cache_all=False, name=''):
        self.shard, self.batch_size, self.shuffle, self.seed = shard, batch_size, shuffle, seed
        self.steps = steps or shard.steps(batch_size)
        self.transformation_fn, self.pin_memory = transformation_fn, pin_memory
        self.keep_decoded, self.name = inmemory_cache_all, name
        self.epoch = 0

    def __len__(self):
        return self.steps

    def _order(self, n, epoch):
        if not self.shuffle:
            return np.arange(n)
        return np.random.RandomState((self.seed * 1000003 + epoch) % (2 ** 31)).permutation(n)

    def _batch(self, data, idx):
        out = {c: _to_tensor(v[idx], self.pin_memory) for c, v in data.items()}
        return self.transformation_fn(out) if self.transformation_fn else out

    def _iterate(self):
        data = self.shard.load()
        n = self.shard.rows
        order = self._order(n, self.epoch)
        self.epoch += 1
        for s in range(self.steps):
            idx = order[(np.arange(self.batch_size) + s * self.batch_size) % n]
            yield self._batch(data, np.sort(idx) if not self.shuffle else idx)
        if not self.keep_decoded:
            self.shard.release()


class PytorchInfiniteDataLoader(PytorchDataLoader):
    """`len()` batches per `__iter__`, continuing where the previous epoch stopped."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.keep_decoded = True
        self._stream = None

    def _endless(self):
        data = self.shard.load()
        n = self.shard.rows
        sweep = 0
        while True:
            order = self._order(n, sweep)
            sweep += 1
            for start in range(0, n - self.batch_size + 1, self.batch_size) if n >= self.batch_size else [0]:
                idx = order[(np.arange(self.batch_size) + start) % n]
                yield self._batch(data, idx)

    def _iterate(self):
        if self._stream is None:
            self._stream = self._endless()
        for _ in range(self.steps):
            yield next(self._stream)


class PytorchInmemDataLoader(PytorchDataLoader):
    """Decodes once and converts the whole shard to tensors up front: batches are pure index_select calls."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.keep_decoded = True
        self._tensors = None

    def _iterate(self):
        if self._tensors is None:
            self._tensors = {c: _to_tensor(v, self.pin_memory) for c, v in self.shard.load().items()}
            self.shard.release()
        n = self.shard.rows
        order = torch.as_tensor(self._order(n, self.epoch))
        self.epoch += 1
        for s in range(self.steps):
            idx = order[(torch.arange(self.batch_size) + s * self.batch_size) % n]
            out = {c: v.index_select(0, idx) for c, v in self._tensors.items()}
            yield self.transformation_fn(out) if self.transformation_fn else out


class PytorchAsyncDataLoader(AsyncDataLoaderMixin, PytorchDataLoader):
    pass


class PytorchInfiniteAsyncDataLoader(AsyncDataLoaderMixin, PytorchInfiniteDataLoader):
    pass


class PytorchInmemAsyncDataLoader(AsyncDataLoaderMixin, PytorchInmemDataLoader):
    pass
