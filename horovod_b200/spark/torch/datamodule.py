"""Data modules of the torch estimators (reference horovod/spark/torch/datamodule.py: `PetastormDataModule` :24-136,
`MapIterable` :138).

The reference reads its intermediate Parquet through Petastorm readers; here the reader is pyarrow.dataset
(`spark.data_loaders.ParquetShard`: every rank takes whole files, pads to equal step counts) and batches come out as dicts of
torch tensors, optionally in pinned memory for the side-stream `DevicePrefetcher`."""
from horovod_b200.spark.common.datamodule import DataModule
from horovod_b200.spark.data_loaders import (ParquetShard, PytorchAsyncDataLoader, PytorchDataLoader, PytorchInmemAsyncDataLoader,
                                              PytorchInmemDataLoader)


class ParquetDataModule(DataModule):
    short_name = 'parquet'

    def __init__(self, *args, store=None, row_shapes=None, seed=0, pin_memory=False, train_reader_num_workers=None,
                 val_reader_num_workers=None, train_async_data_loader_queue_size=64, val_async_data_loader_queue_size=64,
                 debug_data_loader=False, **kwargs):
        """`*_reader_num_workers` >= 1: batches are decoded / pinned on a background thread that runs ahead of the training
        loop by up to `*_async_data_loader_queue_size` batches."""
        super().__init__(*args, **kwargs)
        self.train_async = (train_reader_num_workers or 0, train_async_data_loader_queue_size)
        self.val_async = (val_reader_num_workers or 0, val_async_data_loader_queue_size)
        self.debug_data_loader = debug_data_loader
        if store is None:
            from horovod_b200.spark.common.store import Store
            store = Store.create(self.train_dir.rsplit('/', 1)[0])
        self.store, self.row_shapes, self.seed, self.pin_memory = store, row_shapes, seed, pin_memory
        self._shards, self._async = [], []

    def _loader(self, path, batch_size, shuffle, steps, async_cfg):
        shard = ParquetShard(self.store, path, list(self.schema_fields), self.cur_shard, self.shard_count, self.row_shapes)
        self._shards.append(shard)
        kwargs = dict(batch_size=batch_size, shuffle=shuffle, seed=self.seed, steps=steps, transformation_fn=self.transform_fn,
                      pin_memory=self.pin_memory)
        workers, depth = async_cfg
        if workers >= 1 and depth > 0:
            cls = PytorchInmemAsyncDataLoader if self.inmemory_cache_all else PytorchAsyncDataLoader
            loader = cls(shard, async_loader_queue_size=depth, debug_data_loader=self.debug_data_loader, **kwargs)
            self._async.append(loader)
            return loader
        cls = PytorchInmemDataLoader if self.inmemory_cache_all else PytorchDataLoader
        return cls(shard, **kwargs)

    def train_data(self):
        return self._loader(self.train_dir, self.train_batch_size, self.shuffle, self.steps_per_epoch_train, self.train_async)

    def val_data(self):
        if not self.has_val or not self.val_dir:
            return None
        return self._loader(self.val_dir, self.val_batch_size, False, self.steps_per_epoch_val, self.val_async)

    def __exit__(self, type, value, traceback):
        for loader in self._async:
            if hasattr(loader, 'close_async_loader'):
                loader.close_async_loader()
        for shard in self._shards:
            shard.release()
        self._shards, self._async = [], []


PetastormDataModule = ParquetDataModule          # the name estimators written against the reference pass as data_module


class MapIterable:
    """Re-iterable view of `data` with `map_fn` applied to every item, `epochs` times (None = forever)."""

    def __init__(self, data, epochs=None, map_fn=lambda x: x):
        self.data, self.epochs, self.map_fn = data, epochs, map_fn

    def __iter__(self):
        done = 0
        while self.epochs is None or done < self.epochs:
            for x in self.data:
                yield self.map_fn(x)
            done += 1
