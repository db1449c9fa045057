"""Data module of the Keras estimator (reference horovod/spark/keras/datamodule.py `PetastormDataModule`: tf.data over
Petastorm readers).  `model.fit` is fed by Python generators of numpy batches from the rank's `ParquetShard`, so there is
no tf.data / Petastorm dependency; `train_data()` / `val_data()` return callables that start one pass."""
import numpy as np

from horovod_b200.spark.common.datamodule import DataModule
from horovod_b200.spark.data_loaders import ParquetShard


class _NumpyShardBatches:
    """One pass = `steps` dict-of-numpy batches from this rank's shard."""

    def __init__(self, shard, batch_size, shuffle, seed, steps, transformation_fn):
        self.shard, self.batch_size, self.shuffle, self.seed = shard, batch_size, shuffle, seed
        self.steps = steps or shard.steps(batch_size)
        self.transformation_fn = transformation_fn
        self.passes = 0

    def __call__(self):
        data, n = self.shard.load(), self.shard.rows
        rng = np.random.RandomState((self.seed * 1000003 + self.passes) % (2 ** 31))
        order = rng.permutation(n) if self.shuffle else np.arange(n)
        self.passes += 1
        for s in range(self.steps):
            idx = order[(np.arange(self.batch_size) + s * self.batch_size) % n]
            batch = {c: v[idx] for c, v in data.items()}
            yield self.transformation_fn(batch) if self.transformation_fn else batch


class ParquetDataModule(DataModule):
    short_name = 'parquet'

    def __init__(self, *args, store=None, row_shapes=None, seed=0, **kwargs):
        super().__init__(*args, **kwargs)
        self.store, self.row_shapes, self.seed = store, row_shapes, seed
        self._shards = []

    def _batches(self, path, batch_size, shuffle, steps):
        shard = ParquetShard(self.store, path, list(self.schema_fields), self.cur_shard, self.shard_count, self.row_shapes)
        self._shards.append(shard)
        return _NumpyShardBatches(shard, batch_size, shuffle, self.seed, steps, self.transform_fn)

    def train_data(self):
        return self._batches(self.train_dir, self.train_batch_size, self.shuffle, self.steps_per_epoch_train)

    def val_data(self):
        if not self.has_val or not self.val_dir:
            return None
        return self._batches(self.val_dir, self.val_batch_size, False, self.steps_per_epoch_val)

    def __exit__(self, type, value, traceback):
        for shard in self._shards:
            shard.release()
        self._shards = []


PetastormDataModule = ParquetDataModule
