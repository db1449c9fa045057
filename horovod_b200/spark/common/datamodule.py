"""Base class of the objects that feed an estimator's training loop (reference horovod/spark/common/datamodule.py:18-60).

A data module is created on every rank with that rank's shard coordinates, entered once, asked for its training (and
validation) iterable, and exited when training ends.  Estimators take the class through their `data_module` param, so a
different reader (another file format, a GPU data loader) needs no change to the training function."""
from abc import ABC, abstractmethod


class DataModule(ABC):
    short_name = None        # e.g. 'parquet'

    def __init__(self, train_dir, val_dir, num_train_epochs=1, has_val=True, train_batch_size=32, val_batch_size=32,
                 shuffle=True, transform_fn=None, inmemory_cache_all=False, cur_shard=0, shard_count=1, schema_fields=None,
                 storage_options=None, steps_per_epoch_train=None, steps_per_epoch_val=None, verbose=True, **kwargs):
        self.train_dir, self.val_dir = train_dir, val_dir
        self.num_train_epochs, self.has_val = num_train_epochs, has_val
        self.train_batch_size, self.val_batch_size = train_batch_size, val_batch_size
        self.shuffle, self.transform_fn, self.inmemory_cache_all = shuffle, transform_fn, inmemory_cache_all
        self.cur_shard, self.shard_count = cur_shard, shard_count
        self.schema_fields, self.storage_options = schema_fields, storage_options
        self.steps_per_epoch_train, self.steps_per_epoch_val = steps_per_epoch_train, steps_per_epoch_val
        self.verbose = verbose
        self.extra = kwargs

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        return None

    @abstractmethod
    def train_data(self):
        """The training data in the form the framework's loop iterates over (one pass = one epoch)."""

    @abstractmethod
    def val_data(self):
        """Same for validation data; None when the module was built with has_val=False."""
