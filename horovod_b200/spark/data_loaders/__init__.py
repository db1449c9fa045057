from horovod_b200.spark.data_loaders.parquet_shards import ParquetShard, shard_files  # noqa: F401
from horovod_b200.spark.data_loaders.pytorch_data_loaders import (  # noqa: F401
    PytorchDataLoader, PytorchInfiniteDataLoader, PytorchAsyncDataLoader, PytorchInfiniteAsyncDataLoader,
    PytorchInmemDataLoader, PytorchInmemAsyncDataLoader)
