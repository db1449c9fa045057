from horovod_b200.spark.common.store import (Store, LocalStore, FilesystemStore, HDFSStore, DBFSLocalStore,  # noqa: F401
                                              AbstractFilesystemStore, is_databricks)
from horovod_b200.spark.common.backend import Backend, LocalBackend, SparkBackend  # noqa: F401
from horovod_b200.spark.common.params import EstimatorParams, ModelParams  # noqa: F401
from horovod_b200.spark.common.estimator import HorovodEstimator, HorovodModel  # noqa: F401
from horovod_b200.spark.common.util import prepare_data, clear_training_cache  # noqa: F401
