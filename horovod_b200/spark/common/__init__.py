from horovod_b200.spark.common.store import Store, LocalStore, FilesystemStore  # noqa: F401
from horovod_b200.spark.common.backend import Backend, LocalBackend, SparkBackend  # noqa: F401
