"""(De)serialisation helpers of the Lightning estimator (reference horovod/spark/lightning/util.py; same functions as the torch
estimator's: LightningModules are nn.Modules)."""
from horovod_b200.spark.torch.util import (  # noqa: F401
    _deserialize, _serialize, deserialize_fn, is_module_available, is_module_available_fn, save_into_bio, save_into_bio_fn,
    serialize_fn)
