from horovod_b200.spark.lightning.estimator import LightningEstimator, LightningModel, TorchEstimator, TorchModel  # noqa: F401
from horovod_b200.spark.lightning.trainer import ModuleProtocolTrainer, to_lightning_module  # noqa: F401
