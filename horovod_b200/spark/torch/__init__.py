from horovod_b200.spark.torch.estimator import TorchEstimator, TorchModel  # noqa: F401
