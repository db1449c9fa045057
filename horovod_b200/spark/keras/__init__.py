from horovod_b200.spark.keras.estimator import KerasEstimator, KerasModel  # noqa: F401
