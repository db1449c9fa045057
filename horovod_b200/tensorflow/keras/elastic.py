"""Elastic Keras helpers (parity: horovod/tensorflow/keras/elastic.py)."""
from horovod_b200._keras.callbacks import CommitStateCallback, UpdateBatchStateCallback, UpdateEpochStateCallback  # noqa: F401
from horovod_b200.tensorflow.elastic import TensorFlowKerasState as KerasState, run  # noqa: F401
from horovod_b200.tensorflow.elastic import TensorFlowKerasState  # noqa: E402,F401
