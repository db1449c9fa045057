"""Elastic Keras helpers (parity: horovod/keras/elastic.py)."""
from horovod_b200.tensorflow.keras.elastic import (  # noqa: F401
    CommitStateCallback, UpdateBatchStateCallback, UpdateEpochStateCallback, KerasState, run)
from horovod_b200.tensorflow.elastic import TensorFlowKerasState  # noqa: E402,F401
