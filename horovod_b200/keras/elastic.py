"""Elastic Keras helpers (parity: horovod/keras/elastic.py)."""
from horovod_b200.tensorflow.keras.elastic import (  # noqa: F401
    CommitStateCallback, UpdateBatchStateCallback, UpdateEpochStateCallback, KerasState, run)
