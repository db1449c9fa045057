"""Keras callbacks (parity: horovod/keras/callbacks.py)."""
from horovod_b200._keras.callbacks import (  # noqa: F401
    BroadcastGlobalVariablesCallback, MetricAverageCallback, LearningRateScheduleCallback, LearningRateWarmupCallback,
    BestModelCheckpoint)
