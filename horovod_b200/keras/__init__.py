"""`import horovod_b200.keras as hvd` — stand-alone Keras entry point (parity: horovod/keras/__init__.py); Keras ≥ 2.4 is
tf.keras, so this is the same implementation as `horovod_b200.tensorflow.keras`."""
from horovod_b200.tensorflow.keras import *  # noqa: F401,F403
from horovod_b200.tensorflow.keras import DistributedOptimizer, PartialDistributedOptimizer, load_model  # noqa: F401
from horovod_b200.keras import callbacks, elastic  # noqa: F401
