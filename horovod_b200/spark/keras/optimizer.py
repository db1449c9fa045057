"""Optimizer (de)serialisation of the Keras estimator (reference horovod/spark/keras/optimizer.py + bare.py + tensorflow.py:
the optimizer's config and slot weights written into an in-memory h5 file).

h5py is not a dependency here: an optimizer travels as (class, get_config(), weights) pickled by cloudpickle — the class by
value when the workers cannot import it — and is rebuilt with `from_config`; slot weights are restored once the optimizer
has been built against the model's variables."""
import base64

import cloudpickle
import numpy as np


def _weights_of(optimizer):
    get = getattr(optimizer, 'get_weights', None)
    if get is not None:
        try:
            return [np.asarray(w) for w in get()]
        except Exception:  # noqa: BLE001 - Keras 3 optimizers expose `variables` instead
            pass
    variables = getattr(optimizer, 'variables', None)
    if callable(variables):                       # a method up to Keras 2.x, a property in Keras 3
        variables = variables()
    return [np.asarray(v) for v in variables or []]


def _serialize_keras_optimizer(opt):
    return base64.b64encode(cloudpickle.dumps({'cls': type(opt), 'config': opt.get_config(), 'weights': _weights_of(opt)})).decode('ascii')


def _deserialize_keras_optimizer(serialized_opt, model=None):
    rec = cloudpickle.loads(base64.b64decode(serialized_opt))
    opt = rec['cls'].from_config(rec['config'])
    weights = rec.get('weights') or []
    if weights and model is not None:
        build = getattr(opt, 'build', None)
        if build is not None:
            try:
                build(model.trainable_variables)
            except Exception:  # noqa: BLE001 - legacy optimizers create slots lazily
                pass
        setter = getattr(opt, 'set_weights', None)
        if setter is not None:
            try:
                setter(weights)
            except Exception:  # noqa: BLE001 - shapes unknown until the first step: start from fresh slots
                pass
    return opt


def is_string(obj):
    return isinstance(obj, str)


serialize_tf_keras_optimizer = serialize_bare_keras_optimizer = _serialize_keras_optimizer


def deserialize_tf_keras_optimizer(x, model=None):
    return _deserialize_keras_optimizer(x, model)


def deserialize_bare_keras_optimizer(x):
    return _deserialize_keras_optimizer(x)
