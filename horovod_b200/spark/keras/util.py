"""Keras model / optimizer transport and batch assembly for the Keras estimator.

Role parity: horovod/spark/keras/util.py (`TFKerasUtil`: serialize / deserialize model, keras module selection, batch
preparation :30-285), spark/keras/bare.py / tensorflow.py (save_model / load_model with custom objects) and
spark/keras/optimizer.py (optimizer (de)serialisation).
"""
import io
import os
import tempfile

import numpy as np


def keras_module():
    import tensorflow as tf
    return tf.keras


def serialize_model(model):
    """-> bytes.  Uses the Keras file format when `keras.models.save_model` exists (architecture + weights + compile state),
    otherwise the portable pair (class, get_config(), get_weights())."""
    import cloudpickle
    keras = keras_module()
    models = getattr(keras, 'models', None)
    if models is not None and hasattr(models, 'save_model') and hasattr(models, 'load_model'):
        with tempfile.TemporaryDirectory() as d:
            for ext in ('.keras', '.h5'):
                path = os.path.join(d, 'model' + ext)
                try:
                    models.save_model(model, path)
                except Exception:  # noqa: BLE001 - format not supported by this Keras version
                    continue
                with open(path, 'rb') as f:
                    return cloudpickle.dumps({'format': ext, 'blob': f.read()})
    return cloudpickle.dumps({'format': 'config', 'cls': type(model), 'config': model.get_config(), 'weights': model.get_weights()})


def deserialize_model(data, custom_objects=None):
    import cloudpickle
    rec = cloudpickle.loads(data)
    if rec['format'] == 'config':
        cls = rec['cls']
        try:
            model = cls.from_config(rec['config'], custom_objects=custom_objects) if custom_objects else cls.from_config(rec['config'])
        except TypeError:
            model = cls.from_config(rec['config'])
        model.set_weights(rec['weights'])
        return model
    keras = keras_module()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'model' + rec['format'])
        with open(path, 'wb') as f:
            f.write(rec['blob'])
        return keras.models.load_model(path, custom_objects=custom_objects, compile=False)


def checkpoint_to_serialized_model(ckpt_bytes, model, custom_objects=None):
    """Bytes of a run's checkpoint (what `_StoreCheckpoint` writes: weights + epoch) -> bytes in `serialize_model` format,
    with `model` supplying the architecture.  A checkpoint that already is a serialized model passes through."""
    import cloudpickle
    rec = cloudpickle.loads(ckpt_bytes)
    if isinstance(rec, dict) and 'format' in rec:
        return ckpt_bytes
    model.set_weights(weights_from_bytes(rec['weights']))
    return serialize_model(model)


def serialize_optimizer(optimizer):
    import cloudpickle
    return cloudpickle.dumps((type(optimizer), optimizer.get_config()))


def deserialize_optimizer(data):
    import cloudpickle
    cls, config = cloudpickle.loads(data)
    return cls.from_config(config)


def weights_to_bytes(weights):
    buf = io.BytesIO()
    np.savez(buf, *[np.asarray(w) for w in weights])
    return buf.getvalue()


def weights_from_bytes(data):
    with np.load(io.BytesIO(data), allow_pickle=False) as z:
        return [z['arr_%d' % i] for i in range(len(z.files))]


def batch_generator(loader_epochs, feature_cols, label_cols, sample_weight_col=None):
    """Endless generator of (x, y[, w]) numpy batches for `model.fit`: x / y are single arrays for one column, tuples
    otherwise.  `loader_epochs()` returns a fresh iterable of {column: array-like} batches for each pass."""
    def pick(batch, cols):
        arrays = [np.asarray(batch[c]) for c in cols]
        return arrays[0] if len(arrays) == 1 else tuple(arrays)
    while True:
        for batch in loader_epochs():
            item = (pick(batch, feature_cols), pick(batch, label_cols))
            if sample_weight_col:
                item += (np.asarray(batch[sample_weight_col]),)
            yield item
