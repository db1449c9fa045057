"""(De)serialisation helpers of the torch estimator (reference horovod/spark/torch/util.py: is_module_available :23,
serialize_fn :46, deserialize_fn :65, save_into_bio :97)."""
import base64
import importlib.util
import io

import torch


def is_module_available(module_name):
    return importlib.util.find_spec(module_name) is not None


def is_module_available_fn():
    return is_module_available


def _serialize(obj):
    """torch.save through cloudpickle: a model class defined in a script / notebook / test module that the workers cannot
    import travels by value."""
    import cloudpickle
    if isinstance(obj, torch.nn.Module) and getattr(__import__('sys').modules.get(type(obj).__module__), '__file__', None):
        from horovod_b200.runner import _pickle_by_value_if_not_importable
        _pickle_by_value_if_not_importable(type(obj))
    buf = io.BytesIO()
    torch.save(obj, buf, pickle_module=cloudpickle)
    return buf.getvalue()


def _deserialize(data):
    return torch.load(io.BytesIO(data), weights_only=False)



def save_into_bio(obj, save_obj_fn):
    """save_obj_fn(obj, file) into an in-memory file, rewound for reading."""
    bio = io.BytesIO()
    save_obj_fn(obj, bio)
    bio.seek(0)
    return bio


def save_into_bio_fn():
    return save_into_bio


def serialize_fn():
    """-> fn(model) -> ascii string (base64 of the torch.save bytes): safe inside Spark ML params / JSON metadata."""
    def _ser(model):
        return base64.b64encode(_serialize(model)).decode('ascii')
    return _ser


def deserialize_fn():
    def _de(model_bytes_base64):
        return _deserialize(base64.b64decode(model_bytes_base64))
    return _de
