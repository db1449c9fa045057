"""Saving and loading estimators / models (reference horovod/spark/common/serialization.py `HorovodParamsWriter` :23-68 /
`HorovodParamsReader` :71-90, the MLWritable / MLReadable side of the Spark ML pipeline API).

Layout on disk is Spark ML's: `<path>/metadata/part-00000` holds one JSON document with the class name, a timestamp, the uid
and the param map.  Values JSON cannot carry (models, optimizers, losses, stores, callables) are pickled with cloudpickle —
torch modules through `torch.save`, so classes defined in a notebook travel by value — and stored base64-encoded under a
`{"__pickled__": ...}` marker, the same idea as the reference's `codec.dumps_base64` param values.  Nothing here needs
pyspark; `path` is a local directory or anything the given Store's filesystem understands.
"""
import base64
import importlib
import io
import json
import os
import time
import uuid

_PICKLED, _TORCH = '__pickled__', '__torch__'


def _encode(value):
    try:
        json.dumps(value)
        return value
    except (TypeError, ValueError):
        pass
    try:
        import torch
        if isinstance(value, torch.nn.Module):
            from horovod_b200.spark.torch.util import _serialize
            return {_TORCH: base64.b64encode(_serialize(value)).decode('ascii')}
    except ImportError:  # pragma: no cover
        pass
    import cloudpickle
    return {_PICKLED: base64.b64encode(cloudpickle.dumps(value)).decode('ascii')}


def _decode(value):
    if isinstance(value, dict) and len(value) == 1:
        if _PICKLED in value:
            import cloudpickle
            return cloudpickle.loads(base64.b64decode(value[_PICKLED]))
        if _TORCH in value:
            from horovod_b200.spark.torch.util import _deserialize
            return _deserialize(base64.b64decode(value[_TORCH]))
    return value


class _Target:
    """Local directory or a path inside a Store."""

    def __init__(self, path, store=None):
        self.path, self.store = path.rstrip('/'), store

    def _file(self):
        return self.path + '/metadata/part-00000'

    def exists(self):
        return self.store.exists(self._file()) if self.store is not None else os.path.exists(self._file())

    def write(self, text):
        if self.store is not None:
            return self.store.write_text(self._file(), text)
        os.makedirs(os.path.dirname(self._file()), exist_ok=True)
        with io.open(self._file(), 'w', encoding='utf-8') as f:
            f.write(text)

    def read(self):
        if self.store is not None:
            return self.store.read(self._file()).decode('utf-8')
        with io.open(self._file(), encoding='utf-8') as f:
            return f.read()


class HorovodParamsWriter:
    def __init__(self, instance):
        self.instance = instance
        self._overwrite = False
        self._store = None

    def overwrite(self):
        self._overwrite = True
        return self

    def option(self, key, value):
        """`option('store', store)` writes through that Store's filesystem instead of the local one."""
        if key == 'store':
            self._store = value
        return self

    def save(self, path):
        target = _Target(path, self._store)
        if target.exists() and not self._overwrite:
            raise IOError('Path %s already exists. To overwrite it, use write().overwrite().save(path).' % path)
        self.saveImpl(target)

    def saveImpl(self, target):
        if not isinstance(target, _Target):
            target = _Target(target, self._store)
        target.write(json.dumps(self._get_metadata_to_save(self.instance), separators=(',', ':')))

    @staticmethod
    def _get_metadata_to_save(instance, extra_metadata=None, param_map=None):
        cls = type(instance)
        params = param_map if param_map is not None else instance.param_dict()
        doc = {'class': cls.__module__ + '.' + cls.__qualname__, 'timestamp': int(round(time.time() * 1000)),
               'uid': getattr(instance, 'uid', None) or '%s_%s' % (cls.__name__, uuid.uuid4().hex[:12]),
               'paramMap': {k: _encode(v) for k, v in params.items()},
               'horovodVersion': __import__('horovod_b200').__version__}
        if extra_metadata:
            doc.update(extra_metadata)
        return doc

    saveMetadata = saveImpl


class HorovodParamsReader:
    def __init__(self, cls=None):
        self.cls = cls
        self._store = None

    def option(self, key, value):
        if key == 'store':
            self._store = value
        return self

    def load(self, path):
        doc = json.loads(_Target(path, self._store).read())
        cls = self.cls
        saved = doc['class']
        if cls is None or (cls.__module__ + '.' + cls.__qualname__) != saved:
            mod, _, name = saved.rpartition('.')
            found = getattr(importlib.import_module(mod), name)
            if cls is not None and not issubclass(found, cls):
                raise TypeError('%s holds a %s, not a %s' % (path, saved, cls.__name__))
            cls = found
        params = {k: _decode(v) for k, v in doc['paramMap'].items()}
        known = {k: v for k, v in params.items() if k in cls._table}
        instance = cls.__new__(cls)
        # constructors validate combinations (e.g. REQUIRED) that a saved instance already passed: restore the values directly
        from horovod_b200.spark.common.params import ParamsBase
        ParamsBase.__init__(instance)
        for k, v in known.items():
            instance._values[k] = v
        instance.uid = doc.get('uid')
        return instance
