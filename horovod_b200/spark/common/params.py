"""Estimator / model parameters.

Role parity: horovod/spark/common/params.py (`EstimatorParams` :24-378, `ModelParams` :380-519): every estimator knob is
reachable as a constructor keyword, a `setFooBar(value)` / `getFooBar()` pair and through `fit(df, params={...})`.

The reference declares each knob as a pyspark.ml `Param` plus two hand-written accessor methods.  Here ONE table row per
knob (`P(...)`) is the single source: name, default, validator, doc.  `ParamsBase.__init_subclass__` folds the tables of a
class hierarchy together and generates the camel-case accessors, so the estimators run with or without pyspark and a new
knob is one line.
"""
import copy


class P:
    """One parameter: `P('batch_size', 32, check=positive_int, doc='...')`."""
    __slots__ = ('name', 'default', 'check', 'doc', 'camels')

    def __init__(self, name, default=None, check=None, doc='', camel=()):
        """`camel`: accessor stems other than the mechanical CamelCase of `name` (the reference's `setInMemoryCacheAll`
        for `inmemory_cache_all`); every stem gets a set/get pair."""
        self.name, self.default, self.check, self.doc = name, default, check, doc
        head, *rest = name.split('_')
        own = ''.join([head.capitalize()] + [r.capitalize() for r in rest])
        self.camels = (own,) + tuple(c for c in ((camel,) if isinstance(camel, str) else camel) if c != own)

    @property
    def camel(self):
        return self.camels[0]


def _positive_int(name, v):
    if v is not None and (not isinstance(v, int) or isinstance(v, bool) or v < 1):
        raise ValueError('%s must be a positive integer, got %r' % (name, v))


def _non_negative_int(name, v):
    if v is not None and (not isinstance(v, int) or isinstance(v, bool) or v < 0):
        raise ValueError('%s must be a non-negative integer, got %r' % (name, v))


def _str_list(name, v):
    if v is None:
        return
    if isinstance(v, str) or not all(isinstance(c, str) for c in v):
        raise ValueError('%s must be a list of column names, got %r' % (name, v))


def _validation(name, v):
    if v is None or isinstance(v, str):
        return
    if isinstance(v, bool) or not isinstance(v, (int, float)) or not 0 <= float(v) < 1:
        raise ValueError('%s must be a column name or a fraction in [0, 1), got %r' % (name, v))


def _reader_pool(name, v):
    if v not in (None, 'thread', 'process', 'dummy'):
        raise ValueError("%s must be 'thread', 'process' or 'dummy', got %r" % (name, v))


def _mp_start(name, v):
    if v not in (None, 'spawn', 'fork', 'forkserver'):
        raise ValueError("%s must be 'spawn', 'fork' or 'forkserver', got %r" % (name, v))


def _callable_or_none(name, v):
    if v is not None and not callable(v):
        raise ValueError('%s must be callable, got %r' % (name, v))


class ParamsBase:
    """Holds the values; subclasses list their knobs in `PARAMS`."""
    PARAMS = ()
    _table = {}

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        table = {}
        for klass in reversed(cls.__mro__):
            for p in klass.__dict__.get('PARAMS', ()):
                table[p.name] = p
        cls._table = table
        for p in table.values():
            for camel in p.camels:
                if 'set' + camel not in cls.__dict__:
                    setattr(cls, 'set' + camel, _make_setter(p, camel))
                if 'get' + camel not in cls.__dict__:
                    setattr(cls, 'get' + camel, _make_getter(p, camel))

    def __init__(self, **kwargs):
        self._values = {name: copy.copy(p.default) for name, p in self._table.items()}
        self.setParams(**kwargs)

    def setParams(self, **kwargs):
        for k, v in kwargs.items():
            self._set(k, v)
        return self

    def _set(self, name, value):
        p = self._table.get(name)
        if p is None:
            raise TypeError('%s has no parameter %r (known: %s)' % (type(self).__name__, name, ', '.join(sorted(self._table))))
        if p.check is not None:
            p.check(name, value)
        self._values[name] = value

    def _get(self, name):
        return self._values[name]

    def param_dict(self):
        return dict(self._values)

    def copy(self, extra=None):
        """A shallow clone with `extra` applied (what `fit(df, params=...)` trains with)."""
        other = copy.copy(self)
        other._values = dict(self._values)
        if extra:
            other.setParams(**extra)
        return other

    # -- persistence (Spark ML's MLWritable / MLReadable surface) -----------------------------------------------------------
    def write(self):
        from horovod_b200.spark.common.serialization import HorovodParamsWriter
        return HorovodParamsWriter(self)

    def save(self, path):
        """Shortcut for `write().save(path)`."""
        self.write().save(path)

    @classmethod
    def read(cls):
        from horovod_b200.spark.common.serialization import HorovodParamsReader
        return HorovodParamsReader(cls)

    @classmethod
    def load(cls, path):
        """Shortcut for `read().load(path)`."""
        return cls.read().load(path)

    def explainParams(self):
        return '\n'.join('%s: %s (default: %r, current: %r)' % (n, p.doc, p.default, self._values[n])
                         for n, p in sorted(self._table.items()))

    def __getattr__(self, name):
        # plain attribute access to a knob: est.batch_size
        table = type(self)._table
        if name in table and '_values' in self.__dict__:
            return self.__dict__['_values'][name]
        raise AttributeError(name)


def _make_setter(p, camel=None):
    def setter(self, value):
        self._set(p.name, value)
        return self
    setter.__name__ = 'set' + (camel or p.camel)
    setter.__doc__ = p.doc
    return setter


def _make_getter(p, camel=None):
    def getter(self):
        return self._get(p.name)
    getter.__name__ = 'get' + (camel or p.camel)
    getter.__doc__ = p.doc
    return getter


class EstimatorParams(ParamsBase):
    PARAMS = (
        P('num_proc', None, _positive_int, 'number of training processes (default: the backend decides)'),
        P('backend', None, None, 'Backend that runs the training function; excludes num_proc'),
        P('store', None, None, 'Store (or path prefix) for intermediate data, checkpoints and logs'),
        P('model', None, None, 'the model to train'),
        P('optimizer', None, None, 'optimizer (instance built on the model)'),
        P('loss', None, None, 'loss function(s)'),
        P('loss_weights', None, None, 'one weight per loss / output'),
        P('metrics', None, None, 'metric functions reported with the history'),
        P('feature_cols', None, _str_list, 'feature column names'),
        P('label_cols', None, _str_list, 'label column names'),
        P('sample_weight_col', None, None, 'column with per-row loss weights'),
        P('validation', None, _validation, 'validation column name or fraction of rows in [0, 1)'),
        P('callbacks', None, None, 'framework callbacks'),
        P('batch_size', 32, _positive_int, 'rows per step and process'),
        P('val_batch_size', None, _positive_int, 'rows per validation step (default: batch_size)'),
        P('epochs', 1, _positive_int, 'passes over the training data'),
        P('verbose', 1, _non_negative_int, 'verbosity'),
        P('random_seed', 0, None, 'seed for the split and for shuffling'),
        P('shuffle', True, None, 'shuffle the training rows of a rank every epoch'),
        P('shuffle_buffer_size', None, _non_negative_int, 'accepted for compatibility: shards are shuffled in memory',
          camel='ShufflingBufferSize'),
        P('partitions_per_process', 1, _positive_int, 'Parquet files written per training process'),
        P('run_id', None, None, 'name of the run directory in the store; an existing checkpoint there is resumed'),
        P('train_steps_per_epoch', None, _positive_int, 'steps per epoch (default: rows of the smallest shard // batch_size)'),
        P('validation_steps_per_epoch', None, _positive_int, 'validation steps per epoch'),
        P('transformation_fn', None, _callable_or_none, 'fn(dict of column -> tensor/array) -> dict applied to every batch'),
        P('input_shapes', None, None, 'one shape per feature column ([-1, ...]); rows are reshaped before the model sees them'),
        P('label_shapes', None, None, 'one shape per label column'),
        P('inmemory_cache_all', False, None, 'keep the decoded shard in memory across epochs', camel='InMemoryCacheAll'),
        P('use_gpu', True, None, 'train on cuda:<local_rank> when a GPU is visible'),
        P('gradient_compression', None, None, 'hvd.Compression.* for the gradient allreduce'),
        P('backward_passes_per_step', 1, _positive_int, 'local gradient accumulation steps'),
        P('compress_sparse_cols', False, None, 'accepted for compatibility'),
        P('categorical_cols', None, _str_list, 'categorical feature columns, handed to the data module (tabular readers)'),
        P('continuous_cols', None, _str_list, 'continuous feature columns, handed to the data module (tabular readers)'),
        P('data_module', None, None, 'DataModule class that feeds the training loop (default: the framework\'s Parquet module)'),
        P('train_reader_num_workers', None, _non_negative_int,
          'reader threads for training data: >= 1 decodes batches on a background thread (async loader)', camel='TrainReaderNumWorker'),
        P('val_reader_num_workers', None, _non_negative_int, 'same for validation data', camel='ValReaderNumWorker'),
        P('reader_pool_type', 'thread', _reader_pool, "'thread' (the async loaders' pool), 'process' and 'dummy' are accepted; readers here are threads"),
        P('mp_start_method', None, _mp_start, "accepted for compatibility: training processes are fresh interpreters started by the launcher ('spawn' semantics) whatever is set"),
        P('tensorflow_dataset_prefetch_buffer_size', None, _non_negative_int, 'accepted for compatibility (tf.data prefetch depth of the reference\'s Keras reader)'),
        P('transformation_edit_fields', None, None, 'fields transformation_fn adds or retypes: [(name, numpy dtype, shape, nullable)]'),
        P('transformation_removed_fields', None, _str_list, 'columns transformation_fn drops; they are removed from every batch after it ran'),
    )


class ModelParams(ParamsBase):
    PARAMS = (
        P('model', None, None, 'the trained model'),
        P('history', None, None, 'per-epoch metrics of the run that produced the model'),
        P('feature_columns', None, _str_list, 'feature column names'),
        P('label_columns', None, _str_list, 'label column names', camel='LabelColoumns'),
        P('output_cols', None, _str_list, 'names of the prediction columns appended by transform()'),
        P('run_id', None, None, 'run that produced the model'),
        P('metadata', None, None, 'column metadata of the training data'),
        P('batch_size', 1024, _positive_int, 'rows per inference batch in transform()'),
    )
