"""DataFrame -> Parquet shards in the Store, with metadata and a cache of already materialised DataFrames.

Role parity: horovod/spark/common/util.py (`prepare_data` :686-741, `_get_or_create_dataset` :560-660, `check_validation`
:525-541, `get_simple_meta_from_parquet` :400-470, `to_list` / vector handling :150-260) and common/cache.py
(`TrainingDataCache`).  The reference writes Parquet for Petastorm and keeps Petastorm metadata; here the shards are read
back by `horovod_b200.spark.data_loaders` (pyarrow.dataset), the metadata is computed from the Parquet schema and the
first row group, and the input can be a Spark or a pandas DataFrame.
"""
import contextlib
import os
import threading
import uuid
import weakref

import numpy as np


def is_spark_df(df):
    return type(df).__module__.startswith('pyspark.')


def check_columns(df, columns, what):
    have = set(df.columns)
    missing = [c for c in columns if c not in have]
    if missing:
        raise ValueError('%s column(s) %s not found in the DataFrame (columns: %s)' % (what, missing, sorted(have)))


def check_validation(validation, df=None):
    """`validation` is None, a fraction in [0, 1) or the name of a boolean-ish column."""
    if validation is None:
        return
    if isinstance(validation, str):
        if df is not None and validation not in df.columns:
            raise ValueError('Validation column "%s" does not exist in the DataFrame' % validation)
        return
    if isinstance(validation, bool) or not isinstance(validation, (int, float)):
        raise ValueError('Param validation must be of type "float" or "str", found: %s' % type(validation))
    if not 0 <= validation < 1:
        raise ValueError('Validation split %s must be in the range: [0, 1)' % validation)


def split_validation(df, validation, seed):
    """-> (train_df, val_df or None)."""
    if validation is None or validation == 0:
        return df, None
    if isinstance(validation, str):
        if is_spark_df(df):
            flag = df[validation].cast('boolean')
            return df.filter(~flag).drop(validation), df.filter(flag).drop(validation)
        mask = df[validation].astype(bool)
        return df[~mask].drop(columns=[validation]), df[mask].drop(columns=[validation])
    frac = float(validation)
    if is_spark_df(df):
        train, val = df.randomSplit([1 - frac, frac], seed=seed)
        return train, val
    val = df.sample(frac=frac, random_state=seed)
    return df.drop(val.index), val


def _densify(value):
    """Spark ML vectors (DenseVector / SparseVector), numpy arrays and nested lists -> plain nested lists."""
    if hasattr(value, 'toArray'):
        return value.toArray().tolist()
    if isinstance(value, np.ndarray):
        return value.tolist()
    return value


def _pandas_to_table(pdf, columns):
    import pyarrow as pa
    data = {}
    for c in columns:
        col = pdf[c]
        first = col.iloc[0] if len(col) else None
        if hasattr(first, 'toArray') or isinstance(first, (np.ndarray, list, tuple)):
            data[c] = pa.array([_densify(v) for v in col])
        else:
            data[c] = pa.array(col.to_numpy())
    return pa.table(data)


def write_parquet(df, path, store, num_files, columns=None):
    """Writes `df[columns]` as `num_files` Parquet files under `path`; returns the row count."""
    store.delete(path)
    if is_spark_df(df):
        out = df.select(*columns) if columns else df
        vector_cols = [f.name for f in out.schema.fields if type(f.dataType).__name__ == 'VectorUDT']
        if vector_cols:
            from pyspark.ml.functions import vector_to_array
            for c in vector_cols:
                out = out.withColumn(c, vector_to_array(out[c]))
        # a scheme-less path would be resolved against the cluster's default filesystem (often HDFS): LocalStore means file://
        spark_path = path if '://' in path else 'file://' + path
        out.repartition(num_files).write.mode('overwrite').parquet(spark_path)
        return out.count()
    import pyarrow.parquet as pq
    cols = list(columns) if columns else list(df.columns)
    table = _pandas_to_table(df, cols)
    local = store._local(path)
    store.fs.create_dir(local, recursive=True)
    rows = table.num_rows
    per_file = -(-rows // num_files) if rows else 0
    for i in range(num_files):
        if i * per_file < rows:
            pq.write_table(table.slice(i * per_file, per_file), os.path.join(local, 'part-%05d.parquet' % i), filesystem=store.fs)
    return rows


def parquet_metadata(store, path):
    """{'rows': N, 'avg_row_size': bytes, 'columns': {name: {'dtype': numpy dtype name, 'shape': per-row shape or None}}}.
    The per-row shape of a list column comes from the first row (ragged columns report None)."""
    import pyarrow.dataset as ds
    dataset = ds.dataset(store._local(path), format='parquet', filesystem=store.fs)
    rows = dataset.count_rows()
    head = dataset.head(1)
    total_bytes = 0
    for frag in dataset.get_fragments():
        md = frag.metadata
        total_bytes += sum(md.row_group(i).total_byte_size for i in range(md.num_row_groups))
    cols = {}
    for name in head.schema.names:
        values = head.column(name).to_pylist()
        first = values[0] if values else None
        if isinstance(first, (list, tuple)):
            arr = np.asarray(first)
            cols[name] = {'dtype': arr.dtype.name if arr.dtype != object else 'object',
                          'shape': list(arr.shape) if arr.dtype != object else None}
        else:
            np_dtype = head.schema.field(name).type.to_pandas_dtype()
            cols[name] = {'dtype': np.dtype(np_dtype).name if np_dtype is not object else 'object', 'shape': []}
    return {'rows': rows, 'avg_row_size': (total_bytes / rows) if rows else 0, 'columns': cols}


def check_shape_compatibility(metadata, feature_columns, label_columns, input_shapes=None, output_shapes=None, label_shapes=None):
    """Every declared shape must hold exactly the number of elements a row of its column has (-1 = the batch dimension)."""
    def count(shape):
        n = 1
        for d in shape:
            n *= d if d != -1 else 1
        return n

    def check(kind, cols, shapes):
        if shapes is None:
            return
        if len(shapes) != len(cols):
            raise ValueError('%s column count %d must equal the number of %s shapes %d' % (kind, len(cols), kind, len(shapes)))
        for col, shape in zip(cols, shapes):
            row_shape = metadata['columns'][col]['shape']
            if row_shape is None:
                continue
            if count(row_shape) != count(shape):
                raise ValueError('%s column "%s" with %d elements per row does not match the declared shape %s'
                                 % (kind, col, count(row_shape), list(shape)))
    check('feature', feature_columns, input_shapes)
    check('label', label_columns, label_shapes if label_shapes is not None else output_shapes)


class PreparedDataset:
    """What the estimators train on: Parquet paths in the store plus row counts and column metadata."""

    def __init__(self, idx, train_path, val_path, train_rows, val_rows, metadata):
        self.idx, self.train_path, self.val_path = idx, train_path, val_path
        self.train_rows, self.val_rows, self.metadata = train_rows, val_rows, metadata

    def __repr__(self):
        return 'PreparedDataset(idx=%s, train_rows=%d, val_rows=%d)' % (self.idx, self.train_rows, self.val_rows)


from horovod_b200.spark.common.cache import TrainingDataCache as _DatasetCache  # noqa: E402

_dataset_cache = _DatasetCache()


def clear_training_cache(store=None):
    _dataset_cache.clear(store)


@contextlib.contextmanager
def prepare_data(num_processes, store, df, label_columns, feature_columns, validation=None, sample_weight_col=None,
                 partitions_per_process=1, random_seed=0, verbose=0, keep=True):
    """Materialises `df` (split into train / validation) as Parquet in `store`; yields a PreparedDataset.

    keep=True leaves the files in the store and remembers them for the next fit on the same DataFrame object;
    keep=False deletes them when the block ends."""
    check_validation(validation, df)
    columns = list(feature_columns) + list(label_columns) + ([sample_weight_col] if sample_weight_col else [])
    check_columns(df, columns, 'Training')
    num_files = max(1, num_processes * partitions_per_process)
    key, dataset = _dataset_cache.lookup(df, store, validation, columns, num_files)
    if dataset is None:
        idx = uuid.uuid4().hex[:8]
        select = columns + ([validation] if isinstance(validation, str) else [])
        train_df, val_df = split_validation(df[select] if not is_spark_df(df) else df.select(*select), validation, random_seed)
        train_path, val_path = store.get_train_data_path(idx), store.get_val_data_path(idx)
        train_rows = write_parquet(train_df, train_path, store, num_files, columns)
        if train_rows < num_processes:
            store.delete(train_path)
            raise ValueError('%d training rows cannot be spread over %d processes' % (train_rows, num_processes))
        val_rows = 0
        if val_df is not None:
            val_rows = write_parquet(val_df, val_path, store, num_files, columns)
            if val_rows == 0:
                store.delete(val_path)
        dataset = PreparedDataset(idx, train_path, val_path if val_rows else None, train_rows, val_rows,
                                  parquet_metadata(store, train_path))
        if verbose:
            print('prepared %s: %d train rows, %d validation rows, %.0f bytes/row' %
                  (dataset.idx, train_rows, val_rows, dataset.metadata['avg_row_size']), flush=True)
        set_dataset_properties(dataset.idx, (dataset.train_rows, dataset.val_rows, dataset.metadata, dataset.metadata['avg_row_size']))
        if keep:
            _dataset_cache.insert(key, df, dataset)
    try:
        yield dataset
    finally:
        if keep:
            _dataset_cache.release(key)
        else:
            store.delete(dataset.train_path)
            if dataset.val_path:
                store.delete(dataset.val_path)


def make_transform(transformation_fn, removed_fields=None):
    """The per-batch transform of an estimator: `transformation_fn` (if any), then the columns listed in
    `transformation_removed_fields` are dropped from the batch."""
    removed = set(removed_fields or ())
    if not removed:
        return transformation_fn

    def transform(batch):
        out = transformation_fn(batch) if transformation_fn else batch
        return {k: v for k, v in out.items() if k not in removed}
    return transform


def existing_dataset(store, dataset_idx=None):
    """PreparedDataset for Parquet that is already in the store (`fit_on_parquet`)."""
    train_path = store.get_train_data_path(dataset_idx)
    if not store.exists(train_path) or not store.is_parquet_dataset(train_path):
        raise ValueError('No Parquet training data at %s' % train_path)
    val_path = store.get_val_data_path(dataset_idx)
    has_val = store.exists(val_path) and store.is_parquet_dataset(val_path)
    meta = parquet_metadata(store, train_path)
    val_rows = parquet_metadata(store, val_path)['rows'] if has_val else 0
    return PreparedDataset(dataset_idx, train_path, val_path if has_val else None, meta['rows'], val_rows, meta)


# ---- small helpers of the reference's util module that user code and the estimators' callers rely on --------------------------
def to_list(var, length):
    """None -> None; a scalar or a one-element list -> that element `length` times; a list of `length` entries -> itself."""
    if var is None:
        return None
    var = list(var) if isinstance(var, (list, tuple)) else [var]
    if len(var) == 1:
        return var * length
    if len(var) != length:
        raise ValueError('List must have %d entries (or one), found %d' % (length, len(var)))
    return var


def numpy_type_to_str(dtype):
    """np.float32 / dtype('float32') / 'float32' -> 'float32' (what the metadata dictionaries store)."""
    return np.dtype(dtype).name


_SPARK_TO_NUMPY = {'BooleanType': np.bool_, 'ByteType': np.int8, 'ShortType': np.int16, 'IntegerType': np.int32, 'LongType': np.int64,
                   'FloatType': np.float32, 'DoubleType': np.float64, 'StringType': np.str_, 'BinaryType': np.bytes_}


def data_type_to_str(dtype):
    """Spark SQL type (instance or class) -> its name without the 'Type' suffix: IntegerType() -> 'Integer'; vector and
    array columns report 'Vector' / 'Array' (reference util.py `data_type_to_str`)."""
    name = dtype if isinstance(dtype, str) else getattr(dtype, '__name__', None) or type(dtype).__name__
    if name in ('VectorUDT', 'DenseVector', 'SparseVector'):
        return 'Vector'
    return name[:-4] if name.endswith('Type') else name


def data_type_to_numpy(dtype):
    """Spark SQL scalar type -> numpy scalar type; vectors and arrays of floats -> float64 / float32 element type."""
    name = dtype if isinstance(dtype, str) else getattr(dtype, '__name__', None) or type(dtype).__name__
    if name in ('VectorUDT', 'DenseVector', 'SparseVector', 'Vector'):
        return np.float64
    if name == 'ArrayType':
        return data_type_to_numpy(getattr(dtype, 'elementType', 'DoubleType'))
    key = name if name.endswith('Type') else name + 'Type'
    if key not in _SPARK_TO_NUMPY:
        raise ValueError('Unrecognized data type: %s' % name)
    return _SPARK_TO_NUMPY[key]


def spark_scalar_to_python_type(dtype):
    """Spark SQL scalar type -> the Python type of a collected value."""
    np_type = data_type_to_numpy(dtype)
    if np_type is np.bool_:
        return bool
    if np_type is np.str_:
        return str
    if np_type is np.bytes_:
        return bytes
    return float if np.issubdtype(np_type, np.floating) else int


def pyarrow_to_spark_data_type(dtype):
    """pyarrow DataType -> the NAME of the Spark SQL type that holds it ('IntegerType', 'ArrayType(FloatType)', ...); returns
    the pyspark class instance instead when pyspark is importable."""
    import pyarrow.types as pat
    table = ((pat.is_boolean, 'BooleanType'), (pat.is_int8, 'ByteType'), (pat.is_int16, 'ShortType'), (pat.is_int32, 'IntegerType'),
             (pat.is_int64, 'LongType'), (pat.is_float32, 'FloatType'), (pat.is_float64, 'DoubleType'), (pat.is_string, 'StringType'),
             (pat.is_large_string, 'StringType'), (pat.is_binary, 'BinaryType'), (pat.is_large_binary, 'BinaryType'))
    if pat.is_list(dtype) or pat.is_large_list(dtype) or pat.is_fixed_size_list(dtype):
        inner = pyarrow_to_spark_data_type(dtype.value_type)
        try:
            from pyspark.sql.types import ArrayType
            return ArrayType(inner)
        except ImportError:
            return 'ArrayType(%s)' % inner
    for pred, name in table:
        if pred(dtype):
            try:
                import pyspark.sql.types as st
                return getattr(st, name)()
            except ImportError:
                return name
    raise ValueError('Unrecognized pyarrow data type: %s' % dtype)


def get_simple_meta_from_parquet(store, label_columns, feature_columns, sample_weight_col=None, dataset_idx=None):
    """(train_rows, val_rows, metadata, avg_row_size) of the Parquet the store already holds for `dataset_idx`; raises when a
    needed column is missing or the training set is empty (reference util.py :444-500)."""
    train_path = store.get_train_data_path(dataset_idx)
    if not store.exists(train_path) or not store.is_parquet_dataset(train_path):
        raise ValueError('{} is not a parquet dataset'.format(train_path))
    meta = parquet_metadata(store, train_path)
    if meta['rows'] == 0:
        raise ValueError('Training data is empty: {}'.format(train_path))
    needed = list(label_columns) + list(feature_columns) + ([sample_weight_col] if sample_weight_col else [])
    missing = [c for c in needed if c not in meta['columns']]
    if missing:
        raise ValueError('Column(s) %s not found in %s (columns: %s)' % (missing, train_path, sorted(meta['columns'])))
    val_path = store.get_val_data_path(dataset_idx)
    val_rows = parquet_metadata(store, val_path)['rows'] if store.exists(val_path) and store.is_parquet_dataset(val_path) else 0
    metadata = {c: {'spark_data_type': None, 'is_sparse_vector_only': False, 'shape': info['shape'], 'intermediate_format':
                    'array' if info['shape'] else 'nochange', 'max_size': int(np.prod(info['shape'])) if info['shape'] else 1,
                    'dtype': info['dtype']} for c, info in meta['columns'].items()}
    return meta['rows'], val_rows, metadata, meta['avg_row_size']


_dataset_properties = {}


def get_dataset_properties(dataset_idx):
    """(train_rows, val_rows, metadata, avg_row_size) remembered for a dataset prepared in this process."""
    return _dataset_properties[dataset_idx]


def set_dataset_properties(dataset_idx, props):
    _dataset_properties[dataset_idx] = props


def get_available_devices():
    """GPU addresses Spark assigned to THIS task (`spark.task.resource.gpu.amount`, Spark >= 3); [] outside a Spark task or when
    the job requested no GPU resources (reference util.py `get_available_devices`)."""
    try:
        from pyspark import TaskContext
    except ImportError:
        return []
    ctx = TaskContext.get()
    if ctx is None or not hasattr(ctx, 'resources'):
        return []
    gpu = ctx.resources().get('gpu')
    return list(gpu.addresses) if gpu is not None else []


def gpu_index_for(local_rank, devices=None, environ=None):
    """CUDA device index of a training process: the GPU Spark assigned to the task when there is one, else `local_rank`.
    HOROVOD_SPARK_USE_LOCAL_RANK_GPU_INDEX=1 forces `local_rank` (clusters whose GPU addresses are not CUDA ordinals)."""
    environ = os.environ if environ is None else environ
    if devices is None:
        from horovod_b200.spark.task import get_available_devices as task_devices      # explicitly recorded resources win
        devices = task_devices()
    if devices and environ.get('HOROVOD_SPARK_USE_LOCAL_RANK_GPU_INDEX', '0') == '0':
        return int(devices[0])
    return local_rank
