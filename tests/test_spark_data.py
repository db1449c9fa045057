"""Estimator data path: DataFrame -> Parquet (split, metadata, cache), shard dealing, torch loaders.
Reference coverage model: test/integration/test_spark.py (prepare_data / get_simple_meta / check_shape_compatibility) and
test/single/data/test_data_loader*.py."""
import numpy as np
import pandas as pd
import pytest

from horovod_b200.spark.common import LocalStore, util
from horovod_b200.spark.data_loaders import (ParquetShard, PytorchAsyncDataLoader, PytorchDataLoader, PytorchInfiniteDataLoader,
                                             PytorchInmemDataLoader, shard_files)


def _frame(n=100):
    rng = np.random.RandomState(0)
    return pd.DataFrame({'id': np.arange(n), 'vec': list(rng.randn(n, 4).astype(np.float32)), 'img': [np.arange(6.0) + i for i in range(n)],
                         'label': (np.arange(n) % 3).astype(np.int64), 'is_val': np.arange(n) % 5 == 0})


def test_shard_files_deal():
    files = ['b', 'a', 'd', 'c', 'e']
    assert shard_files(files, 0, 2) == ['a', 'c', 'e'] and shard_files(files, 1, 2) == ['b', 'd']
    assert shard_files(['x'], 3, 4) == ['x']                       # more ranks than files: share
    every = sum((shard_files(files, r, 3) for r in range(3)), [])
    assert sorted(every) == sorted(files)


def test_prepare_data_split_metadata_and_cache(tmp_path):
    store = LocalStore(str(tmp_path))
    df = _frame()
    with util.prepare_data(2, store, df, ['label'], ['vec', 'img'], validation='is_val', keep=True) as ds:
        assert ds.train_rows == 80 and ds.val_rows == 20
        meta = ds.metadata
        assert meta['rows'] == 80 and meta['avg_row_size'] > 0
        assert meta['columns']['vec']['shape'] == [4] and meta['columns']['label']['shape'] == [] and 'int' in meta['columns']['label']['dtype']
        first_idx = ds.idx
        assert store.is_parquet_dataset(ds.train_path)
    with util.prepare_data(2, store, df, ['label'], ['vec', 'img'], validation='is_val', keep=True) as again:
        assert again.idx == first_idx                                    # same DataFrame object: written once
    with util.prepare_data(2, store, df, ['label'], ['vec'], validation=0.25, random_seed=1, keep=False) as other:
        assert other.idx != first_idx and other.train_rows + other.val_rows == 100 and 15 <= other.val_rows <= 35
        path = other.train_path
    assert not store.exists(path)                                        # keep=False cleans up
    util.clear_training_cache(store)
    assert not store.exists(store.get_train_data_path(first_idx))
    with pytest.raises(ValueError, match='not found'):
        with util.prepare_data(1, store, df, ['nope'], ['vec']):
            pass
    with pytest.raises(ValueError, match='cannot be spread'):
        with util.prepare_data(200, store, df, ['label'], ['vec']):
            pass
    for bad in (1.0, -0.1, True, 'missing_col'):
        with pytest.raises(ValueError):
            util.check_validation(bad, df)


def test_shape_compatibility():
    meta = {'columns': {'img': {'shape': [6]}, 'label': {'shape': []}}}
    util.check_shape_compatibility(meta, ['img'], ['label'], input_shapes=[[-1, 2, 3]], label_shapes=[[-1]])
    with pytest.raises(ValueError, match='does not match'):
        util.check_shape_compatibility(meta, ['img'], ['label'], input_shapes=[[-1, 4, 2]])
    with pytest.raises(ValueError, match='column count'):
        util.check_shape_compatibility(meta, ['img'], ['label'], input_shapes=[[-1, 6], [-1, 1]])


def _shards(tmp_path, size, n=96):
    store = LocalStore(str(tmp_path))
    df = _frame(n)
    util.write_parquet(df, store.get_train_data_path(), store, 4, ['id', 'vec', 'img', 'label'])
    return [ParquetShard(store, store.get_train_data_path(), ['id', 'img', 'label'], r, size, {'img': [-1, 2, 3]}) for r in range(size)]


def test_parquet_shard_partition_and_reshape(tmp_path):
    shards = _shards(tmp_path, 3)
    ids = [set(s.load()['id'].tolist()) for s in shards]
    assert sum(len(i) for i in ids) == 96 and set.union(*ids) == set(range(96))
    assert shards[0].load()['img'].shape[1:] == (2, 3) and shards[0].load()['img'].dtype == np.float32
    assert len({s.steps(8) for s in shards}) == 1                       # identical step count on every rank
    assert shards[0].total_rows == 96 and shards[0].min_rows_per_rank == 24


def test_torch_loaders(tmp_path):
    import torch
    shard = _shards(tmp_path, 2)[0]                                       # 48 rows
    loader = PytorchDataLoader(shard, batch_size=8, shuffle=True, seed=5)
    e1 = [b['id'].tolist() for b in loader]
    e2 = [b['id'].tolist() for b in loader]
    assert len(loader) == 6 and len(e1) == 6 and sorted(sum(e1, [])) == sorted(shard.load()['id'].tolist()) and e1 != e2
    assert [b['id'].tolist() for b in PytorchDataLoader(shard, batch_size=8, shuffle=True, seed=5)] == e1   # seeded
    plain = PytorchDataLoader(shard, batch_size=8, shuffle=False, transformation_fn=lambda b: {**b, 'twice': b['id'] * 2})
    b0 = next(iter(plain))
    assert torch.equal(b0['twice'], b0['id'] * 2) and b0['img'].shape == (8, 2, 3)
    inf = PytorchInfiniteDataLoader(shard, batch_size=10, shuffle=False, steps=3)
    a = [b['id'].tolist() for b in inf]
    b = [b['id'].tolist() for b in inf]
    assert len(a) == 3 and a[0] != b[0]                                   # the second epoch continues, it does not restart
    inmem = PytorchInmemDataLoader(shard, batch_size=8, shuffle=True, seed=1)
    assert sorted(sum([x['id'].tolist() for x in inmem], [])) == sorted(shard.load()['id'].tolist())
    asy = PytorchAsyncDataLoader(shard, batch_size=8, shuffle=False, async_loader_queue_size=2)
    try:
        assert [x['id'].tolist() for x in asy] == [x['id'].tolist() for x in PytorchDataLoader(shard, batch_size=8, shuffle=False)]
    finally:
        asy.close_async_loader()


def test_reference_util_helpers(tmp_path):
    import pyarrow as pa
    assert util.to_list(None, 3) is None and util.to_list(1, 3) == [1, 1, 1] and util.to_list([1, 2], 2) == [1, 2]
    with pytest.raises(ValueError):
        util.to_list([1, 2], 3)
    assert util.numpy_type_to_str(np.float32) == 'float32' and util.numpy_type_to_str('int64') == 'int64'
    assert util.data_type_to_str('IntegerType') == 'Integer' and util.data_type_to_str('VectorUDT') == 'Vector'
    assert util.data_type_to_numpy('FloatType') is np.float32 and util.data_type_to_numpy('Long') is np.int64
    assert util.spark_scalar_to_python_type('DoubleType') is float and util.spark_scalar_to_python_type('ShortType') is int
    assert util.spark_scalar_to_python_type('BooleanType') is bool and util.spark_scalar_to_python_type('StringType') is str
    with pytest.raises(ValueError):
        util.data_type_to_numpy('MapType')
    assert util.pyarrow_to_spark_data_type(pa.int32()) == 'IntegerType'
    assert util.pyarrow_to_spark_data_type(pa.list_(pa.float32())) == 'ArrayType(FloatType)'
    store = LocalStore(str(tmp_path))
    with pytest.raises(ValueError, match='not a parquet dataset'):
        util.get_simple_meta_from_parquet(store, ['label'], ['vec'])
    df = _frame(40)
    with util.prepare_data(2, store, df, ['label'], ['vec'], validation=0.25, keep=True) as ds:
        tr, va, meta, avg = util.get_simple_meta_from_parquet(store, ['label'], ['vec'], dataset_idx=ds.idx)
        assert tr == ds.train_rows and va == ds.val_rows and tr + va == 40 and avg > 0
        assert meta['vec']['intermediate_format'] == 'array' and meta['label']['intermediate_format'] == 'nochange'
        assert meta['vec']['max_size'] == int(np.prod(meta['vec']['shape']))
        assert util.get_dataset_properties(ds.idx)[0] == tr
        with pytest.raises(ValueError, match='nope'):
            util.get_simple_meta_from_parquet(store, ['nope'], ['vec'], dataset_idx=ds.idx)


def test_gpu_index_for_spark_task_resources():
    assert util.get_available_devices() == []                       # no pyspark / no task context here
    assert util.gpu_index_for(3, devices=[], environ={}) == 3
    assert util.gpu_index_for(3, devices=['5', '6'], environ={}) == 5
    assert util.gpu_index_for(3, devices=['5'], environ={'HOROVOD_SPARK_USE_LOCAL_RANK_GPU_INDEX': '1'}) == 3


def test_task_info_resources():
    from horovod_b200.spark import task
    assert task.get_available_devices() == []

    class Res:
        addresses = ['3', '4']
    try:
        task.set_resources({'gpu': Res()})
        assert task.get_available_devices() == ['3', '4'] and util.gpu_index_for(0, environ={}) == 3
        task.set_resources({'gpu': ['7']})
        assert task.get_available_devices() == ['7'] and util.gpu_index_for(1, environ={'HOROVOD_SPARK_USE_LOCAL_RANK_GPU_INDEX': '1'}) == 1
    finally:
        task.set_resources({})
    assert util.gpu_index_for(2, environ={}) == 2
