"""TorchEstimator on a pandas DataFrame with the LocalBackend (2 processes) and a LocalStore.
Reference coverage model: test/integration/test_spark_torch.py::test_fit_model."""
import numpy as np
import pandas as pd
import pytest
import torch

from horovod_b200.spark.common import LocalBackend, LocalStore, Store
from horovod_b200.spark.torch import TorchEstimator


def test_store_layout_and_io(tmp_path):
    st = Store.create(str(tmp_path / 'store'))
    assert isinstance(st, LocalStore)
    assert st.get_train_data_path().endswith('intermediate_train_data') and st.get_val_data_path(3).endswith('intermediate_val_data.3')
    ck = st.get_checkpoint_path('r1')
    assert ck.endswith('runs/r1/checkpoint.pt') and st.get_logs_path('r1').endswith('runs/r1/logs')
    assert not st.exists(ck)
    st.write(ck, b'abc')
    assert st.exists(ck) and st.read(ck) == b'abc'
    st.write_text(st.get_logs_path('r1') + '/note.txt', 'hello')
    assert st.read(st.get_logs_path('r1') + '/note.txt') == b'hello'
    assert Store.create('file://' + str(tmp_path / 's2')).prefix_path == str(tmp_path / 's2')
    assert LocalStore(str(tmp_path / 's3'), save_runs=False).get_checkpoint_path('x') is None


def test_store_remote_snapshot_sync_and_paths(tmp_path):
    """The rest of the reference's Store interface: to_remote snapshot, local scratch dir + sync, checkpoint listing, URI
    helpers, scheme matching."""
    import os
    import pickle
    from horovod_b200.spark.common.store import DBFSLocalStore, FilesystemStore, HDFSStore, split_protocol
    st = LocalStore(str(tmp_path / 'store'))
    remote = st.to_remote('run7', 2)
    assert remote.train_data_path == st.get_train_data_path(2) and remote.checkpoint_path == st.get_checkpoint_path('run7')
    assert remote.saving_runs and remote.logs_subdir == 'logs' and remote.checkpoint_filename == 'checkpoint.pt'
    assert remote.runs_path == st.get_runs_path() and remote.run_path == st.get_run_path('run7')
    with remote.get_local_output_dir() as d:
        os.makedirs(os.path.join(d, 'logs'))
        for name in ('epoch=0.ckpt', 'epoch=1.ckpt', os.path.join('logs', 'events.txt')):
            with open(os.path.join(d, name), 'w') as f:
                f.write(name)
        remote.sync(d)
        scratch = d
    assert not os.path.exists(scratch)                                   # the scratch directory is removed on exit
    assert [os.path.basename(p) for p in st.get_checkpoints('run7')] == ['epoch=0.ckpt', 'epoch=1.ckpt']
    assert st.get_checkpoints('run7', suffix='.txt')[0].endswith('logs/events.txt') and st.get_checkpoints('nope') == []
    assert st.read(st.get_run_path('run7') + '/logs/events.txt') == b'logs/events.txt'
    st.copy(__file__, st.get_run_path('run7') + '/copied.py')
    assert st.exists(st.get_run_path('run7') + '/copied.py')
    with pytest.raises(IsADirectoryError):
        st.copy(str(tmp_path), st.get_run_path('run7') + '/dir')
    assert st.get_full_path('/a/b') == 'file:///a/b' and st.get_full_path_fn()('s3://x/y') == 's3://x/y'
    assert st.get_localized_path('file:///a/b') == '/a/b' and st.get_data_metadata_path('/a/b/') == '/a/b/_metadata.json'
    assert split_protocol('s3://bucket/k') == ('s3', 'bucket/k') and split_protocol('/plain') == (None, '/plain')
    assert LocalStore.matches('/x') and LocalStore.matches('file:///x') and not LocalStore.matches('s3://b/x')
    assert HDFSStore.matches('hdfs://nn:8020/x') and not HDFSStore.matches('/x') and FilesystemStore.matches('gs://b/x')
    assert DBFSLocalStore.matches_dbfs('dbfs:/ml/x') and DBFSLocalStore.normalize_path('dbfs:/ml/x') == '/dbfs/ml/x'
    h = HDFSStore('hdfs://nn:8020/user/me')
    assert h.get_full_path('/user/me/runs') == 'hdfs://nn:8020/user/me/runs' and h.get_full_path_fn()('/d') == 'hdfs://nn:8020/d'
    assert h.get_localized_path('hdfs://nn:8020/user/me/x') == '/user/me/x'
    from horovod_b200.spark.common import util
    import pandas as pd
    util.write_parquet(pd.DataFrame({'a': [1, 2, 3]}), st.get_train_data_path(), st, 1, ['a'])
    assert st.is_parquet_dataset(st.get_train_data_path()) and st.get_parquet_dataset(st.get_train_data_path()).read().num_rows == 3
    nosave = LocalStore(str(tmp_path / 'n'), save_runs=False).to_remote('r', None)
    assert not nosave.saving_runs and nosave.checkpoint_path is None


def test_estimator_param_validation(tmp_path):
    m = torch.nn.Linear(2, 1)
    o = torch.optim.SGD(m.parameters(), lr=0.1)
    with pytest.raises(ValueError):
        TorchEstimator(model=m, optimizer=o, loss=torch.nn.functional.mse_loss, feature_cols=['x'], label_cols=['y'])
    with pytest.raises(ValueError):
        TorchEstimator(model=m, optimizer=o, loss=None, feature_cols=['x'], label_cols=['y'], store=str(tmp_path))
    with pytest.raises(ValueError):
        TorchEstimator(model=m, optimizer=o, loss=torch.nn.functional.mse_loss, feature_cols=['x'], label_cols=['y'],
                       store=str(tmp_path), backend=LocalBackend(1), num_proc=2)


def test_fit_linear_regression_two_procs(native_built, tmp_path):
    rng = np.random.RandomState(0)
    x = rng.randn(512, 3).astype(np.float32)
    w = np.array([1.5, -2.0, 0.5], dtype=np.float32)
    y = x @ w + 0.25
    df = pd.DataFrame({'features': list(x), 'label': y})
    torch.manual_seed(0)
    model = torch.nn.Linear(3, 1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.5)
    est = TorchEstimator(model=model, optimizer=opt, loss=torch.nn.functional.mse_loss, feature_cols=['features'],
                         label_cols=['label'], batch_size=32, epochs=6, validation=0.2, store=str(tmp_path / 'store'),
                         backend=LocalBackend(2), use_gpu=False, verbose=0, run_id='t1')
    tm = est.fit(df)
    hist = tm.getHistory()
    assert len(hist) == 6 and hist[-1]['loss'] < 0.05 * hist[0]['loss'] and hist[-1]['val_loss'] < 0.05
    np.testing.assert_allclose(tm.getModel().weight.detach().numpy().ravel(), w, atol=0.1)
    out = tm.transform(df.head(8))
    assert 'label__output' in out.columns
    np.testing.assert_allclose(np.array(out['label__output'].tolist()), y[:8], atol=0.3)
    st = est.store
    assert st.exists(st.get_checkpoint_path('t1'))
    ck = torch.load(__import__('io').BytesIO(st.read(st.get_checkpoint_path('t1'))), weights_only=False)
    assert ck['epoch'] == 5 and 'model' in ck


def test_store_selection_and_path_rules(tmp_path):
    from horovod_b200.spark.common import DBFSLocalStore, FilesystemStore, HDFSStore
    assert isinstance(Store.create('hdfs://namenode:8020/user/x'), HDFSStore)
    assert isinstance(Store.create('dbfs:/ml/run'), DBFSLocalStore) and isinstance(Store.create('/dbfs/ml/run'), DBFSLocalStore)
    assert type(Store.create('s3://bucket/prefix')) is FilesystemStore and isinstance(Store.create(str(tmp_path)), LocalStore)
    assert HDFSStore.parse_url('hdfs://nn:8020/a/b') == ('nn', 8020, '/a/b')
    assert HDFSStore.parse_url('hdfs:///a/b') == ('default', 0, '/a/b') and HDFSStore.parse_url('hdfs://nn/a') == ('nn', 0, '/a')
    h = HDFSStore('hdfs://nn:8020/base', user='me')
    assert h.get_train_data_path(2) == 'hdfs://nn:8020/base/intermediate_train_data.2' and h._local(h.get_run_path('r')) == '/base/runs/r'
    for given, want in (('dbfs:/a/b', '/dbfs/a/b'), ('dbfs:///a/b', '/dbfs/a/b'), ('/dbfs/a/b', '/dbfs/a/b'), ('/other', '/other')):
        assert DBFSLocalStore.normalize_path(given) == want
    d = DBFSLocalStore('dbfs:/ml')
    assert d.prefix_path == '/dbfs/ml' and d.get_checkpoint_path('r').endswith('runs/r/checkpoint.tf')


def _make_counting_module():
    from horovod_b200.spark.torch.datamodule import ParquetDataModule

    class CountingDataModule(ParquetDataModule):
        short_name = 'counting'

        def train_data(self):
            import os
            with open(os.path.join(self.extra_dir, 'rank%d.txt' % self.cur_shard), 'w') as f:
                f.write('%d/%d %s' % (self.cur_shard, self.shard_count, ','.join(self.schema_fields)))
            return super().train_data()
    return CountingDataModule


def test_fit_with_loss_constructors_async_readers_removed_fields_and_data_module(native_built, tmp_path):
    """The estimator knobs of the reference that change what the training function does: loss built on the workers, batches
    decoded on a reader thread, columns dropped after the transformation, a user-supplied DataModule class."""
    rng = np.random.RandomState(1)
    x = rng.randn(256, 2).astype(np.float32)
    y = x @ np.array([2.0, -1.0], dtype=np.float32)
    df = pd.DataFrame({'features': list(x), 'noise': rng.randn(256).astype(np.float32), 'label': y})
    torch.manual_seed(0)
    model = torch.nn.Linear(2, 1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    cls = _make_counting_module()
    cls.extra_dir = str(tmp_path)

    def drop_check(batch):
        return dict(batch, scratch=batch['label'] * 0)     # a column the transformation adds and the estimator drops again
    est = TorchEstimator(model=model, optimizer=opt, loss_constructors=[lambda: torch.nn.MSELoss()], feature_cols=['features'],
                         label_cols=['label'], batch_size=32, epochs=4, validation=0.25, store=str(tmp_path / 'store'),
                         backend=LocalBackend(2), use_gpu=False, verbose=0, data_module=cls, transformation_fn=drop_check)
    est.setTrainReaderNumWorker(2).setValReaderNumWorker(1).setTransformationRemovedFields(['scratch'])
    est.setInMemoryCacheAll(True).setShufflingBufferSize(0)
    assert est.getTrainReaderNumWorker() == 2 and est.getInMemoryCacheAll() is True and est.getLossConstructors()
    with pytest.raises(ValueError):
        est.setReaderPoolType('fibers')
    tm = est.fit(df)
    hist = tm.getHistory()
    assert len(hist) == 4 and hist[-1]['loss'] < 0.1 * hist[0]['loss'] and 'val_loss' in hist[-1]
    assert sorted(f for f in __import__('os').listdir(tmp_path) if f.startswith('rank')) == ['rank0.txt', 'rank1.txt']
    assert (tmp_path / 'rank1.txt').read_text() == '1/2 features,label'
    assert tm.getLossConstructors() and isinstance(tm.getOptimizer(), torch.optim.SGD) and tm.getLoss() is None
    with pytest.raises(ValueError, match='loss'):
        TorchEstimator(model=model, optimizer=opt, feature_cols=['features'], label_cols=['label'], store=str(tmp_path / 's'))


def test_make_transform_drops_removed_fields():
    from horovod_b200.spark.common.util import make_transform
    assert make_transform(None, None) is None
    f = make_transform(lambda b: dict(b, extra=1), ['a'])
    assert f({'a': 1, 'b': 2}) == {'b': 2, 'extra': 1}
    assert make_transform(None, ['a'])({'a': 1, 'b': 2}) == {'b': 2}


def test_estimator_and_model_save_load_round_trip(native_built, tmp_path):
    """write().save(path) / load(path) (Spark ML's MLWritable / MLReadable surface; reference test_spark_torch.py
    ::test_torch_direct_parquet_train + serialization tests): a trained TorchModel comes back with the same weights and
    predictions, an estimator with its knobs, losses and store."""
    import json
    from horovod_b200.spark.common.serialization import HorovodParamsReader
    from horovod_b200.spark.torch import TorchModel
    rng = np.random.RandomState(3)
    x = rng.randn(128, 2).astype(np.float32)
    df = pd.DataFrame({'features': list(x), 'label': x @ np.array([1.0, -1.0], dtype=np.float32)})
    torch.manual_seed(3)
    model = torch.nn.Linear(2, 1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    est = TorchEstimator(model=model, optimizer=opt, loss=torch.nn.functional.mse_loss, feature_cols=['features'], label_cols=['label'],
                         batch_size=16, epochs=3, store=str(tmp_path / 'store'), backend=LocalBackend(1), use_gpu=False, verbose=0)
    est_path = str(tmp_path / 'saved_estimator')
    est.save(est_path)
    with pytest.raises(IOError, match='overwrite'):
        est.write().save(est_path)
    est.setEpochs(2).write().overwrite().save(est_path)
    meta = json.load(open(est_path + '/metadata/part-00000'))
    assert meta['class'].endswith('TorchEstimator') and meta['paramMap']['epochs'] == 2 and '__torch__' in meta['paramMap']['model']
    est2 = TorchEstimator.load(est_path)
    assert est2.getEpochs() == 2 and est2.getFeatureCols() == ['features'] and isinstance(est2.getStore(), LocalStore)
    assert isinstance(est2.getOptimizer(), torch.optim.SGD) and est2.getLoss() is torch.nn.functional.mse_loss
    assert torch.equal(est2.getModel().weight, model.weight)
    tm = est2.fit(df)                                                # the loaded estimator trains
    assert len(tm.getHistory()) == 2
    model_path = str(tmp_path / 'saved_model')
    tm.write().save(model_path)
    tm2 = TorchModel.load(model_path)
    assert tm2.getHistory() == tm.getHistory() and tm2.getRunId() == tm.getRunId() and tm2.getOutputCols() == ['label__output']
    np.testing.assert_allclose(np.array(tm2.transform(df.head(6))['label__output'].tolist()),
                               np.array(tm.transform(df.head(6))['label__output'].tolist()))
    assert isinstance(HorovodParamsReader().load(model_path), TorchModel)          # class resolved from the metadata
    with pytest.raises(TypeError):
        TorchEstimator.load(model_path)
    # through a Store's filesystem
    st = est.getStore()
    tm.write().option('store', st).save(st.get_run_path(tm.getRunId()) + '/model')
    assert st.exists(st.get_run_path(tm.getRunId()) + '/model/metadata/part-00000')
    assert TorchModel.read().option('store', st).load(st.get_run_path(tm.getRunId()) + '/model').getRunId() == tm.getRunId()
