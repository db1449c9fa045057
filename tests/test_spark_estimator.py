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
