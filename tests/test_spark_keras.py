"""KerasEstimator control flow on a numpy stand-in for TensorFlow/Keras (tests/fakes): model / optimizer transport,
DistributedOptimizer wrapping, broadcast + metric-average callbacks, store checkpoints and resume, weights hand-back,
transform.  Reference coverage model: test/integration/test_spark_keras.py (fit_model, restore_from_checkpoint,
keras_direct_parquet_train, serialization round trips)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SCRIPT = r'''
import sys, numpy as np, pandas as pd
import tensorflow as tf                      # tests/fakes/tensorflow
from tensorflow.mini_keras import LinearModel, SGD
from horovod_b200.spark.common import LocalBackend
from horovod_b200.spark.keras import KerasEstimator, KerasModel
from horovod_b200.spark.keras import util as kutil

# transport round trips
m = LinearModel(3, 1, seed=3)
m2 = kutil.deserialize_model(kutil.serialize_model(m))
assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), m2.get_weights()))
o2 = kutil.deserialize_optimizer(kutil.serialize_optimizer(SGD(0.25)))
assert isinstance(o2, SGD) and abs(o2.get_config()['learning_rate'] - 0.25) < 1e-7
ws = kutil.weights_from_bytes(kutil.weights_to_bytes(m.get_weights()))
assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), ws))

rng = np.random.RandomState(0)
x = rng.randn(512, 3).astype(np.float32)
w = np.array([1.5, -2.0, 0.5], np.float32)
df = pd.DataFrame({'features': list(x), 'label': x @ w + 0.25})
store = sys.argv[1]
env = {'PYTHONPATH': sys.argv[2]}
est = KerasEstimator(model=LinearModel(3, 1), optimizer=SGD(0.1), loss='mse', metrics=['mae'], feature_cols=['features'],
                     label_cols=['label'], batch_size=32, epochs=5, validation=0.25, store=store, backend=LocalBackend(2, env=env),
                     verbose=0, run_id='k1')
for bad in (dict(optimizer=None), dict(loss=None), dict(model=object())):
    try:
        KerasEstimator(**{**dict(model=LinearModel(3), optimizer=SGD(), loss='mse', feature_cols=['f'], label_cols=['l'], store=store), **bad})
        raise SystemExit('expected ValueError for %r' % bad)
    except ValueError:
        pass
km = est.fit(df)
h = km.getHistory()
assert len(h['loss']) == 5 and h['loss'][-1] < 0.05 * h['loss'][0] and h['val_loss'][-1] < 0.05 and 'val_mae' in h, h
np.testing.assert_allclose(km.getModel().get_weights()[0].ravel(), w, atol=0.1)
out = km.transform(df.head(6))
np.testing.assert_allclose(np.array(out['label__output'].tolist()), df['label'].values[:6], atol=0.3)
assert est.store.exists(est.store.get_checkpoint_path('k1'))
more = est.fit(df, params={'epochs': 7})              # resumes after epoch 4
assert len(more.getHistory()['loss']) == 2, more.getHistory()

# the run's checkpoint -> a serialized model (Store.read_serialized_keras_model), optimizer transport, backend_env, data module
blob = est.store.read_serialized_keras_model(est.store.get_checkpoint_path('k1'), LinearModel(3, 1), None)
np.testing.assert_allclose(kutil.deserialize_model(blob).get_weights()[0].ravel(), w, atol=0.1)
from horovod_b200.spark.keras import optimizer as kopt, datamodule as kdm, remote as kremote
o3 = kopt.deserialize_tf_keras_optimizer(kopt.serialize_tf_keras_optimizer(SGD(0.5)), model=m)
assert isinstance(o3, SGD) and abs(o3.get_config()['learning_rate'] - 0.5) < 1e-7 and kopt.is_string(kopt.serialize_bare_keras_optimizer(SGD()))
assert kdm.PetastormDataModule is kdm.ParquetDataModule and callable(kremote.RemoteTrainer({}))
import os
marker = os.path.join(store, 'env_seen')


class EnvCheck:                                    # a callback that records what the training process sees
    def __init__(self, path):
        self.path = path

    def set_model(self, model):
        pass

    def on_epoch_end(self, epoch, logs=None):
        import horovod_b200.tensorflow.keras as hvd
        if hvd.rank() == 1:
            open(self.path, 'w').write(os.environ.get('HVD_TEST_KERAS_BACKEND', 'missing'))

    def __getattr__(self, name):
        if name.startswith('on_'):
            return lambda *a, **k: None
        raise AttributeError(name)


est2 = KerasEstimator(model=LinearModel(3, 1), optimizer=SGD(0.1), loss='mse', feature_cols=['features'], label_cols=['label'],
                      batch_size=32, epochs=1, store=store, backend=LocalBackend(2, env=env), verbose=0,
                      backend_env={'HVD_TEST_KERAS_BACKEND': 'numpy'}, callbacks=[EnvCheck(marker)], data_module=kdm.ParquetDataModule)
assert est2.getBackendEnv() == {'HVD_TEST_KERAS_BACKEND': 'numpy'} and est2.getDataModule() is kdm.ParquetDataModule
est2.fit(df)
assert open(marker).read() == 'numpy'
print('KERAS ESTIMATOR OK')
'''


def test_keras_estimator_on_fake_keras(native_built, tmp_path):
    env = dict(os.environ)
    fakes = os.path.join(HERE, 'fakes')
    env['PYTHONPATH'] = os.pathsep.join([fakes, ROOT, env.get('PYTHONPATH', '')])
    script = tmp_path / 'keras_est.py'
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script), str(tmp_path / 'store'), env['PYTHONPATH']], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and 'KERAS ESTIMATOR OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
