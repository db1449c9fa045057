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
