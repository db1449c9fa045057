"""Estimator API: fit a torch model on a DataFrame with data-parallel training, get a transformer back.

    python examples/spark_torch_estimator.py --num-proc 2            # pandas DataFrame, processes on this machine
    spark-submit ... examples/spark_torch_estimator.py --spark       # Spark DataFrame, training inside Spark barrier tasks

The same estimator classes run on both (`LocalBackend` / `SparkBackend`); the intermediate data format is Parquet in the
Store either way.  `--lightning` trains the same network through the LightningModule-protocol estimator instead.
"""
import argparse
import tempfile

import numpy as np
import pandas as pd
import torch

from horovod_b200.spark.common import LocalBackend, SparkBackend
from horovod_b200.spark.lightning import LightningEstimator
from horovod_b200.spark.torch import TorchEstimator


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(8, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3))

    def forward(self, features):
        return self.body(features)

    # --- LightningModule protocol (only used with --lightning) ---
    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=0.01)

    def training_step(self, batch, batch_idx):
        loss = torch.nn.functional.cross_entropy(self(batch['features'].float()), batch['label'])
        self.log('train_acc', (self(batch['features'].float()).argmax(1) == batch['label']).float().mean())
        return loss

    def validation_step(self, batch, batch_idx):
        return {'val_loss': torch.nn.functional.cross_entropy(self(batch['features'].float()), batch['label'])}


def make_frame(rows=1200, seed=0):
    centers = np.random.RandomState(42).randn(3, 8) * 2.0        # the classes; `seed` only draws the samples
    rng = np.random.RandomState(seed)
    label = rng.randint(0, 3, rows)
    feats = centers[label] + rng.randn(rows, 8)
    return pd.DataFrame({'features': list(feats.astype(np.float32)), 'label': label.astype(np.int64)})


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    p.add_argument('--num-proc', type=int, default=2)
    p.add_argument('--epochs', type=int, default=4)
    p.add_argument('--spark', action='store_true')
    p.add_argument('--lightning', action='store_true')
    p.add_argument('--store', default=None, help='Store prefix (a directory, hdfs://..., s3://..., dbfs:/...)')
    p.add_argument('--save-model', default=None, help='directory to save the fitted model to (Spark ML layout) and load it back from')
    a = p.parse_args()
    df = make_frame()
    if a.spark:
        from pyspark.sql import SparkSession
        df = SparkSession.builder.getOrCreate().createDataFrame(df)
        backend = SparkBackend(a.num_proc)
    else:
        backend = LocalBackend(a.num_proc)
    store = a.store or tempfile.mkdtemp(prefix='hvd-store-')
    torch.manual_seed(0)
    model = Net()
    common = dict(model=model, feature_cols=['features'], label_cols=['label'], batch_size=32, epochs=a.epochs, validation=0.2,
                  store=store, backend=backend, use_gpu=torch.cuda.is_available(), verbose=1)
    if a.lightning:
        est = LightningEstimator(**common)
    else:
        est = TorchEstimator(optimizer=torch.optim.Adam(model.parameters(), lr=0.01), loss=torch.nn.functional.cross_entropy, **common)
    fitted = est.fit(df)
    sample = make_frame(200, seed=1) if not a.spark else df.limit(200)
    out = fitted.transform(sample)
    if not a.spark:
        pred = np.array(out['label__output'].tolist()).argmax(1)
        acc = float((pred == out['label'].values).mean())
        print('history:', fitted.getHistory()[-1])
        print('held-out accuracy: %.3f' % acc)
        if a.save_model:
            fitted.write().overwrite().save(a.save_model)            # <dir>/metadata/part-00000, like any Spark ML stage
            again = type(fitted).load(a.save_model)
            pred2 = np.array(again.transform(sample)['label__output'].tolist()).argmax(1)
            assert (pred2 == pred).all() and again.getRunId() == fitted.getRunId()
            print('saved to and reloaded from', a.save_model)
        print('ESTIMATOR EXAMPLE OK' if acc > 0.8 else 'accuracy too low')
