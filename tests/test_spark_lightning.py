"""LightningEstimator: protocol modules (training_step / validation_step / configure_optimizers / self.log) and the legacy
nn.Module + optimizer + loss form, trained by 2 local processes.  Reference coverage model:
test/integration/test_spark_lightning.py (fit_model, legacy, restore from checkpoint, direct parquet train)."""
import io

import numpy as np
import pandas as pd
import pytest
import torch

from horovod_b200.spark.common import LocalBackend
from horovod_b200.spark.lightning import LightningEstimator, ModuleProtocolTrainer
from horovod_b200.spark.lightning.trainer import parse_optimizers


class Regressor(torch.nn.Module):
    """Follows the LightningModule protocol without importing pytorch_lightning."""

    def __init__(self, lr=0.1):
        super().__init__()
        self.net = torch.nn.Linear(3, 1)
        self.lr = lr
        self.epoch_starts = 0

    def forward(self, x):
        return self.net(x)

    def configure_optimizers(self):
        opt = torch.optim.SGD(self.parameters(), lr=self.lr, momentum=0.5)
        return [opt], [{'scheduler': torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=0.5), 'interval': 'epoch'}]

    def on_train_epoch_start(self):
        self.epoch_starts += 1

    def training_step(self, batch, batch_idx):
        pred = self(batch['features'].float()).squeeze(-1)
        loss = torch.nn.functional.mse_loss(pred, batch['label'].float())
        self.log('train_mae', (pred - batch['label']).abs().mean())
        return loss

    def validation_step(self, batch, batch_idx):
        pred = self(batch['features'].float()).squeeze(-1)
        return {'val_loss': torch.nn.functional.mse_loss(pred, batch['label'].float())}


def _frame(n=512, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.randn(n, 3).astype(np.float32)
    w = np.array([1.5, -2.0, 0.5], dtype=np.float32)
    return pd.DataFrame({'features': list(x), 'label': x @ w + 0.25}), w


def test_parse_optimizers_forms():
    m = torch.nn.Linear(2, 1)
    o = torch.optim.SGD(m.parameters(), lr=0.1)
    s = torch.optim.lr_scheduler.StepLR(o, 1)
    assert parse_optimizers(o) == (o, None, 'epoch')
    assert parse_optimizers([o]) == (o, None, 'epoch')
    assert parse_optimizers(([o], [s])) == (o, s, 'epoch')
    assert parse_optimizers({'optimizer': o, 'lr_scheduler': {'scheduler': s, 'interval': 'step'}}) == (o, s, 'step')
    assert parse_optimizers({'optimizer': o, 'lr_scheduler': s}) == (o, s, 'epoch')
    with pytest.raises(ValueError):
        parse_optimizers([o, o])
    with pytest.raises(ValueError):
        parse_optimizers('sgd')


def test_param_accessors_and_validation(tmp_path):
    est = LightningEstimator(model=Regressor(), feature_cols=['features'], label_cols=['label'], store=str(tmp_path))
    assert est.getBatchSize() == 32 and est.setBatchSize(16) is est and est.getBatchSize() == 16 and est.batch_size == 16
    assert est.setGradientClipVal(1.0).getGradientClipVal() == 1.0
    with pytest.raises(ValueError):
        est.setBatchSize(0)
    with pytest.raises(ValueError):
        est.setValidation(1.5)
    with pytest.raises(TypeError):
        est.setParams(no_such_knob=1)
    with pytest.raises(ValueError):   # plain module without optimizer / loss
        LightningEstimator(model=torch.nn.Linear(3, 1), feature_cols=['features'], label_cols=['label'], store=str(tmp_path))
    clone = est.copy({'epochs': 7})
    assert clone.getEpochs() == 7 and est.getEpochs() == 1
    assert 'batch_size' in est.explainParams()


def test_fit_protocol_module_two_procs_and_resume(native_built, tmp_path):
    df, w = _frame()
    torch.manual_seed(0)
    est = LightningEstimator(model=Regressor(), feature_cols=['features'], label_cols=['label'], batch_size=32, epochs=4,
                             validation=0.2, store=str(tmp_path / 'store'), backend=LocalBackend(2), use_gpu=False, verbose=0,
                             run_id='pl1', gradient_clip_val=5.0)
    model = est.fit(df)
    hist = model.getHistory()
    assert [h['epoch'] for h in hist] == [0, 1, 2, 3]
    assert hist[-1]['loss'] < 0.1 * hist[0]['loss'] and hist[-1]['val_loss'] < 0.1 and 'train_mae' in hist[0]
    np.testing.assert_allclose(model.getModel().net.weight.detach().numpy().ravel(), w, atol=0.15)
    out = model.transform(df.head(5))
    np.testing.assert_allclose(np.array(out['label__output'].tolist()), df['label'].values[:5], atol=0.4)
    # same run id, more epochs: continues after the stored epoch instead of starting over
    more = est.fit(df, params={'epochs': 6})
    assert [h['epoch'] for h in more.getHistory()] == [4, 5]
    ck = torch.load(io.BytesIO(est.store.read(est.store.get_checkpoint_path('pl1'))), weights_only=False)
    assert ck['epoch'] == 5


def test_fit_legacy_module_and_fit_on_parquet(native_built, tmp_path):
    from horovod_b200.spark.common import util
    df, w = _frame(256, seed=1)
    torch.manual_seed(1)
    net = torch.nn.Linear(3, 1)
    est = LightningEstimator(model=net, optimizer=torch.optim.SGD(net.parameters(), lr=0.1), loss=torch.nn.functional.mse_loss,
                             feature_cols=['features'], label_cols=['label'], batch_size=16, epochs=5, store=str(tmp_path / 's'),
                             backend=LocalBackend(2), use_gpu=False, verbose=0)
    store = est.store
    rows = util.write_parquet(df, store.get_train_data_path(), store, 2, ['features', 'label'])
    assert rows == 256
    model = est.fit_on_parquet()
    assert model.getHistory()[-1]['loss'] < 0.05
    np.testing.assert_allclose(model.getModel().model.weight.detach().numpy().ravel(), w, atol=0.1)


def test_protocol_trainer_single_process_hooks(native_built):
    """world size 1, in process: hooks fire, logged values land in the history, step-interval scheduler steps per batch."""
    import horovod_b200.torch as hvd
    hvd.init()
    try:
        class M(Regressor):
            def configure_optimizers(self):
                opt = torch.optim.SGD(self.parameters(), lr=0.1)
                return {'optimizer': opt, 'lr_scheduler': {'scheduler': torch.optim.lr_scheduler.StepLR(opt, 1, 0.9), 'interval': 'step'}}
        df, _ = _frame(64)
        x, y = torch.tensor(np.stack(df['features'])), torch.tensor(df['label'].values, dtype=torch.float32)
        batches = [{'features': x[i:i + 16], 'label': y[i:i + 16]} for i in range(0, 64, 16)]
        m = M()
        seen = []
        tr = ModuleProtocolTrainer(hvd, torch.device('cpu'), epochs=3, callbacks=[lambda e, r: seen.append(e)])
        hist = tr.fit(m, batches, batches)
        assert seen == [0, 1, 2] and m.epoch_starts == 3 and hist[2]['loss'] < hist[0]['loss']
        assert abs(tr.optimizer.param_groups[0]['lr'] - 0.1 * 0.9 ** 12) < 1e-9
        assert {'loss', 'val_loss', 'train_mae', 'epoch'} <= set(hist[0])

        class Logger:
            def __init__(self):
                self.calls = []

            def log_metrics(self, metrics, step=None):
                self.calls.append((step, dict(metrics)))
        lg = Logger()
        ModuleProtocolTrainer(hvd, torch.device('cpu'), epochs=2, logger=lg, log_every_n_steps=3).fit(M(), batches, batches)
        steps = [s for s, _ in lg.calls]
        assert steps == [3, 4, 6, 8] and set(lg.calls[0][1]) == {'loss'} and {'loss', 'val_loss'} <= set(lg.calls[1][1]), lg.calls

        class Diverges(M):
            def training_step(self, batch, batch_idx):
                return super().training_step(batch, batch_idx) * float('nan') if batch_idx == 2 else super().training_step(batch, batch_idx)
        with pytest.raises(ValueError, match='step 2'):
            ModuleProtocolTrainer(hvd, torch.device('cpu'), epochs=1, terminate_on_nan=True).fit(Diverges(), batches)
    finally:
        hvd.shutdown()


def test_trainer_args_loss_constructors_async_readers_and_module_paths(native_built, tmp_path):
    """The reference's Lightning-estimator knobs: trainer_args (translated for the protocol trainer), loss_constructors for a
    plain nn.Module, async reader threads with their queue sizes; plus the reference's module paths."""
    from horovod_b200.spark.lightning import datamodule, legacy, remote, util as lutil
    assert remote.translate_trainer_args({'max_epochs': 3, 'gpus': 1, 'accumulate_grad_batches': 2}) == {'epochs': 3, 'backward_passes_per_step': 2}
    with pytest.raises(ValueError, match='limit_train_batches'):
        remote.translate_trainer_args({'limit_train_batches': 0.5})
    assert datamodule.PetastormDataModule is datamodule.ParquetDataModule and callable(remote.RemoteTrainer({}))
    m = torch.nn.Linear(2, 1)
    back = lutil.deserialize_fn()(lutil.serialize_fn()(m))
    assert torch.equal(back.weight, m.weight) and lutil.is_module_available('torch') and not lutil.is_module_available('pytorch_lightning_x')
    assert lutil.save_into_bio({'a': 1}, torch.save).read(2) == b'PK'

    df, w = _frame(256, seed=2)
    torch.manual_seed(2)
    net = torch.nn.Linear(3, 1)
    est = LightningEstimator(model=net, optimizer=torch.optim.SGD(net.parameters(), lr=0.1), loss_constructors=[lambda: torch.nn.MSELoss()],
                             feature_cols=['features'], label_cols=['label'], batch_size=16, epochs=1, validation=0.25,
                             store=str(tmp_path / 's'), backend=LocalBackend(2), use_gpu=False, verbose=0,
                             trainer_args={'max_epochs': 4, 'enable_progress_bar': False})
    est.setTrainReaderNumWorker(1).setTrainAsyncDataLoaderQueueSize(4).setValAsyncDataLoaderQueueSize(2).setNumGPUs(1)
    assert est.getTrainAsyncDataLoaderQueueSize() == 4 and est.getTrainerArgs()['max_epochs'] == 4 and est.getDebugDataLoader() is False
    model = est.fit(df)
    hist = model.getHistory()
    assert [h['epoch'] for h in hist] == [0, 1, 2, 3] and hist[-1]['loss'] < 0.1 * hist[0]['loss']      # max_epochs won over epochs=1
    assert isinstance(model.getModel(), legacy.LegacyModule) and model.getLossConstructors() and model.getLoss() is None
    with pytest.raises(ValueError, match='trainer_args'):
        est.copy({'trainer_args': {'precision': 16}}).fit(df)
