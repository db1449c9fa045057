"""A plain nn.Module + optimizer + loss(es) presented through the LightningModule protocol (reference
horovod/spark/lightning/legacy.py `to_lightning_module` :23-115)."""
import torch


class LegacyModule(torch.nn.Module):
    """A plain nn.Module + optimizer + loss(es) presented through the protocol (reference: spark/lightning/legacy.py:20-115)."""

    def __init__(self, model, optimizer, loss_fns, loss_weights, feature_cols, label_cols, sample_weight_col=None):
        super().__init__()
        from horovod_b200.spark.torch.remote import _BatchLoss
        self.model = model
        self._optimizer = optimizer
        self._loss = _BatchLoss(model, loss_fns, loss_weights, None, list(feature_cols), list(label_cols), sample_weight_col)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def configure_optimizers(self):
        return self._optimizer

    def training_step(self, batch, batch_idx):
        return {'loss': self._loss(batch)}

    def validation_step(self, batch, batch_idx):
        return {'val_loss': self._loss(batch)}


def to_lightning_module(model, optimizer, loss_fns, loss_weights, feature_cols, label_cols, sample_weight_col=None):
    return LegacyModule(model, optimizer, loss_fns, loss_weights, feature_cols, label_cols, sample_weight_col)
