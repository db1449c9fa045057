"""Training-loop helpers with the semantics of the reference's Keras callbacks (horovod/_keras/callbacks.py), for plain
PyTorch loops: metric averaging across ranks, LR warm-up / schedules scaled by hvd.size(), state broadcast at start."""
import torch

from horovod_b200.torch import mpi_ops
from horovod_b200.torch.functions import broadcast_optimizer_state, broadcast_parameters


def broadcast_global_state(model, optimizer=None, root_rank=0):
    """BroadcastGlobalVariablesCallback: make every rank start from root_rank's weights (and optimizer state)."""
    broadcast_parameters(model.state_dict(), root_rank=root_rank)
    if optimizer is not None:
        broadcast_optimizer_state(optimizer, root_rank=root_rank)


class MetricAverage(object):
    """MetricAverageCallback: `avg = MetricAverage(); logs = avg({'loss': 0.3, 'acc': 0.9})` returns the metrics
    averaged over all ranks (one fused allreduce for the whole dict)."""

    def __init__(self, device=None):
        self.device = device

    def __call__(self, metrics):
        keys = sorted(metrics.keys())
        t = torch.tensor([float(metrics[k]) for k in keys], dtype=torch.float64, device=self.device or 'cpu')
        t = mpi_ops.allreduce(t, op=mpi_ops.Average, name='metric_average.' + '.'.join(keys)[:200])
        return {k: v for k, v in zip(keys, t.tolist())}


class LearningRateSchedule(object):
    """LearningRateScheduleCallback: lr = initial_lr * multiplier(epoch) for start_epoch <= epoch < end_epoch.
    `multiplier` may be a constant or a function of the (fractional) epoch; momentum correction is applied for SGD."""

    def __init__(self, optimizer, initial_lr, multiplier, start_epoch=0, end_epoch=None, staircase=True,
                 momentum_correction=True, steps_per_epoch=None):
        self.optimizer = optimizer
        self.initial_lr = initial_lr
        self.start_epoch = start_epoch
        self.end_epoch = end_epoch
        self.staircase = staircase
        self.momentum_correction = momentum_correction
        self.steps_per_epoch = steps_per_epoch
        self.multiplier = multiplier if callable(multiplier) else (lambda epoch: multiplier)
        self._restore = None

    def _adjust(self, epoch):
        old = self.optimizer.param_groups[0]['lr']
        new = self.initial_lr * self.multiplier(epoch)
        for g in self.optimizer.param_groups:
            g['lr'] = new
            if self.momentum_correction and 'momentum' in g and old > 0:
                # keep the effective step of the momentum buffer continuous across an LR change (Goyal et al. 2017)
                self._restore = g['momentum']
                g['momentum'] = g['momentum'] * new / old

    def on_batch_begin(self, epoch, batch):
        if epoch < self.start_epoch or (self.end_epoch is not None and epoch >= self.end_epoch):
            return
        if self.staircase and batch == 0:
            self._adjust(epoch)
        elif not self.staircase:
            assert self.steps_per_epoch, 'steps_per_epoch is required for smooth schedules'
            self._adjust(epoch + float(batch) / self.steps_per_epoch)

    def on_batch_end(self):
        if self._restore is not None:
            for g in self.optimizer.param_groups:
                if 'momentum' in g:
                    g['momentum'] = self._restore
            self._restore = None


class LearningRateWarmup(LearningRateSchedule):
    """LearningRateWarmupCallback: ramp from initial_lr / size to initial_lr over `warmup_epochs` (gradual warm-up for
    large-batch training, arXiv:1706.02677)."""

    def __init__(self, optimizer, initial_lr, warmup_epochs=5, momentum_correction=True, steps_per_epoch=None, verbose=0):
        def multiplier(epoch):
            epoch += 1.0 / steps_per_epoch if steps_per_epoch else 0
            return 1.0 / mpi_ops.size() * (epoch * (mpi_ops.size() - 1) / warmup_epochs + 1)
        super().__init__(optimizer, initial_lr, multiplier, start_epoch=0, end_epoch=warmup_epochs, staircase=False,
                         momentum_correction=momentum_correction, steps_per_epoch=steps_per_epoch)
        self.verbose = verbose
