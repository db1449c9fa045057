"""A small trainer for objects that follow the LightningModule protocol.

The reference's Lightning estimator hands the module to `pytorch_lightning.Trainer(strategy=HorovodStrategy)`
(horovod/spark/lightning/remote.py:100-300).  pytorch_lightning is an optional dependency that pins its own distributed
strategy API; this trainer needs only the PROTOCOL — `configure_optimizers`, `training_step`, optional `validation_step` and
epoch hooks, `self.log` — so any `LightningModule` (or a plain class with those methods) trains on the hvd runtime with
`hvd.DistributedOptimizer`, parameter / optimizer-state broadcast, metric averaging and rank-0 checkpoints.

Protocol subset that is honoured:
  configure_optimizers() -> optimizer | [optimizer] | ([optimizers], [schedulers]) | {'optimizer': o, 'lr_scheduler': s | {'scheduler': s, 'interval': 'epoch'|'step'}}
  training_step(batch, batch_idx) -> loss tensor | {'loss': tensor, ...}
  validation_step(batch, batch_idx) -> tensor | {'val_loss': tensor, ...} | None   (values are averaged)
  on_fit_start / on_train_start / on_train_epoch_start / on_train_epoch_end / on_validation_epoch_end / on_fit_end / on_train_end
  self.log(name, value, ...), self.log_dict({...})  -> averaged per epoch, across ranks, into the history
One optimizer is supported (the reference has the same restriction through its Horovod strategy).
"""
import torch


def _first(x):
    return x[0] if isinstance(x, (list, tuple)) else x


def parse_optimizers(configured):
    """-> (optimizer, scheduler or None, 'epoch' | 'step')"""
    sched, interval = None, 'epoch'
    if isinstance(configured, dict):
        opt = configured['optimizer']
        s = configured.get('lr_scheduler')
        if isinstance(s, dict):
            sched, interval = s['scheduler'], s.get('interval', 'epoch')
        else:
            sched = s
    elif isinstance(configured, tuple) and len(configured) == 2 and isinstance(configured[0], (list, tuple)):
        opts, scheds = configured
        if len(opts) != 1:
            raise ValueError('exactly one optimizer is supported, configure_optimizers returned %d' % len(opts))
        opt = opts[0]
        s = _first(scheds) if scheds else None
        if isinstance(s, dict):
            sched, interval = s['scheduler'], s.get('interval', 'epoch')
        else:
            sched = s
    elif isinstance(configured, (list, tuple)):
        if len(configured) != 1:
            raise ValueError('exactly one optimizer is supported, configure_optimizers returned %d' % len(configured))
        opt = configured[0]
    else:
        opt = configured
    if not isinstance(opt, torch.optim.Optimizer):
        raise ValueError('configure_optimizers must return a torch optimizer, got %r' % type(opt))
    return opt, sched, interval


class _LogSink:
    """Collects `self.log` calls of one epoch: name -> (sum, count)."""

    def __init__(self):
        self.acc = {}

    def log(self, name, value, *args, **kwargs):
        v = value.detach().float().mean() if torch.is_tensor(value) else torch.tensor(float(value))
        s, c = self.acc.get(name, (0.0, 0))
        self.acc[name] = (s + v, c + 1)

    def log_dict(self, values, *args, **kwargs):
        for k, v in values.items():
            self.log(k, v)

    def drain(self, device):
        out = {k: (torch.as_tensor(s, device=device, dtype=torch.float32) / max(c, 1)) for k, (s, c) in self.acc.items()}
        self.acc = {}
        return out


def _hook(module, name, *args):
    fn = getattr(module, name, None)
    if callable(fn):
        return fn(*args)


class ModuleProtocolTrainer:
    """fit(module, train_loader, val_loader=None) -> history (list of per-epoch dicts, identical on every rank)."""

    def __init__(self, hvd, device, epochs=1, first_epoch=0, compression=None, backward_passes_per_step=1, gradient_clip_val=None,
                 callbacks=(), checkpoint=None, verbose=0, prefetcher=None, logger=None, log_every_n_steps=50, terminate_on_nan=False):
        """`logger`: object with `log_metrics(dict, step=)` (any pytorch_lightning logger qualifies), called on rank 0 every
        `log_every_n_steps` training steps with the step's loss and once per epoch with the averaged record.
        `terminate_on_nan`: a non-finite training loss raises ValueError on the rank that sees it (costs one device -> host read
        per step)."""
        self.logger, self.log_every_n_steps, self.terminate_on_nan = logger, max(1, int(log_every_n_steps or 50)), terminate_on_nan
        self.hvd, self.device, self.epochs, self.first_epoch = hvd, device, epochs, first_epoch
        self.compression, self.accumulate = compression, backward_passes_per_step
        self.clip, self.callbacks, self.checkpoint, self.verbose = gradient_clip_val, list(callbacks), checkpoint, verbose
        self.prefetcher = prefetcher or (lambda loader: loader)
        self.optimizer = None

    def setup(self, module, optimizer_state=None):
        hvd = self.hvd
        module.to(self.device)
        base, self.scheduler, self.interval = parse_optimizers(module.configure_optimizers())
        if optimizer_state is not None:
            base.load_state_dict(optimizer_state)
        named = module.named_parameters() if hasattr(module, 'named_parameters') else None
        self.optimizer = hvd.DistributedOptimizer(base, named_parameters=named, compression=self.compression or hvd.Compression.none,
                                                  backward_passes_per_step=self.accumulate)
        hvd.broadcast_parameters(module.state_dict(), root_rank=0)
        hvd.broadcast_optimizer_state(self.optimizer, root_rank=0)
        return self.optimizer

    def _average(self, values, tag):
        names = sorted(values)
        if not names:
            return {}
        vec = self.hvd.allreduce(torch.stack([values[n].to(self.device) for n in names]), name='pl.%s' % tag)
        return dict(zip(names, vec.tolist()))

    def fit(self, module, train_loader, val_loader=None):
        if self.optimizer is None:
            self.setup(module)
        sink = _LogSink()
        module.log, module.log_dict = sink.log, sink.log_dict
        opt, hvd = self.optimizer, self.hvd
        history = []
        global_step = 0
        _hook(module, 'on_fit_start')
        _hook(module, 'on_train_start')
        for epoch in range(self.first_epoch, self.epochs):
            module.train()
            module.current_epoch_ = epoch
            _hook(module, 'on_train_epoch_start')
            total, count = torch.zeros((), device=self.device), 0
            opt.zero_grad()
            for i, batch in enumerate(self.prefetcher(train_loader)):
                out = module.training_step(batch, i)
                loss = out['loss'] if isinstance(out, dict) else out
                if self.terminate_on_nan and not bool(torch.isfinite(loss.detach()).all()):
                    raise ValueError('The loss returned in `training_step` is %s at epoch %d, step %d.' % (loss.detach().tolist(), epoch, i))
                global_step += 1
                if self.logger is not None and hvd.rank() == 0 and global_step % self.log_every_n_steps == 0:
                    self.logger.log_metrics({'loss': float(loss.detach())}, step=global_step)
                (loss / self.accumulate).backward()
                if (i + 1) % self.accumulate == 0:
                    if self.clip:
                        opt.synchronize()
                        torch.nn.utils.clip_grad_norm_(module.parameters(), self.clip)
                        with opt.skip_synchronize():
                            opt.step()
                    else:
                        opt.step()
                    opt.zero_grad()
                    if self.scheduler is not None:
                        # the scheduler watches the optimizer it was built on; the step ran through the distributed wrapper
                        setattr(self.scheduler.optimizer, '_opt_called', True)
                    if self.scheduler is not None and self.interval == 'step':
                        self.scheduler.step()
                total += loss.detach()
                count += 1
            if self.scheduler is not None and self.interval == 'epoch':
                self.scheduler.step()
            _hook(module, 'on_train_epoch_end')
            logged = sink.drain(self.device)
            logged['loss'] = total / max(count, 1)
            record = {'epoch': epoch}
            record.update(self._average(logged, 'train'))
            if val_loader is not None and callable(getattr(module, 'validation_step', None)):
                module.eval()
                sums, n = {}, 0
                with torch.no_grad():
                    for i, batch in enumerate(self.prefetcher(val_loader)):
                        out = module.validation_step(batch, i)
                        if out is not None:
                            items = out.items() if isinstance(out, dict) else [('val_loss', out)]
                            for k, v in items:
                                if torch.is_tensor(v):
                                    sums[k] = sums.get(k, 0.0) + v.detach().float().mean()
                        n += 1
                _hook(module, 'on_validation_epoch_end')
                vals = {k: v / max(n, 1) for k, v in sums.items()}
                vals.update(sink.drain(self.device))
                record.update(self._average(vals, 'val'))
            history.append(record)
            if self.logger is not None and hvd.rank() == 0:
                self.logger.log_metrics({k: v for k, v in record.items() if k != 'epoch'}, step=global_step)
            for cb in self.callbacks:
                cb(epoch, record) if callable(cb) else _hook(cb, 'on_epoch_end', epoch, record)
            if self.verbose and hvd.rank() == 0:
                print('epoch %d: %s' % (epoch, record), flush=True)
            if self.checkpoint is not None and hvd.rank() == 0:
                self.checkpoint(module, opt, epoch)
        _hook(module, 'on_train_end')
        _hook(module, 'on_fit_end')
        return history


from horovod_b200.spark.lightning.legacy import LegacyModule, to_lightning_module  # noqa: E402,F401
