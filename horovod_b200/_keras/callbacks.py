"""Keras callbacks (parity: horovod/_keras/callbacks.py:23-215 and _keras/elastic.py:17-85): broadcast of the initial
state, metric averaging, LR schedule / warm-up with momentum correction, elastic commit / batch / epoch bookkeeping.
Written against the public Keras callback protocol only (`self.model`, `self.params`, `logs`)."""

import tensorflow as tf

import horovod_b200.tensorflow as hvd


class BroadcastGlobalVariablesCallback(tf.keras.callbacks.Callback):
    """Broadcasts model + optimizer variables from `root_rank` after the first batch (the optimizer slots only exist
    once one step ran)."""

    def __init__(self, root_rank=0, device='', process_set=hvd.global_process_set):
        super().__init__()
        self.root_rank, self.process_set = root_rank, process_set
        self.broadcast_done = False
        self._local = []

    def register_local_var(self, var):
        self._local.append(var.ref() if hasattr(var, 'ref') else id(var))

    def _vars(self):
        opt = getattr(self.model, 'optimizer', None)
        ov = []
        if opt is not None:
            ov = opt.variables() if callable(getattr(opt, 'variables', None)) else list(getattr(opt, 'variables', []))
        allv = list(self.model.variables) + list(ov)
        key = (lambda v: v.ref()) if allv and hasattr(allv[0], 'ref') else id
        return [v for v in allv if key(v) not in self._local]

    def on_batch_end(self, batch, logs=None):
        if self.broadcast_done:
            return
        hvd.broadcast_variables(self._vars(), self.root_rank, process_set=self.process_set)
        self.broadcast_done = True


class MetricAverageCallback(tf.keras.callbacks.Callback):
    """Averages the epoch-end metrics over ranks, in place in `logs`, in sorted key order on every rank."""

    def __init__(self, device='', process_set=hvd.global_process_set):
        super().__init__()
        self.process_set = process_set

    def on_epoch_end(self, epoch, logs=None):
        if not logs:
            return
        keys = sorted(k for k, v in logs.items() if isinstance(v, (int, float)) or hasattr(v, 'dtype'))
        if not keys:
            return
        vals = tf.constant([float(logs[k]) for k in keys], dtype=tf.float64)
        avg = hvd.allreduce(vals, op=hvd.Average, name='metric_average', process_set=self.process_set).numpy()
        for k, v in zip(keys, avg):
            logs[k] = float(v)


def _get_lr(opt):
    lr = opt.learning_rate if hasattr(opt, 'learning_rate') else opt.lr
    return float(tf.keras.backend.get_value(lr))


def _set_lr(opt, value):
    lr = opt.learning_rate if hasattr(opt, 'learning_rate') else opt.lr
    if hasattr(lr, 'assign'):
        lr.assign(value)
    else:
        tf.keras.backend.set_value(lr, value)


class LearningRateScheduleCallback(tf.keras.callbacks.Callback):
    """lr = initial_lr * multiplier(epoch) for start_epoch <= epoch < end_epoch; `staircase=False` evaluates the multiplier
    at fractional epochs every batch.  With `momentum_correction` the momentum is rescaled by new_lr/old_lr for the one
    batch in which the LR changed (Goyal et al. 2017, as the reference does)."""

    def __init__(self, initial_lr, multiplier, start_epoch=0, end_epoch=None, staircase=True, momentum_correction=True,
                 steps_per_epoch=None):
        super().__init__()
        if initial_lr is None:
            raise ValueError('Parameter `initial_lr` is required')
        self.initial_lr, self.start_epoch, self.end_epoch = initial_lr, start_epoch, end_epoch
        self.staircase, self.momentum_correction, self.steps_per_epoch = staircase, momentum_correction, steps_per_epoch
        self.multiplier = multiplier if callable(multiplier) else (lambda epoch: multiplier)
        if not callable(multiplier):
            self.staircase = True
        self.current_epoch = None
        self._restore_momentum = None

    def _steps(self):
        if self.steps_per_epoch:
            return self.steps_per_epoch
        p = self.params or {}
        if p.get('steps'):
            return p['steps']
        if p.get('samples') and p.get('batch_size'):
            return -(-p['samples'] // p['batch_size'])
        raise ValueError('Could not autodetect the number of steps per epoch. Please specify the steps_per_epoch parameter.')

    def _in_range(self, epoch):
        return epoch >= self.start_epoch and (self.end_epoch is None or epoch < self.end_epoch)

    def _adjust(self, epoch):
        opt = self.model.optimizer
        old = _get_lr(opt)
        new = self.initial_lr * self.multiplier(epoch)
        _set_lr(opt, new)
        if self.momentum_correction and hasattr(opt, 'momentum') and old > 0:
            m = opt.momentum
            self._restore_momentum = float(tf.keras.backend.get_value(m))
            corrected = self._restore_momentum * new / old
            m.assign(corrected) if hasattr(m, 'assign') else setattr(opt, 'momentum', corrected)

    def _restore(self):
        if self._restore_momentum is not None:
            m = self.model.optimizer.momentum
            m.assign(self._restore_momentum) if hasattr(m, 'assign') else setattr(self.model.optimizer, 'momentum', self._restore_momentum)
            self._restore_momentum = None

    def on_train_begin(self, logs=None):
        if not self.staircase:
            self._steps()

    def on_epoch_begin(self, epoch, logs=None):
        self.current_epoch = epoch

    def on_batch_begin(self, batch, logs=None):
        if not self._in_range(self.current_epoch):
            return
        if self.staircase and batch == 0:
            self._adjust(self.current_epoch)
        elif not self.staircase:
            self._adjust(self.current_epoch + float(batch) / self._steps())

    def on_batch_end(self, batch, logs=None):
        self._restore()

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs['lr'] = _get_lr(self.model.optimizer)


class LearningRateWarmupCallback(LearningRateScheduleCallback):
    """Ramps from initial_lr/size to initial_lr over `warmup_epochs` (linear in fractional epochs)."""

    def __init__(self, initial_lr, warmup_epochs=5, momentum_correction=True, steps_per_epoch=None, verbose=0):
        def multiplier(epoch):
            epoch += 1.0 / self._steps()
            return 1.0 / hvd.size() * (epoch * (hvd.size() - 1) / warmup_epochs + 1)
        super().__init__(initial_lr, multiplier, start_epoch=0, end_epoch=warmup_epochs, staircase=False,
                         momentum_correction=momentum_correction, steps_per_epoch=steps_per_epoch)
        self.verbose = verbose

    def on_epoch_end(self, epoch, logs=None):
        super().on_epoch_end(epoch, logs)
        if epoch == self.end_epoch - 1 and self.verbose > 0:
            print('\nEpoch %d: finished gradual learning rate warmup to %g.' % (epoch + 1, _get_lr(self.model.optimizer)))


# ---- elastic (logic in _keras/elastic.py, mixed with this Keras' Callback class) ---------------------------------------------
from horovod_b200._keras import elastic as _elastic_impl  # noqa: E402


class CommitStateCallback(_elastic_impl.CommitStateCallbackImpl, tf.keras.callbacks.Callback):
    """state.commit() every `batches_per_commit` batches and at every epoch end."""

    def __init__(self, state, batches_per_commit=1):
        super().__init__(tf.keras.backend if hasattr(tf.keras, 'backend') else None, state, batches_per_commit)


class UpdateBatchStateCallback(_elastic_impl.UpdateBatchStateCallbackImpl, tf.keras.callbacks.Callback):
    """Tracks state.batch so that a restarted epoch skips the batches already consumed."""

    def __init__(self, state):
        super().__init__(tf.keras.backend if hasattr(tf.keras, 'backend') else None, state)


class UpdateEpochStateCallback(_elastic_impl.UpdateEpochStateCallbackImpl, tf.keras.callbacks.Callback):
    """Tracks state.epoch across resets."""

    def __init__(self, state):
        super().__init__(tf.keras.backend if hasattr(tf.keras, 'backend') else None, state)


class BestModelCheckpoint(tf.keras.callbacks.Callback):
    """Keeps the best model seen so far (by `monitor`) at `filepath`; the estimators set `filepath` to the run's checkpoint
    in the Store (reference _keras/callbacks.py `BestModelCheckpoint`: a ModelCheckpoint with save_best_only=True whose
    path is filled in later).  Written on the Callback protocol only, so it does not depend on the ModelCheckpoint class of
    a particular Keras version; `save_fn(model, filepath)` overrides how the model is written."""

    def __init__(self, monitor='val_loss', verbose=0, mode='auto', save_freq='epoch', filepath=None, save_weights_only=False,
                 save_fn=None):
        super().__init__()
        if mode not in ('auto', 'min', 'max'):
            raise ValueError("mode must be 'auto', 'min' or 'max'")
        if mode == 'auto':
            mode = 'max' if any(key in monitor for key in ('acc', 'auc', 'fmeasure', 'f1')) else 'min'
        self.monitor, self.verbose, self.mode, self.save_freq = monitor, verbose, mode, save_freq
        self.filepath, self.save_weights_only, self.save_fn = filepath, save_weights_only, save_fn
        self.best = float('inf') if mode == 'min' else float('-inf')
        self.best_epoch = None

    def _improved(self, value):
        return value < self.best if self.mode == 'min' else value > self.best

    def on_epoch_end(self, epoch, logs=None):
        value = (logs or {}).get(self.monitor)
        if value is None:
            return
        value = float(value)
        if not self._improved(value):
            return
        self.best, self.best_epoch = value, epoch
        if self.filepath is None:
            return
        path = self.filepath.format(epoch=epoch + 1, **(logs or {})) if '{' in self.filepath else self.filepath
        if self.save_fn is not None:
            self.save_fn(self.model, path)
        elif self.save_weights_only:
            self.model.save_weights(path)
        else:
            self.model.save(path)
        if self.verbose:
            print('Epoch %d: %s improved to %.5f, saving model to %s' % (epoch + 1, self.monitor, value, path))


# the reference splits every callback into a backend-independent `...Impl` and the Keras subclass; here the classes above are
# written against the Callback protocol directly, so the Impl names are the classes themselves
BroadcastGlobalVariablesCallbackImpl = BroadcastGlobalVariablesCallback
MetricAverageCallbackImpl = MetricAverageCallback
LearningRateScheduleCallbackImpl = LearningRateScheduleCallback
LearningRateWarmupCallbackImpl = LearningRateWarmupCallback
