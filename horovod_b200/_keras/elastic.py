"""Backend-independent logic of the elastic Keras callbacks (reference horovod/_keras/elastic.py: CommitStateCallbackImpl :17,
UpdateBatchStateCallbackImpl :42, UpdateEpochStateCallbackImpl :66).  `_keras/callbacks.py` mixes these with the Callback class
of the Keras in use; `backend` is that Keras' backend module (unused by the logic, kept for signature parity)."""


class CommitStateCallbackImpl:
    """state.commit() every `batches_per_commit` batches and at every epoch end."""

    def __init__(self, backend, state, batches_per_commit=1, *args):
        super().__init__(*args)
        self.backend, self.state, self.batches_per_commit = backend, state, batches_per_commit
        self.batches_remaining = batches_per_commit

    def on_train_begin(self, logs=None):
        self.batches_remaining = self.batches_per_commit

    def on_batch_end(self, batch, logs=None):
        self.batches_remaining -= 1
        if self.batches_remaining == 0:
            self.commit()

    def on_epoch_end(self, epoch, logs=None):
        self.commit()

    def commit(self):
        self.state.commit()
        self.batches_remaining = self.batches_per_commit


class UpdateBatchStateCallbackImpl:
    """Tracks state.batch so that a restarted epoch skips the batches already consumed."""

    def __init__(self, backend, state, *args):
        super().__init__(*args)
        self.backend, self.state = backend, state
        self.steps_per_epoch = None

    def on_train_begin(self, logs=None):
        self.steps_per_epoch = (getattr(self, 'params', None) or {}).get('steps')

    def on_epoch_begin(self, epoch, logs=None):
        params = getattr(self, 'params', None) or {}
        if self.steps_per_epoch and 'steps' in params:
            params['steps'] = self.steps_per_epoch - self.state.batch

    def on_batch_end(self, batch, logs=None):
        self.state.batch = batch

    def on_epoch_end(self, epoch, logs=None):
        self.state.batch = 0


class UpdateEpochStateCallbackImpl:
    """Tracks state.epoch: Keras restarts its epoch counter at 0 after a reset, the state keeps counting."""

    def __init__(self, backend, state, *args):
        super().__init__(*args)
        self.backend, self.state = backend, state
        self._initial = 0

    def on_train_begin(self, logs=None):
        self._initial = self.state.epoch

    def on_epoch_end(self, epoch, logs=None):
        self.state.epoch = self._initial + epoch + 1 if epoch < self._initial else epoch + 1
