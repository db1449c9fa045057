"""A few lines of "Keras" on numpy for the estimator tests: a linear model with mean-squared-error, SGD, the callback
protocol and a fit() loop over a generator.  It exists so that the Keras estimator's control flow (optimizer wrapping,
broadcast / metric callbacks, checkpoints, weights hand-back) runs without TensorFlow installed."""
import numpy as np

import tensorflow as tf


class History:
    def __init__(self):
        self.history = {}


class SGD:
    def __init__(self, learning_rate=0.1):
        self.learning_rate = tf.Variable(np.float32(learning_rate), trainable=False, name='lr')
        self.iterations = 0

    def get_config(self):
        return {'learning_rate': float(self.learning_rate.numpy())}

    @classmethod
    def from_config(cls, cfg):
        return cls(**cfg)

    def variables(self):
        return []

    def apply_gradients(self, grads_and_vars):
        lr = float(self.learning_rate.numpy())
        for g, v in grads_and_vars:
            if g is not None:
                v.assign(v.numpy() - lr * np.asarray(g.numpy() if hasattr(g, 'numpy') else g))
        self.iterations += 1


class LinearModel:
    """y = x @ w + b, loss 'mse'."""

    def __init__(self, in_dim, out_dim=1, seed=0):
        rng = np.random.RandomState(seed)
        self.in_dim, self.out_dim, self.seed = in_dim, out_dim, seed
        self.w = tf.Variable(rng.randn(in_dim, out_dim).astype(np.float32) * 0.1, name='w')
        self.b = tf.Variable(np.zeros(out_dim, np.float32), name='b')
        self.optimizer, self.loss, self.metrics = None, None, None
        self.stop_training = False

    variables = property(lambda self: [self.w, self.b])
    trainable_variables = variables

    def get_config(self):
        return {'in_dim': self.in_dim, 'out_dim': self.out_dim, 'seed': self.seed}

    @classmethod
    def from_config(cls, cfg):
        return cls(**cfg)

    def get_weights(self):
        return [self.w.numpy().copy(), self.b.numpy().copy()]

    def set_weights(self, ws):
        self.w.assign(np.asarray(ws[0], np.float32))
        self.b.assign(np.asarray(ws[1], np.float32))

    def compile(self, optimizer=None, loss=None, loss_weights=None, metrics=None):
        assert loss in ('mse', 'mean_squared_error'), loss
        self.optimizer, self.loss, self.metrics = optimizer, loss, metrics or []

    def predict(self, x):
        return np.asarray(x, np.float32).reshape(len(x), self.in_dim) @ self.w.numpy() + self.b.numpy()

    def _step(self, x, y, w=None, train=True):
        x = np.asarray(x, np.float32).reshape(len(x), self.in_dim)
        y = np.asarray(y, np.float32).reshape(len(x), self.out_dim)
        err = x @ self.w.numpy() + self.b.numpy() - y
        sw = np.ones(len(x), np.float32) if w is None else np.asarray(w, np.float32)
        loss = float(np.mean(sw[:, None] * err ** 2))
        if train:
            scale = 2.0 * sw[:, None] * err / err.size
            self.optimizer.apply_gradients([(tf.constant(x.T @ scale), self.w), (tf.constant(scale.sum(0)), self.b)])
        return loss, float(np.mean(np.abs(err)))

    def fit(self, x, steps_per_epoch=None, epochs=1, initial_epoch=0, callbacks=(), verbose=0, validation_data=None, validation_steps=None):
        history = History()
        for cb in callbacks:
            cb.set_model(self)
            cb.params = {'epochs': epochs, 'steps': steps_per_epoch}
        for cb in callbacks:
            getattr(cb, 'on_train_begin', lambda logs=None: None)()
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                getattr(cb, 'on_epoch_begin', lambda e, logs=None: None)(epoch)
            losses = []
            for step in range(steps_per_epoch):
                item = next(x)
                for cb in callbacks:
                    getattr(cb, 'on_batch_begin', lambda b, logs=None: None)(step)
                loss, _ = self._step(*item)
                losses.append(loss)
                for cb in callbacks:
                    getattr(cb, 'on_batch_end', lambda b, logs=None: None)(step, {'loss': loss})
            logs = {'loss': float(np.mean(losses))}
            if validation_data is not None:
                vals = [self._step(*next(validation_data), train=False) for _ in range(validation_steps)]
                logs['val_loss'] = float(np.mean([v[0] for v in vals]))
                if 'mae' in self.metrics:
                    logs['val_mae'] = float(np.mean([v[1] for v in vals]))
            for cb in callbacks:
                getattr(cb, 'on_epoch_end', lambda e, logs=None: None)(epoch, logs)
            for k, v in logs.items():
                history.history.setdefault(k, []).append(v)
            if self.stop_training:
                break
        for cb in callbacks:
            getattr(cb, 'on_train_end', lambda logs=None: None)()
        return history
