"""Numpy-backed stand-in for the small part of MXNet that horovod_b200.mxnet touches (MXNet is not installed in this
image).  NOT MXNet; it only lets the front end's control flow run."""
import types

import numpy as _np

__version__ = '0.0-fake'


class _Ctx:
    device_type = 'cpu'


class NDArray:
    def __init__(self, a):
        self._a = _np.array(a)
        self.context = _Ctx()

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def shape(self):
        return self._a.shape

    def wait_to_read(self):
        pass

    def asnumpy(self):
        return self._a

    def astype(self, dt):
        return NDArray(self._a.astype(dt))

    def as_in_context(self, ctx):
        return self

    def __setitem__(self, k, v):
        self._a[k] = v._a if isinstance(v, NDArray) else v


def _array(a, dtype=None):
    return NDArray(_np.asarray(a, dtype=dtype))


nd = types.SimpleNamespace(array=_array, NDArray=NDArray)


class _Optimizer:
    def __init__(self, learning_rate=0.1, rescale_grad=1.0):
        self.lr, self.rescale_grad = learning_rate, rescale_grad

    def create_state(self, index, weight):
        return None

    def create_state_multi_precision(self, index, weight):
        return None

    def update(self, index, weight, grad, state):
        pairs = zip(weight, grad) if isinstance(index, (list, tuple)) else [(weight, grad)]
        for w, g in pairs:
            w[:] = w.asnumpy() - self.lr * self.rescale_grad * g.asnumpy()

    update_multi_precision = update

    def set_learning_rate(self, lr):
        self.lr = lr

    def set_lr_mult(self, m):
        self.lr_mult = m

    def set_wd_mult(self, m):
        self.wd_mult = m


optimizer = types.SimpleNamespace(Optimizer=_Optimizer)


class DeferredInitializationError(Exception):
    pass


class Parameter:
    def __init__(self, name, value=None, grad=None, grad_req='write'):
        self.name, self._value, self._grad, self.grad_req = name, value, grad, grad_req

    def data(self):
        if self._value is None:
            raise DeferredInitializationError(self.name)
        return self._value

    def list_grad(self):
        return [self._grad]

    def _init_impl(self, value):
        self._value = value


class _Trainer:
    def __init__(self, params, optimizer, optimizer_params=None, kvstore=None):
        self._params = list(params.values()) if isinstance(params, dict) else list(params)
        self._optimizer = optimizer
        self._scale = 1.0

    def step(self, batch_size):
        self._allreduce_grads()
        for p in self._params:
            if p.grad_req != 'null':
                p.data()[:] = p.data().asnumpy() - self._optimizer.lr * self._scale / batch_size * p.list_grad()[0].asnumpy()


gluon = types.SimpleNamespace(Trainer=_Trainer, parameter=types.SimpleNamespace(Parameter=Parameter,
                                                                                 DeferredInitializationError=DeferredInitializationError))
