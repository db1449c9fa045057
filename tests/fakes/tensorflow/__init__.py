"""A tiny eager-only stand-in for TensorFlow (numpy-backed) used to exercise horovod_b200.tensorflow's control flow in
an image that has no TensorFlow.  It implements just the surface that front end touches; it is NOT TensorFlow and
passing these tests does not replace running the module against the real library."""
import types

import numpy as _np

__version__ = '0.0-fake'


class DType:
    def __init__(self, np_dtype, floating):
        self.np, self.is_floating = _np.dtype(np_dtype), floating

    def __eq__(self, o):
        return isinstance(o, DType) and self.np == o.np

    def __hash__(self):
        return hash(self.np)


float16, float32, float64 = DType('float16', True), DType('float32', True), DType('float64', True)
bfloat16 = DType('float16', True)  # stand-in
int32, int64, uint8 = DType('int32', False), DType('int64', False), DType('uint8', False)
_BY_NP = {d.np: d for d in (float32, float64, float16, int32, int64, uint8)}


class Tensor:
    device = '/job:localhost/replica:0/task:0/device:CPU:0'

    def __init__(self, a):
        self._a = _np.asarray(a)

    def numpy(self):
        return self._a

    @property
    def dtype(self):
        return _BY_NP[self._a.dtype]

    @property
    def shape(self):
        return self._a.shape

    def set_shape(self, s):
        pass

    def __getitem__(self, k):
        if isinstance(k, slice):
            k = slice(*[int(v.numpy()) if isinstance(v, Tensor) else v for v in (k.start, k.stop, k.step)])
        elif isinstance(k, Tensor):
            k = int(k.numpy())
        return Tensor(self._a[k])

    def __truediv__(self, o):
        return Tensor(self._a / (o._a if isinstance(o, Tensor) else o))

    def __mul__(self, o):
        return Tensor(self._a * (o._a if isinstance(o, Tensor) else o))

    def __sub__(self, o):
        return Tensor(self._a - (o._a if isinstance(o, Tensor) else o))

    def __add__(self, o):
        return Tensor(self._a + (o._a if isinstance(o, Tensor) else o))

    def __floordiv__(self, o):
        return Tensor(self._a // (o._a if isinstance(o, Tensor) else o))

    def __float__(self):
        return float(self._a)

    def __int__(self):
        return int(self._a)

    def __index__(self):
        return int(self._a)


class Variable(Tensor):
    def __init__(self, initial, trainable=True, name=None):
        super().__init__(_np.array(initial._a if isinstance(initial, Tensor) else initial))
        self.name = name or 'Variable:0'
        self.trainable = trainable

    def value(self):
        return Tensor(self._a.copy())

    read_value = value

    def assign(self, v):
        self._a = _np.array(v._a if isinstance(v, Tensor) else v, dtype=self._a.dtype).reshape(self._a.shape)
        return self

    def assign_add(self, v):
        self._a = self._a + (v._a if isinstance(v, Tensor) else v)
        return self

    def ref(self):
        return id(self)


class IndexedSlices:
    def __init__(self, values, indices, dense_shape=None):
        self.values, self.indices, self.dense_shape = values, indices, dense_shape


def is_tensor(x):
    return isinstance(x, Tensor)


def convert_to_tensor(x, dtype=None):
    if isinstance(x, IndexedSlices):
        dense = _np.zeros(tuple(int(d) for d in x.dense_shape), dtype=x.values._a.dtype)
        _np.add.at(dense, x.indices._a, x.values._a)
        return Tensor(dense)
    if isinstance(x, Tensor):
        return Tensor(x._a) if not isinstance(x, Variable) else Tensor(x._a.copy())
    return Tensor(_np.asarray(x, dtype=dtype.np if dtype else None))


def constant(v, dtype=None):
    return Tensor(_np.asarray(v, dtype=dtype.np if dtype else None))


def cast(x, dtype):
    return Tensor(convert_to_tensor(x)._a.astype(dtype.np))


def zeros_like(x):
    return Tensor(_np.zeros_like(x._a))


def fill(dims, value):
    return Tensor(_np.full([int(d) for d in dims], int(value) if isinstance(value, Tensor) else value))


def shape(x, out_type=None):
    return Tensor(_np.asarray(x._a.shape, dtype=(out_type or int32).np))


def reshape(x, s):
    return Tensor(x._a.reshape(s))


def reduce_sum(x, axis=None):
    return Tensor(_np.sum(x._a, axis=axis))


def reduce_mean(x, axis=None, keepdims=False):
    return Tensor(_np.mean(x._a, axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims))


def square(x):
    return Tensor(_np.square(x._a))


def executing_eagerly():
    return True


def custom_gradient(f):
    def wrapped(*args):
        out, grad = f(*args)
        wrapped.last_grad = grad
        return out
    return wrapped


class GradientTape:
    """Records nothing: `gradient` returns d(sum(w * x))/dw-style canned values provided by the test."""

    def __init__(self, persistent=False):
        self.canned = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources, output_gradients=None):
        return list(self.canned)


class _Callback:
    def __init__(self):
        self.model, self.params = None, {}

    def set_model(self, m):
        self.model = m


class _BN:
    def __init__(self, **kw):
        self.name = kw.get('name', 'bn')


class _Backend:
    @staticmethod
    def get_value(v):
        return v.numpy() if isinstance(v, Tensor) else v

    @staticmethod
    def set_value(v, x):
        v.assign(x)


keras = types.SimpleNamespace(callbacks=types.SimpleNamespace(Callback=_Callback),
                              layers=types.SimpleNamespace(BatchNormalization=_BN), backend=_Backend,
                              optimizers=types.SimpleNamespace(Optimizer=object), models=types.SimpleNamespace())
experimental = types.SimpleNamespace(dlpack=None)


class _LegacyOptimizer:
    """tf.compat.v1.train.Optimizer stand-in: `canned` gradients instead of differentiation, plain SGD in apply_gradients."""

    def __init__(self, use_locking=False, name='Optimizer'):
        self._use_locking, self._name = use_locking, name
        self.canned = None
        self.applied = 0

    def compute_gradients(self, loss, var_list=None, **kwargs):
        return list(zip(self.canned, var_list))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        for g, v in grads_and_vars:
            if g is not None:
                v.assign(v - Tensor(_np.asarray(g.numpy()) * self._lr))
        self.applied += 1
        return self.applied

    def get_slot(self, var, name):
        return None

    def get_slot_names(self):
        return ['momentum']

    def variables(self):
        return []


class _GradientDescentOptimizer(_LegacyOptimizer):
    def __init__(self, learning_rate, use_locking=False, name='GradientDescent'):
        super().__init__(use_locking, name)
        self._lr = learning_rate


compat = types.SimpleNamespace(v1=types.SimpleNamespace(global_variables=lambda: [], train=types.SimpleNamespace(
    Optimizer=_LegacyOptimizer, GradientDescentOptimizer=_GradientDescentOptimizer)))
