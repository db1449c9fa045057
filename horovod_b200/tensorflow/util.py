"""Small TensorFlow helpers shared by the front-end modules (reference horovod/tensorflow/util.py:22-57)."""
import functools

import tensorflow as tf


def _executing_eagerly():
    """True outside tf.function graphs (TF1 builds without `executing_eagerly` count as graph mode)."""
    fn = getattr(tf, 'executing_eagerly', None)
    return bool(fn()) if fn is not None else False


def _make_subgraph(f):
    """Traces `f` as a tf.function where TensorFlow has one (the reference falls back to `tf.contrib.eager.defun`)."""
    wrap = getattr(tf, 'function', None)
    return wrap(f) if wrap is not None else f


def _cache(f):
    """Memoises `f` on its positional arguments; unhashable arguments (lists of variables) are keyed by identity."""
    memo = {}

    def key_of(a):
        try:
            hash(a)
            return a
        except TypeError:
            return ('id', id(a))

    @functools.wraps(f)
    def cached(*args):
        k = tuple(key_of(a) for a in args)
        if k not in memo:
            memo[k] = f(*args)
        return memo[k]
    return cached


def vars_to_refs(vars):
    """tf.Variables are unhashable under TF2: hand back their `.ref()` handles (lists and tuples element-wise)."""
    if isinstance(vars, (list, tuple)):
        return type(vars)(vars_to_refs(v) for v in vars)
    return vars.ref() if hasattr(vars, 'ref') else vars


def refs_to_vars(refs):
    """Inverse of `vars_to_refs`."""
    if isinstance(refs, (list, tuple)):
        return type(refs)(refs_to_vars(r) for r in refs)
    return refs.deref() if hasattr(refs, 'deref') else refs
