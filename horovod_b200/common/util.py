"""Small shared helpers (reference horovod/common/util.py)."""
import os
import warnings
from contextlib import contextmanager


@contextmanager
def env(**kwargs):
    """Temporarily set environment variables (None deletes)."""
    backup = {}
    for k, v in kwargs.items():
        backup[k] = os.environ.get(k)
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    try:
        yield
    finally:
        for k, v in backup.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def split_list(l, n):
    """Splits list l into n approximately even sized chunks."""
    d, r = divmod(len(l), n)
    return [l[i * d + min(i, r):(i + 1) * d + min(i + 1, r)] for i in range(n)]


def num_rank_is_power_2(num_rank):
    return num_rank != 0 and (num_rank & (num_rank - 1)) == 0


def is_iterable(x):
    try:
        iter(x)
    except TypeError:
        return False
    return True


def resolve_op(op, average, Average, Sum):
    """`op` supersedes the legacy `average=` kwarg (reference common/util.py:214-232)."""
    if op is not None:
        if average is not None:
            raise ValueError('The op parameter supersedes average. Please provide only one of them.')
        return op
    if average is not None:
        warnings.warn('Parameter `average` has been replaced with `op` and will be removed', DeprecationWarning)
        return Average if average else Sum
    return Average


def is_version_greater_equal_than(ver, target):
    from packaging import version
    return version.parse(ver) >= version.parse(target)
