"""Spark integration (role parity: horovod/spark): `horovod_b200.spark.run(fn, ...)` runs a training function inside
Spark tasks. pyspark is not part of this image; everything that needs it is imported lazily and raises a clear error."""


def _require_pyspark():
    try:
        import pyspark  # noqa: F401
    except ImportError as e:
        raise ImportError('horovod_b200.spark requires pyspark, which is not installed in this environment') from e


def run(fn, args=(), kwargs=None, num_proc=None, start_timeout=None, env=None, stdout=None, stderr=None, verbose=1, nics=None):
    """Runs `fn` on `num_proc` Spark tasks, each becoming one rank (reference spark/runner.py:200-310)."""
    _require_pyspark()
    from horovod_b200.spark.runner import run as _run
    return _run(fn, args=args, kwargs=kwargs or {}, num_proc=num_proc, start_timeout=start_timeout, env=env, stdout=stdout,
                stderr=stderr, verbose=verbose, nics=nics)
