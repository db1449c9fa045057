"""Where an estimator's training function runs (parity: horovod/spark/common/backend.py `Backend`/`SparkBackend`)."""
import os


class Backend:
    def run(self, fn, args=(), kwargs=None, env=None):
        raise NotImplementedError

    def num_processes(self):
        raise NotImplementedError


def default_num_proc():
    """The active Spark context's default parallelism (what SparkBackend uses when num_proc is not given)."""
    import pyspark
    ctx = pyspark.SparkContext._active_spark_context
    if ctx is None:
        raise RuntimeError('Could not find an active SparkContext, are you running in a PySpark session?')
    return ctx.defaultParallelism


class SparkBackend(Backend):
    """Spark barrier tasks through `horovod_b200.spark.run`."""

    def __init__(self, num_proc=None, env=None, **run_kwargs):
        self._num_proc, self._env, self._kw = num_proc, env, run_kwargs

    def num_processes(self):
        if self._num_proc is None:
            import pyspark
            self._num_proc = pyspark.SparkContext._active_spark_context.defaultParallelism
        return self._num_proc

    def run(self, fn, args=(), kwargs=None, env=None):
        import horovod_b200.spark as hs
        full_env = dict(os.environ if self._env is None else self._env)
        full_env.update(env or {})
        return hs.run(fn, args=args, kwargs=kwargs, num_proc=self.num_processes(), env=full_env, **self._kw)


class LocalBackend(Backend):
    """`num_proc` processes on this machine through `horovod_b200.run` (hvdrun's programmatic entry point): the estimators
    work without a Spark cluster, e.g. on one 8-GPU box."""

    def __init__(self, num_proc=1, env=None, **run_kwargs):
        self._num_proc, self._env, self._kw = num_proc, env, run_kwargs

    def num_processes(self):
        return self._num_proc

    def run(self, fn, args=(), kwargs=None, env=None):
        import horovod_b200
        extra = dict(self._env or {})
        extra.update(env or {})
        saved = {k: os.environ.get(k) for k in extra}
        os.environ.update({k: str(v) for k, v in extra.items()})  # the launcher forwards the driver's environment
        try:
            return horovod_b200.run(fn, args=args, kwargs=kwargs, num_proc=self._num_proc, **self._kw)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
