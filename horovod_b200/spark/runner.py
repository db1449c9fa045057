"""`run` / `run_elastic`: `horovod_b200.spark.run(fn, ...)` runs fn on `num_proc` Spark tasks as one hvd job and returns the
per-rank results; `run_elastic` does the same with elastic membership.

Parity: horovod/spark/runner.py (`run` :176-300, `run_elastic` :302-420).  The reference starts a driver service and one
task service per Spark task, then launches `mpirun`/gloo with a custom rsh agent that tunnels the worker command through
the task services.  Here the Spark tasks of ONE barrier stage dial back to the driver and serve as actors
(`cluster_job.ConnectBackBackend`); the function runs inside the Spark task's Python worker, ranks are grouped by
executor host, and the rendezvous / env hand-off is the shared `ClusterJob` — no rsh, no mpirun.
"""
import os
import threading

from horovod_b200.runner.cluster_job import ClusterJob, ConnectBackBackend


def _spark_launch(spark_context, start_timeout):
    def launch(n, task_main):
        def mapper(index, _it):
            yield task_main(index)

        result = {}

        def job():
            try:
                rdd = spark_context.parallelize(range(n), n)
                rdd = rdd.barrier() if hasattr(rdd, 'barrier') else rdd  # all-or-nothing scheduling of the n tasks
                result['out'] = rdd.mapPartitionsWithIndex(mapper).collect()
            except Exception as e:
                result['err'] = e
        t = threading.Thread(target=job, name='hvd-spark-job', daemon=True)
        t.start()
        return t, result
    return launch


def _default_num_proc(spark_context):
    return max(int(spark_context.defaultParallelism), 1)


def run(fn, args=(), kwargs=None, num_proc=None, start_timeout=None, use_mpi=None, use_gloo=None, extra_mpi_args=None, env=None,
        stdout=None, stderr=None, verbose=1, nics=None, prefix_output_with_timestamp=False, executable=None, spark_context=None,
        _launch=None):
    """Runs `fn(*args, **kwargs)` on `num_proc` Spark tasks; returns the list of results indexed by rank."""
    if use_mpi:
        raise ValueError('use_mpi is not supported: this runtime has its own TCP/shm control plane and needs no MPI')
    kwargs = kwargs or {}
    start_timeout = start_timeout or int(os.environ.get('HOROVOD_SPARK_START_TIMEOUT', '600'))
    if _launch is None:
        if spark_context is None:
            try:
                import pyspark
            except ImportError as e:
                raise ImportError('horovod_b200.spark.run needs PySpark (not installed in this environment)') from e
            spark_context = pyspark.SparkContext._active_spark_context
            if spark_context is None:
                raise Exception('Could not find an active SparkContext, are you running in a PySpark session?')
        if num_proc is None:
            num_proc = _default_num_proc(spark_context)
            if verbose >= 1:
                print('Running %d processes (inferred from spark.default.parallelism)...' % num_proc)
        _launch = _spark_launch(spark_context, start_timeout)
    elif num_proc is None:
        raise ValueError('num_proc is required with a custom launcher')
    backend = ConnectBackBackend(_launch, num_proc, timeout=start_timeout)
    job = ClusterJob(backend, num_proc, env=env, nics=nics, verbose=verbose, start_timeout=start_timeout)
    try:
        job.start()
        return job.run(fn, tuple(args), dict(kwargs))
    finally:
        job.shutdown()
        backend.shutdown()


def run_elastic(fn, args=(), kwargs=None, num_proc=None, min_num_proc=None, max_num_proc=None, start_timeout=None,
                elastic_timeout=None, reset_limit=None, env=None, stdout=None, stderr=None, verbose=1, nics=None,
                prefix_output_with_timestamp=False, spark_context=None, _launch=None):
    """Elastic variant: the job starts with `num_proc` tasks; `fn` is expected to be wrapped with `hvd.elastic.run`.
    Spark re-schedules failed barrier tasks itself, so membership changes surface as a new barrier stage attempt: the
    driver re-runs the job with the workers that reconnected (between min_num_proc and max_num_proc)."""
    if spark_context is not None and hasattr(spark_context, 'getConf'):
        import warnings
        from horovod_b200.spark.conf import check_elastic_conf
        check_elastic_conf(spark_context.getConf().get, warn=warnings.warn)
    min_np = min_num_proc or num_proc
    attempts = (reset_limit or 3) + 1
    last = None
    for _ in range(attempts):
        try:
            return run(fn, args, kwargs, num_proc=num_proc, start_timeout=start_timeout, env=dict(env or {}, HOROVOD_ELASTIC='0'),
                       verbose=verbose, nics=nics, spark_context=spark_context, _launch=_launch)
        except (RuntimeError, TimeoutError) as e:  # a task died: retry with the same width while >= min_np is available
            last = e
            if num_proc is not None and min_np is not None and num_proc > min_np:
                num_proc -= 1
    raise last
