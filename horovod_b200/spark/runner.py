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


def _spark_launch(spark_context, start_timeout, barrier=True):
    def launch(n, task_main):
        def mapper(index, _it):
            yield task_main(index)

        result = {}

        def job():
            try:
                rdd = spark_context.parallelize(range(n), n)
                rdd = rdd.barrier() if barrier and hasattr(rdd, 'barrier') else rdd  # all-or-nothing scheduling of the n tasks
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
                elastic_timeout=None, reset_limit=None, cooldown_range=None, env=None, stdout=None, stderr=None, verbose=1, nics=None,
                prefix_output_with_timestamp=False, spark_context=None, _launch=None):
    """Elastic variant (reference spark/runner.py `run_elastic` :312-420): `fn` is an `@hvd.elastic.run` training function.

    `max_num_proc` (default: num_proc) Spark tasks dial back to the driver and become the SLOTS of an elastic job
    (`cluster_job.ConnectBackPool`): discovery counts the live tasks per executor host, the elastic driver
    (`runner.elastic.driver.ElasticDriver`) plans ranks over them, and "spawning a worker" hands the function to an idle task
    of the requested host.  A task that dies takes its slot with it — the survivors roll back to their last commit and carry on
    as long as `min_num_proc` remain; the attempt Spark schedules for the failed task (`spark.task.maxFailures`, see
    `spark/conf.py`) dials back as a fresh slot and is picked up at the next reset.  Returns the results of the final round in
    rank order."""
    import functools
    import warnings
    from horovod_b200.ray.elastic import ElasticRayExecutor as _ElasticExecutor     # scheduler-independent despite the name
    from horovod_b200.runner.cluster_job import ConnectBackPool
    from horovod_b200.runner.elastic.discovery import HostDiscovery
    kwargs = kwargs or {}
    start_timeout = start_timeout or int(os.environ.get('HOROVOD_SPARK_START_TIMEOUT', '600'))
    if _launch is None:
        if spark_context is None:
            try:
                import pyspark
            except ImportError as e:
                raise ImportError('horovod_b200.spark.run_elastic needs PySpark (not installed in this environment)') from e
            spark_context = pyspark.SparkContext._active_spark_context
            if spark_context is None:
                raise Exception('Could not find an active SparkContext, are you running in a PySpark session?')
        if hasattr(spark_context, 'getConf'):
            from horovod_b200.spark.conf import check_elastic_conf
            check_elastic_conf(spark_context.getConf().get, warn=warnings.warn)
        if num_proc is None:
            num_proc = _default_num_proc(spark_context)
        _launch = _spark_launch(spark_context, start_timeout, barrier=False)   # all-or-nothing scheduling is what elastic avoids
    elif num_proc is None:
        raise ValueError('num_proc is required with a custom launcher')
    min_np = min_num_proc or num_proc
    max_np = max_num_proc or num_proc
    if not min_np <= num_proc <= max_np:
        raise ValueError('need min_num_proc <= num_proc <= max_num_proc, got %s <= %s <= %s' % (min_np, num_proc, max_np))
    pool = ConnectBackPool(_launch, max_np, timeout=start_timeout)

    class _PoolDiscovery(HostDiscovery):
        def find_available_hosts_and_slots(self):
            return pool.hosts_and_slots()

    try:
        pool.wait_for(min_np, start_timeout)
        try:                                      # the job may start with min_np tasks; give the rest a moment to dial in first
            pool.wait_for(num_proc, min(10.0, float(start_timeout)))
        except TimeoutError:
            pass
        extra = {'cooldown_range': cooldown_range} if cooldown_range else {}
        settings = _ElasticExecutor.create_settings(min_num_proc=min_np, max_num_proc=max_np, reset_limit=reset_limit,
                                                    elastic_timeout=elastic_timeout or 600, timeout_s=start_timeout, nics=nics, **extra)
        settings.discovery = _PoolDiscovery()
        settings.num_proc = num_proc
        settings.verbose = 2 if verbose and verbose > 1 else 0
        executor = _ElasticExecutor(settings, env_vars=dict(env or {}), override_discovery=False, actor_factory=pool.actor_factory)
        executor.start()
        return executor.run(functools.partial(fn, *tuple(args), **dict(kwargs)))
    finally:
        pool.shutdown()
