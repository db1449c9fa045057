"""tf.data service workers as a Spark job:  spark-submit ... -m horovod_b200.spark.tensorflow.compute_worker /shared/compute.json

The Spark driver hosts the ComputeService (the notice board dispatchers and workers register with) and writes its config to
`configfile`; one tf.data worker then runs in every Spark task (`spark.default.parallelism` of them) through
`horovod_b200.spark.run`.  The training job reads the config with `TfDataServiceConfig.read(configfile,
wait_for_file_creation=True)`.  Role parity: horovod/spark/tensorflow/compute_worker.py; the hvdrun-launched flavour is
`horovod_b200.tensorflow.data.compute_worker`.
"""
import argparse
import logging
import signal
import sys


def plan(workers, dispatchers):
    """-> workers per dispatcher; the task count must divide evenly."""
    if dispatchers < 1 or workers % dispatchers:
        raise ValueError('Number of processes (%d) must be a multiple of number of dispatchers (%d).' % (workers, dispatchers))
    return workers // dispatchers


def main(configfile, dispatchers=1, dispatcher_side='compute', timeout=60, spark_context=None, run=None, workers=None):
    """`spark_context` / `run` / `workers` are injectable for tests; by default the active SparkSession and
    `horovod_b200.spark.run` are used."""
    from horovod_b200.runner.common.service.compute_service import ComputeService
    from horovod_b200.runner.common.util import secret
    from horovod_b200.tensorflow.data.compute_service import TfDataServiceConfig, compute_worker_fn
    spark = None
    if run is None:
        from pyspark.sql import SparkSession
        from horovod_b200.spark import run
        spark = SparkSession.builder.getOrCreate()
        spark_context = spark.sparkContext
    if workers is None:
        workers = spark_context.defaultParallelism
    per_dispatcher = plan(workers, dispatchers)
    key = secret.make_secret_key()
    compute = ComputeService(dispatchers, per_dispatcher, key=key)
    try:
        config = TfDataServiceConfig(dispatchers=dispatchers, workers_per_dispatcher=per_dispatcher, dispatcher_side=dispatcher_side,
                                     addresses=compute.addresses(), key=key, timeout=timeout)
        config.write(configfile)

        def stop(*_):
            logging.info('compute worker driver received SIGTERM: stopping the Spark context')
            if spark_context is not None and hasattr(spark_context, 'stop'):
                spark_context.stop()
        try:
            signal.signal(signal.SIGTERM, stop)
        except ValueError:           # not the main thread (tests)
            pass
        kwargs = dict(args=(config,), num_proc=workers, verbose=2)
        if spark_context is not None:
            kwargs['spark_context'] = spark_context
        return run(compute_worker_fn, **kwargs)
    finally:
        compute.shutdown()
        if spark is not None:
            spark.stop()


def parse_args(argv=None):
    p = argparse.ArgumentParser(description='tf.data service workers as a Spark job')
    p.add_argument('configfile', help='where the driver writes the compute service config')
    p.add_argument('--dispatchers', type=int, default=1, help='number of dispatchers (the task count must be a multiple)')
    p.add_argument('--dispatcher-side', default='compute', choices=['compute', 'training'], help='which job hosts the dispatchers')
    p.add_argument('--timeout', type=int, default=60, help='seconds to wait for registrations')
    return p.parse_args(argv)


if __name__ == '__main__':
    a = parse_args()
    main(a.configfile, a.dispatchers, a.dispatcher_side, a.timeout)
    sys.exit(0)
