"""Spark integration: `horovod_b200.spark.run(fn, ...)` runs fn on `num_proc` Spark tasks as one hvd job and returns the
per-rank results; `run_elastic` does the same with elastic membership.

Parity: horovod/spark/runner.py (`run` :176-300, `run_elastic` :302-420).  The reference starts a driver service and one
task service per Spark task, then launches `mpirun`/gloo with a custom rsh agent that tunnels the worker command through
the task services.  Here the Spark tasks of ONE barrier stage dial back to the driver and serve as actors
(`cluster_job.ConnectBackBackend`); the function runs inside the Spark task's Python worker, ranks are grouped by
executor host, and the rendezvous / env hand-off is the shared `ClusterJob` — no rsh, no mpirun.
"""
from horovod_b200.spark.runner import _default_num_proc, _spark_launch, run, run_elastic  # noqa: F401
