"""Spark runner: every Spark task of a barrier stage becomes one rank.  The Spark driver hosts the rendezvous KV server;
tasks learn their rank from their partition id, group by host to derive local_rank, export the HOROVOD_* environment and
call the user function (which calls hvd.init()).  Role parity: horovod/spark/runner.py (without the mpirun/rsh agent
path: the native mesh transport needs no external launcher)."""
import os
import socket

import cloudpickle

from horovod_b200.runner.common.util import hosts as hosts_util
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.util import network


def _task_fn(index, driver_addr, driver_port, num_proc, fn_bytes, extra_env):
    """Body of one Spark task."""
    from pyspark import BarrierTaskContext
    ctx = BarrierTaskContext.get()
    host = socket.gethostname()
    # exchange (partition, host) through the barrier so that every task can compute the layout
    infos = ctx.allGather(f'{index}:{host}')
    by_index = dict((int(i.split(':')[0]), i.split(':', 1)[1]) for i in infos)
    order = []
    for i in range(num_proc):
        if by_index[i] not in order:
            order.append(by_index[i])
    slots = {h: [i for i in range(num_proc) if by_index[i] == h] for h in order}
    layout = hosts_util.get_host_assignments([hosts_util.HostInfo(h, len(slots[h])) for h in order], num_proc)
    # ranks are laid out host by host; this task's position inside its host decides its slot
    pos = slots[host].index(index)
    mine = [s for s in layout if s.hostname == host][pos]
    os.environ.update({'HOROVOD_HOSTNAME': host, 'HOROVOD_RANK': str(mine.rank), 'HOROVOD_SIZE': str(mine.size),
                       'HOROVOD_LOCAL_RANK': str(mine.local_rank), 'HOROVOD_LOCAL_SIZE': str(mine.local_size),
                       'HOROVOD_CROSS_RANK': str(mine.cross_rank), 'HOROVOD_CROSS_SIZE': str(mine.cross_size),
                       'HOROVOD_GLOO_RENDEZVOUS_ADDR': driver_addr, 'HOROVOD_GLOO_RENDEZVOUS_PORT': str(driver_port)})
    os.environ.update(extra_env or {})
    fn, args, kwargs = cloudpickle.loads(fn_bytes)
    return mine.rank, fn(*args, **kwargs)


def run(fn, args=(), kwargs=None, num_proc=None, start_timeout=None, env=None, stdout=None, stderr=None, verbose=1, nics=None):
    import pyspark
    spark_context = pyspark.SparkContext._active_spark_context
    if spark_context is None:
        raise Exception('Could not find an active SparkContext, are you running in a PySpark session?')
    if num_proc is None:
        num_proc = spark_context.defaultParallelism
    server = RendezvousServer(verbose)
    port = server.start_server()
    addr = network.get_driver_ip(set(nics) if nics else None)
    fn_bytes = cloudpickle.dumps((fn, args, kwargs or {}))
    try:
        rdd = spark_context.parallelize(range(num_proc), numSlices=num_proc).barrier()
        results = rdd.mapPartitionsWithIndex(
            lambda index, _: [_task_fn(index, addr, port, num_proc, fn_bytes, env)]).collect()
    finally:
        server.stop()
    return [r for _, r in sorted(results)]
