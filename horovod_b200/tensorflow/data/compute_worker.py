"""Entry point of the compute job:  hvdrun -np 4 python -m horovod_b200.tensorflow.data.compute_worker /shared/compute.json

Rank 0 hosts the ComputeService and writes its config to `configfile` (the training job reads it with
`TfDataServiceConfig.read(configfile, wait_for_file_creation=True)`); every rank then runs `compute_worker_fn`.
Role parity: horovod/tensorflow/data/compute_worker.py.
"""
import argparse

from horovod_b200.runner.common.service.compute_service import ComputeService
from horovod_b200.runner.common.util import secret
from horovod_b200.tensorflow.data.compute_service import TfDataServiceConfig, compute_worker_fn


def main(dispatchers, dispatcher_side, configfile, timeout, hvd=None, servers=None):
    if hvd is None:
        import horovod_b200.tensorflow as hvd
    hvd.init()
    rank, size = hvd.rank(), hvd.size()
    if size % dispatchers:
        raise ValueError('Number of processes (%d) must be a multiple of number of dispatchers (%d).' % (size, dispatchers))
    service = None
    try:
        config = None
        if rank == 0:
            key = secret.make_secret_key()
            service = ComputeService(dispatchers, size // dispatchers, key=key)
            config = TfDataServiceConfig(dispatchers=dispatchers, workers_per_dispatcher=size // dispatchers,
                                         dispatcher_side=dispatcher_side, addresses=service.addresses(), key=key, timeout=timeout)
            config.write(configfile)
        config = hvd.broadcast_object(config, root_rank=0, name='TfDataServiceConfig')
        compute_worker_fn(config, rank=rank, servers=servers)
    finally:
        if service is not None:
            service.shutdown()


def parse_args(argv=None):
    p = argparse.ArgumentParser(description='tf.data service workers as an hvd job')
    p.add_argument('configfile', help='where rank 0 writes the compute service config')
    p.add_argument('--dispatchers', type=int, default=1, help='number of dispatchers (the job size must be a multiple)')
    p.add_argument('--dispatcher-side', default='compute', choices=['compute', 'training'], help='which job hosts the dispatchers')
    p.add_argument('--timeout', type=int, default=60, help='seconds to wait for registrations')
    return p.parse_args(argv)


if __name__ == '__main__':
    a = parse_args()
    main(a.dispatchers, a.dispatcher_side, a.configfile, a.timeout)
