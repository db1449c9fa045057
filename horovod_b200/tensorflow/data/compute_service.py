"""tf.data service on top of the compute service: training ranks stream batches from dispatcher / worker servers that run
in a separate hvd job.

Role parity: horovod/tensorflow/data/compute_service.py (`TfDataServiceConfig`, `tf_data_service`,
`send_to_data_service` (also patched onto tf.data.Dataset), `compute_worker_fn`).

The TensorFlow servers are reached through `_service_api()`, so the orchestration can be exercised with stand-ins; pass
`servers=` explicitly to do that.
"""
import binascii
import contextlib
import json
import os
import tempfile
import time

from horovod_b200.runner.common.service.compute_service import ComputeClient

_SIDES = ('compute', 'training')


class TfDataServiceConfig:
    """Immutable description of a running compute service; travels as JSON through a file both jobs can read."""
    __slots__ = ('dispatchers', 'workers_per_dispatcher', 'dispatcher_side', 'addresses', 'key', 'timeout')

    def __init__(self, dispatchers, workers_per_dispatcher, dispatcher_side, addresses, key, timeout=60):
        if dispatcher_side not in _SIDES:
            raise ValueError('dispatcher_side must be one of %s: %r' % (_SIDES, dispatcher_side))
        for name, value in zip(self.__slots__, (dispatchers, workers_per_dispatcher, dispatcher_side, addresses, key, timeout)):
            object.__setattr__(self, name, value)

    def __setattr__(self, name, value):
        raise AttributeError('TfDataServiceConfig is read-only')

    def __eq__(self, other):
        return isinstance(other, TfDataServiceConfig) and self.to_dict() == other.to_dict()

    def __reduce__(self):
        return (TfDataServiceConfig.from_dict, (self.to_dict(),))

    def compute_client(self, verbose=1):
        return ComputeClient(self.addresses, self.key, verbose=verbose)

    def to_dict(self):
        d = {name: getattr(self, name) for name in self.__slots__}
        d['key'] = binascii.hexlify(self.key).decode()
        d['addresses'] = {intf: [list(a) for a in addrs] for intf, addrs in self.addresses.items()}
        return d

    @staticmethod
    def from_dict(d):
        return TfDataServiceConfig(dispatchers=d['dispatchers'], workers_per_dispatcher=d['workers_per_dispatcher'],
                                   dispatcher_side=d['dispatcher_side'],
                                   addresses={intf: [(a[0], a[1]) for a in addrs] for intf, addrs in d['addresses'].items()},
                                   key=binascii.unhexlify(d['key']), timeout=d.get('timeout', 60))

    def write(self, filename):
        """Atomic: readers polling for the file never see a partial config."""
        directory = os.path.dirname(os.path.abspath(filename))
        fd, tmp = tempfile.mkstemp(dir=directory, prefix=os.path.basename(filename) + '.')
        with os.fdopen(fd, 'w') as f:
            json.dump(self.to_dict(), f)
        os.replace(tmp, filename)

    @staticmethod
    def read(filename, wait_for_file_creation=False, poll_seconds=0.5, timeout=None):
        deadline = None if timeout is None else time.monotonic() + timeout
        while wait_for_file_creation and not os.path.exists(filename):
            if deadline is not None and time.monotonic() > deadline:
                raise TimeoutError('%s did not appear within %s s' % (filename, timeout))
            time.sleep(poll_seconds)
        with open(filename) as f:
            return TfDataServiceConfig.from_dict(json.load(f))


def _service_api():
    import tensorflow as tf
    return tf.data.experimental.service


def _stop(server):
    # the TF servers have no public stop(): `_stop` + `join` is what the TF tests use
    getattr(server, '_stop', getattr(server, 'stop', lambda: None))()
    if hasattr(server, 'join'):
        server.join()


@contextlib.contextmanager
def tf_data_service(compute_config, rank, servers=None):
    """Yields the address of the dispatcher this training rank talks to (starting it first when dispatchers live on the
    training side: one per rank, or rank 0's for everybody when there is a single dispatcher)."""
    api = servers or _service_api()
    compute = compute_config.compute_client(verbose=2)
    mine = None
    if compute_config.dispatcher_side == 'training' and (compute_config.dispatchers > 1 or rank == 0):
        mine = api.DispatchServer()
        compute.register_dispatcher(rank if compute_config.dispatchers > 1 else 0, mine.target)
    dispatcher_id = rank if compute_config.dispatchers > 1 else 0
    address = compute.wait_for_dispatcher_registration(dispatcher_id, compute_config.timeout)
    compute.wait_for_dispatcher_worker_registration(dispatcher_id, compute_config.timeout)
    try:
        yield address
    finally:
        if mine is not None:
            _stop(mine)


def send_to_data_service(dataset, compute_config, rank, size=None, processing_mode='distributed_epoch', reuse_dataset=False,
                         round_robin=False, servers=None):
    """dataset -> the same dataset produced by the data-service workers.  `reuse_dataset` shares one job between the
    ranks (each element goes to exactly one rank); `round_robin` additionally makes the hand-out deterministic."""
    if compute_config.dispatcher_side == 'training':
        raise RuntimeError('training side dispatcher not supported, use tf_data_service context manager instead')
    api = servers or _service_api()
    with tf_data_service(compute_config, rank, servers=api) as address:
        shared, ordered = reuse_dataset, reuse_dataset and round_robin
        return dataset.apply(api.distribute(processing_mode=processing_mode, service=address, job_name='job' if shared else None,
                                            consumer_index=rank if ordered else None, num_consumers=size if ordered else None))


def _patch_dataset():
    try:
        import tensorflow as tf
        tf.data.Dataset.send_to_data_service = send_to_data_service
    except Exception:  # noqa: BLE001 - TensorFlow absent or a stand-in without tf.data
        pass


_patch_dataset()


def compute_worker_fn(compute_config, rank=None, servers=None, wait_for_shutdown=True):
    """Runs on every rank of the compute job: worker `rank` serves dispatcher `rank // workers_per_dispatcher`; the first
    worker of each group also hosts the dispatcher when dispatchers live on the compute side.  Returns when the training
    side posts the shutdown."""
    if rank is None:
        import horovod_b200.tensorflow as hvd
        hvd.init()
        rank = hvd.rank()
    api = servers or _service_api()
    group, first_of_group = divmod(rank, compute_config.workers_per_dispatcher)
    compute = compute_config.compute_client(verbose=2)
    dispatcher = None
    if compute_config.dispatcher_side == 'compute' and first_of_group == 0:
        dispatcher = api.DispatchServer()
        compute.register_dispatcher(group, dispatcher.target)
    address = compute.wait_for_dispatcher_registration(group, compute_config.timeout)
    worker = api.WorkerServer(api.WorkerConfig(dispatcher_address=address.split('://', 1)[-1], heartbeat_interval_ms=1000,
                                               dispatcher_timeout_ms=compute_config.timeout * 1000))
    if hasattr(worker, 'start'):
        worker.start()
    compute.register_worker_for_dispatcher(group, rank)
    try:
        if wait_for_shutdown:
            compute.wait_for_shutdown()
    finally:
        if wait_for_shutdown:
            _stop(worker)
            if dispatcher is not None:
                _stop(dispatcher)
    return worker, dispatcher
