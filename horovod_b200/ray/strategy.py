"""How workers are packed onto nodes.

Role parity: horovod/ray/strategy.py (`ColocatedStrategy` :66-137: `num_hosts` x `num_workers_per_host` workers with a
STRICT_SPREAD placement group of one bundle per host; `PGStrategy` :139-225: `num_workers` single-worker bundles with a PACK
group).  A strategy here is a pure description — bundles, placement-group strategy, worker -> bundle index, per-worker
resources — that `RayBackend` turns into actors; nothing in this file imports ray, so the layout rules are unit-testable.
"""


def _resources(cpus, gpus):
    res = {'CPU': cpus}
    if gpus:
        res['GPU'] = gpus
    return res


class PlacementStrategy:
    """bundles: one resource dict per placement-group bundle; placement: Ray's strategy string;
    worker_bundle[i]: bundle index of worker i; worker_resources[i]: what worker i reserves inside its bundle."""
    placement = 'PACK'

    def __init__(self, cpus_per_worker=1, gpus_per_worker=0):
        if cpus_per_worker < 0 or gpus_per_worker < 0:
            raise ValueError('resources per worker must be non-negative')
        self.cpus_per_worker, self.gpus_per_worker = cpus_per_worker, gpus_per_worker

    @property
    def num_workers(self):
        raise NotImplementedError

    @property
    def bundles(self):
        raise NotImplementedError

    @property
    def worker_bundle(self):
        raise NotImplementedError

    @property
    def worker_resources(self):
        return [_resources(self.cpus_per_worker, self.gpus_per_worker) for _ in range(self.num_workers)]

    def total_resources(self):
        total = {}
        for b in self.bundles:
            for k, v in b.items():
                total[k] = total.get(k, 0) + v
        return total

    def describe(self):
        return (self.bundles, self.placement, self.worker_bundle, self.worker_resources)


class ColocatedStrategy(PlacementStrategy):
    """Exactly `num_workers_per_host` workers on each of `num_hosts` DIFFERENT nodes: one bundle per node sized for all its
    workers, bundles spread strictly (the job does not start on fewer nodes)."""
    placement = 'STRICT_SPREAD'

    def __init__(self, num_hosts, num_workers_per_host, cpus_per_worker=1, gpus_per_worker=0):
        super().__init__(cpus_per_worker, gpus_per_worker)
        if num_hosts < 1 or num_workers_per_host < 1:
            raise ValueError('num_hosts and num_workers_per_host must be >= 1')
        self.num_hosts, self.num_workers_per_host = num_hosts, num_workers_per_host

    @property
    def num_workers(self):
        return self.num_hosts * self.num_workers_per_host

    @property
    def bundles(self):
        per_host = _resources(self.cpus_per_worker * self.num_workers_per_host, self.gpus_per_worker * self.num_workers_per_host)
        return [dict(per_host) for _ in range(self.num_hosts)]

    @property
    def worker_bundle(self):
        return [i // self.num_workers_per_host for i in range(self.num_workers)]


class PackStrategy(PlacementStrategy):
    """`num_workers` workers wherever they fit, as few nodes as possible: one bundle per worker, PACK."""
    placement = 'PACK'

    def __init__(self, num_workers, cpus_per_worker=1, gpus_per_worker=0):
        super().__init__(cpus_per_worker, gpus_per_worker)
        if num_workers < 1:
            raise ValueError('num_workers must be >= 1')
        self._n = num_workers

    @property
    def num_workers(self):
        return self._n

    @property
    def bundles(self):
        return [_resources(self.cpus_per_worker, self.gpus_per_worker) for _ in range(self._n)]

    @property
    def worker_bundle(self):
        return list(range(self._n))


PGStrategy = PackStrategy      # the reference's name


def colocated_bundles(num_hosts, num_workers_per_host, cpus_per_worker=1, gpus_per_worker=0):
    s = ColocatedStrategy(num_hosts, num_workers_per_host, cpus_per_worker, gpus_per_worker)
    return s.bundles, s.placement


def pack_bundles(num_workers, cpus_per_worker=1, gpus_per_worker=0):
    s = PackStrategy(num_workers, cpus_per_worker, gpus_per_worker)
    return s.bundles, s.placement
