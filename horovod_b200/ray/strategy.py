"""How workers are packed onto nodes (role parity: horovod/ray/strategy.py): a *colocated* strategy asks for
`num_hosts` bundles of `num_workers_per_host` workers (placement group, STRICT_SPREAD over hosts), a *pack* strategy
asks for `num_workers` single-worker bundles packed as tightly as possible."""


def colocated_bundles(num_hosts, num_workers_per_host, cpus_per_worker=1, gpus_per_worker=0):
    bundle = {'CPU': cpus_per_worker * num_workers_per_host}
    if gpus_per_worker:
        bundle['GPU'] = gpus_per_worker * num_workers_per_host
    return [dict(bundle) for _ in range(num_hosts)], 'STRICT_SPREAD'


def pack_bundles(num_workers, cpus_per_worker=1, gpus_per_worker=0):
    bundle = {'CPU': cpus_per_worker}
    if gpus_per_worker:
        bundle['GPU'] = gpus_per_worker
    return [dict(bundle) for _ in range(num_workers)], 'PACK'


def assign_ranks(worker_hostnames):
    """worker index -> env dict (HOROVOD_RANK/SIZE/LOCAL_*/CROSS_*), hosts ordered by first appearance
    (the Coordinator of the reference, ray/runner.py:45-131)."""
    size = len(worker_hostnames)
    hosts = []
    for h in worker_hostnames:
        if h not in hosts:
            hosts.append(h)
    per_host = {h: [i for i, x in enumerate(worker_hostnames) if x == h] for h in hosts}
    envs = [None] * size
    rank = 0
    for cross_idx, h in enumerate(hosts):
        for local_rank, widx in enumerate(per_host[h]):
            cross_size = sum(1 for hh in hosts if len(per_host[hh]) > local_rank)
            cross_rank = sum(1 for hh in hosts[:cross_idx] if len(per_host[hh]) > local_rank)
            envs[widx] = {'HOROVOD_HOSTNAME': h, 'HOROVOD_RANK': str(rank), 'HOROVOD_SIZE': str(size),
                          'HOROVOD_LOCAL_RANK': str(local_rank), 'HOROVOD_LOCAL_SIZE': str(len(per_host[h])),
                          'HOROVOD_CROSS_RANK': str(cross_rank), 'HOROVOD_CROSS_SIZE': str(cross_size)}
            rank += 1
    return envs
