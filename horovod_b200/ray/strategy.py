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
