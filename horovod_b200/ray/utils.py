"""Helpers of the Ray integration (reference horovod/ray/utils.py: `detect_nics` :36-80, `nics_to_env_var` :82-87,
`map_blocking` :90-92)."""
import socket


def map_blocking(fn, collection):
    """ray.get over fn(x) for x in collection (fn returns an ObjectRef)."""
    import ray
    return ray.get([fn(w) for w in collection])


def nics_to_env_var(nics):
    """Environment that pins the rendezvous / host data plane and NCCL to `nics`."""
    nics = sorted(nics)
    return {'HOROVOD_GLOO_IFACE': nics[0], 'NCCL_SOCKET_IFNAME': ','.join(nics)}


def _interfaces():
    """name -> IPv4 addresses of this machine."""
    try:
        import psutil
        out = {}
        for name, addrs in psutil.net_if_addrs().items():
            v4 = [a.address for a in addrs if a.family == socket.AF_INET]
            if v4:
                out[name] = v4
        return out
    except ImportError:  # pragma: no cover
        return {'lo': ['127.0.0.1']}


def detect_nics(settings, all_host_names, node_workers=None, call=None):
    """Interfaces every node of the job has (by name), loopback excluded when the job spans several nodes.

    `settings.nics` wins when given.  Otherwise one worker per node reports its interfaces and the intersection is taken —
    the reference starts its task servers inside the actors and lets them probe each other (driver_service.py); interface
    names are enough here because the rendezvous address each rank advertises is derived from the chosen interface
    (`HOROVOD_GLOO_IFACE`, see csrc/common/engine.cc) and every peer dials that address directly.

    `call(worker, fn)` runs fn on a worker and returns its result (default: `ray.get(worker.execute.remote(fn))`)."""
    if getattr(settings, 'nics', None):
        return set(settings.nics)
    hosts = list(dict.fromkeys(all_host_names))
    if node_workers is None or len(hosts) <= 1:
        local = _interfaces()
        names = {n for n, addrs in local.items() if len(hosts) <= 1 or not all(a.startswith('127.') for a in addrs)}
        return names or set(local)
    if call is None:
        def call(worker, fn):
            import ray
            return ray.get(worker.execute.remote(fn))
    common = None
    for w in node_workers:
        table = call(w, _interfaces)
        names = {n for n, addrs in table.items() if not all(a.startswith('127.') for a in addrs)}
        common = names if common is None else common & names
    if not common:
        raise RuntimeError('Unable to find a set of common task-to-task communication interfaces: the nodes %s share no '
                           'non-loopback interface name. Pass nics= to create_settings().' % hosts)
    return common
