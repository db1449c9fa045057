"""Local network helpers (reference runner/util/network.py)."""
import socket

import psutil


def get_local_host_addresses():
    return [a.address for addrs in psutil.net_if_addrs().values() for a in addrs if a.family == socket.AF_INET]


def get_local_intfs():
    return set(psutil.net_if_addrs().keys())


def resolve_host_address(host_name):
    try:
        return socket.gethostbyname(host_name)
    except socket.gaierror:
        return None


def filter_local_addresses(all_host_names):
    """Host names that do NOT refer to this machine."""
    local = set(get_local_host_addresses()) | {'127.0.0.1'}
    return [h for h in all_host_names if resolve_host_address(h) not in local and h not in ('localhost', socket.gethostname())]


def is_local_host(host_name):
    return not filter_local_addresses([host_name])


def find_port(server_factory=None):
    if server_factory is None:
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.bind(('', 0))
        port = s.getsockname()[1]
        s.close()
        return port
    from horovod_b200.runner.common.util.network import find_port as _fp
    return _fp(server_factory)


def get_driver_ip(nics):
    """IPv4 address of the first requested NIC (or the first non-loopback one)."""
    for iface, addrs in psutil.net_if_addrs().items():
        if nics and iface not in nics:
            continue
        for a in addrs:
            if a.family == socket.AF_INET and (nics or not a.address.startswith('127.')):
                return a.address
    return '127.0.0.1'
