"""Pre-flight NIC probe, driver side: find the network interfaces over which ALL hosts of a job can reach each other,
so that the mesh transport (and NCCL_SOCKET_IFNAME) can be pinned to them.

Protocol (HMAC-authenticated RPC of runner/common/util/network.py):
  1. the launcher starts a `ProbeCoordinator` and one probe agent per host (`python -m horovod_b200.runner.task.task_service`,
     through ssh for remote hosts);
  2. every agent opens a listener on each of its IPv4 interfaces and `Enroll`s {interface: [(ip, port)]} with the coordinator;
  3. once all hosts enrolled, agent i asks for the address table of agent (i+1) % n (`NextPeer`), tries every
     (interface, address) of it — accepting only connections that arrive through the SAME-named interface, which filters
     NAT'ed / docker bridges — and `Report`s the interfaces that worked;
  4. the coordinator intersects the reports: those interfaces are routable around the whole ring.

Capability parity: horovod/runner/driver/driver_service.py (get_common_interfaces, _launch_task_servers,
_run_probe) + runner/common/service/{driver,task}_service.py — with a three-message protocol instead of the reference's
driver/task service class hierarchy.
"""
import shlex
import sys
import threading

from horovod_b200.runner.common.util import codec, network, safe_shell_exec, timeout as timeout_util
from horovod_b200.runner.util import network as net_util
from horovod_b200.runner.util import threads


class Enroll(object):
    def __init__(self, index, addresses, host_id):
        self.index, self.addresses, self.host_id = index, addresses, host_id


class NextPeer(object):
    """Agent -> coordinator: give me the addresses of the agent after me (blocks until everyone enrolled)."""

    def __init__(self, index):
        self.index = index


class NextPeerReply(object):
    def __init__(self, peer_index, addresses):
        self.peer_index, self.addresses = peer_index, addresses


class Report(object):
    def __init__(self, index, reachable_interfaces):
        self.index, self.reachable_interfaces = index, reachable_interfaces


class ProbeCoordinator(network.BasicService):
    NAME = 'hvd nic probe coordinator'

    def __init__(self, num_hosts, key, nic=None):
        self._n = num_hosts
        self._cv = threading.Condition()
        self._enrolled = {}
        self._reports = {}
        super(ProbeCoordinator, self).__init__(ProbeCoordinator.NAME, key, nic)

    def _handle(self, req, client_address):
        if isinstance(req, Enroll):
            with self._cv:
                self._enrolled[req.index] = (req.addresses, req.host_id)
                self._cv.notify_all()
            return network.AckResponse()
        if isinstance(req, NextPeer):
            with self._cv:
                self._cv.wait_for(lambda: len(self._enrolled) == self._n, timeout=120)
                peer = (req.index + 1) % self._n
                return NextPeerReply(peer, self._enrolled.get(peer, ({}, None))[0])
        if isinstance(req, Report):
            with self._cv:
                self._reports[req.index] = set(req.reachable_interfaces)
                self._cv.notify_all()
            return network.AckResponse()
        return super(ProbeCoordinator, self)._handle(req, client_address)

    def wait_for_reports(self, deadline):
        with self._cv:
            while len(self._reports) < self._n:
                self._cv.wait(min(1.0, max(0.01, deadline.remaining())))
                deadline.check_time_out_for('all hosts to finish the interface probe')
            return dict(self._reports)

    def host_ids(self):
        with self._cv:
            return {i: v[1] for i, v in self._enrolled.items()}


class ProbeCoordinatorClient(network.BasicClient):
    def __init__(self, addresses, key, verbose=0):
        super(ProbeCoordinatorClient, self).__init__(ProbeCoordinator.NAME, addresses, key, verbose)

    def enroll(self, index, addresses, host_id):
        self._send(Enroll(index, addresses, host_id))

    def next_peer(self, index):
        return self._send(NextPeer(index))

    def report(self, index, interfaces):
        self._send(Report(index, sorted(interfaces)))


def _start_agents(all_host_names, local_host_names, coordinator_addresses, settings):
    from horovod_b200.runner.mesh_run import get_ssh_command
    n = len(all_host_names)

    def run(cmd):
        rc = safe_shell_exec.execute(cmd)
        if rc != 0:
            print(f'nic probe agent failed with exit code {rc}: {cmd}', file=sys.stderr)
        return rc

    jobs = []
    for i, host in enumerate(all_host_names):
        cmd = ' '.join([shlex.quote(sys.executable), '-m', 'horovod_b200.runner.task.task_service', codec.dumps_base64(i),
                        codec.dumps_base64(n), codec.dumps_base64(coordinator_addresses), codec.dumps_base64(settings.key),
                        codec.dumps_base64(sorted(settings.nics) if settings.nics else None)])
        if host not in local_host_names:
            cmd = get_ssh_command(cmd, host=host, port=settings.ssh_port, identity_file=settings.ssh_identity_file)
        jobs.append([cmd])
    threads.execute_function_multithreaded(run, jobs, block_until_all_done=False)


def _loopback_interfaces(restrict_to=None):
    import socket
    import psutil
    out = set()
    for iface, addrs in psutil.net_if_addrs().items():
        if restrict_to and iface not in restrict_to:
            continue
        if any(a.family == socket.AF_INET and a.address == '127.0.0.1' for a in addrs):
            out.add(iface)
    return out


def get_common_interfaces(settings, all_host_names, remote_host_names=None, fn_cache=None):
    """Interfaces usable between all hosts. All-local jobs short-circuit to the loopback interface(s)."""
    if remote_host_names is None:
        remote_host_names = net_util.filter_local_addresses(all_host_names)
    if not remote_host_names:
        nics = _loopback_interfaces(settings.nics)
        if not nics:
            raise ValueError('No interface is found for address 127.0.0.1.')
        return nics
    if settings.nics:
        return settings.nics  # the user pinned the interfaces: trust them
    deadline = settings.start_timeout if isinstance(settings.start_timeout, timeout_util.Timeout) else timeout_util.Timeout(
        settings.start_timeout or 30, message='Timed out waiting for {activity}. Please check connectivity between servers.')
    coordinator = ProbeCoordinator(len(all_host_names), settings.key)
    try:
        _start_agents(all_host_names, set(all_host_names) - set(remote_host_names), coordinator.addresses(), settings)
        reports = coordinator.wait_for_reports(deadline)
    finally:
        coordinator.shutdown()
    common = set.intersection(*reports.values()) if reports else set()
    if not common:
        raise Exception('Unable to find a set of common task-to-task communication interfaces: %s' % reports)
    return common
