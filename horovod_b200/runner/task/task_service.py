"""Pre-flight NIC probe, host side (see runner/driver/driver_service.py for the protocol): listens on every local IPv4
interface, enrolls with the coordinator, probes the next host of the ring and reports the interfaces that worked.

Capability parity: horovod/runner/task/task_service.py + task_fn.py.
"""
import sys
import time

from horovod_b200.runner.common.util import codec, host_hash, network


class ProbeAgent(network.BasicService):
    NAME_FORMAT = 'hvd nic probe agent #%d'

    def __init__(self, index, key, nics=None):
        # BasicService binds one port on all interfaces and reports {iface: [(ip, port)]}; optionally restricted
        super(ProbeAgent, self).__init__(ProbeAgent.NAME_FORMAT % index, key, None)
        if nics:
            self._addresses = {i: a for i, a in self._addresses.items() if i in nics}


def probe(index, num_hosts, coordinator_addresses, key, nics=None, linger_s=2.0):
    from horovod_b200.runner.driver.driver_service import ProbeCoordinatorClient
    agent = ProbeAgent(index, key, nics)
    try:
        coordinator = ProbeCoordinatorClient(coordinator_addresses, key)
        coordinator.enroll(index, agent.addresses(), host_hash.host_hash())
        reply = coordinator.next_peer(index)
        reachable = set()
        try:
            # match_intf: a connection only counts if it arrives through the interface of the same name
            client = network.BasicClient(ProbeAgent.NAME_FORMAT % reply.peer_index, reply.addresses, key, verbose=0,
                                         match_intf=num_hosts > 1, probe_timeout=10, attempts=3)
            reachable = set(client.addresses().keys())
        except network.NoValidAddressesFound:
            reachable = set()
        coordinator.report(index, reachable)
        time.sleep(linger_s)  # stay reachable for the host that probes us
    finally:
        agent.shutdown()


def main(argv=None):
    idx, n, addrs, key, nics = (codec.loads_base64(a) for a in (sys.argv if argv is None else argv)[1:6])
    probe(idx, n, addrs, key, nics)


if __name__ == '__main__':
    main()
