"""Driver-side RPC service of the pre-flight network probe: every task registers its addresses, the driver tells each
task which peer to ping, tasks report which interfaces reached it (reference runner/common/service/driver_service.py)."""
import threading

from horovod_b200.runner.common.util import network


class RegisterTaskRequest(object):
    def __init__(self, index, task_addresses, host_hash):
        self.index = index
        self.task_addresses = task_addresses
        self.host_hash = host_hash


class RegisterTaskToTaskAddressesRequest(object):
    def __init__(self, index, task_addresses):
        self.index = index
        self.task_addresses = task_addresses


class AllTaskAddressesRequest(object):
    def __init__(self, index):
        self.index = index


class AllTaskAddressesResponse(object):
    def __init__(self, all_task_addresses):
        self.all_task_addresses = all_task_addresses


class BasicDriverService(network.BasicService):
    def __init__(self, num_proc, name, key, nic):
        super(BasicDriverService, self).__init__(name, key, nic)
        self._num_proc = num_proc
        self._all_task_addresses = {}
        self._task_addresses_for_driver = {}
        self._task_addresses_for_tasks = {}
        self._task_index_host_hash = {}
        self._task_host_hash_indices = {}
        self._wait_cond = threading.Condition()

    def _handle(self, req, client_address):
        if isinstance(req, RegisterTaskRequest):
            with self._wait_cond:
                assert 0 <= req.index < self._num_proc
                self._all_task_addresses[req.index] = req.task_addresses
                # Just use source address for service for fast probing.
                self._task_addresses_for_driver[req.index] = self._filter_by_ip(req.task_addresses, client_address[0])
                # Remove host hash earlier registered under this index.
                if req.index in self._task_index_host_hash:
                    earlier_host_hash = self._task_index_host_hash[req.index]
                    if earlier_host_hash != req.host_hash:
                        self._task_host_hash_indices[earlier_host_hash].remove(req.index)
                # Make index -> host hash map.
                self._task_index_host_hash[req.index] = req.host_hash
                # Make host hash -> indices map.
                self._task_host_hash_indices.setdefault(req.host_hash, [])
                if req.index not in self._task_host_hash_indices[req.host_hash]:
                    self._task_host_hash_indices[req.host_hash].append(req.index)
                    self._task_host_hash_indices[req.host_hash].sort()
                self._wait_cond.notify_all()
            return network.AckResponse()
        if isinstance(req, RegisterTaskToTaskAddressesRequest):
            self.register_task_to_task_addresses(req.index, req.task_addresses)
            return network.AckResponse()
        if isinstance(req, AllTaskAddressesRequest):
            return AllTaskAddressesResponse(self._all_task_addresses[req.index])
        return super(BasicDriverService, self)._handle(req, client_address)

    def _filter_by_ip(self, addresses, target_ip):
        for intf, intf_addresses in addresses.items():
            for ip, port in intf_addresses:
                if ip == target_ip:
                    return {intf: [(ip, port)]}
        return addresses

    def all_task_addresses(self, index):
        with self._wait_cond:
            return self._all_task_addresses[index].copy()

    def task_addresses_for_driver(self, index):
        with self._wait_cond:
            return self._task_addresses_for_driver[index].copy()

    def task_addresses_for_tasks(self, index):
        with self._wait_cond:
            return self._task_addresses_for_tasks[index].copy()

    def register_task_to_task_addresses(self, index, task_addresses):
        with self._wait_cond:
            assert 0 <= index < self._num_proc
            self._task_addresses_for_tasks[index] = task_addresses
            self._wait_cond.notify_all()

    def task_indices(self):
        with self._wait_cond:
            return list(self._task_index_host_hash.keys())

    def task_host_hash_indices(self):
        with self._wait_cond:
            return self._task_host_hash_indices.copy()

    def task_index_host_hash(self, index):
        with self._wait_cond:
            return self._task_index_host_hash[index]

    def wait_for_initial_registration(self, timeout):
        with self._wait_cond:
            while len(self._all_task_addresses) < self._num_proc:
                self._wait_cond.wait(timeout.remaining())
                timeout.check_time_out_for('tasks to start')

    def wait_for_task_to_task_address_updates(self, timeout):
        with self._wait_cond:
            while len(self._task_addresses_for_tasks) < self._num_proc:
                self._wait_cond.wait(timeout.remaining())
                timeout.check_time_out_for('tasks to update task-to-task addresses')


class BasicDriverClient(network.BasicClient):
    def __init__(self, name, driver_addresses, key, verbose, match_intf=False):
        super(BasicDriverClient, self).__init__(name, driver_addresses, key, verbose, match_intf=match_intf)

    def register_task(self, index, task_addresses, host_hash):
        self._send(RegisterTaskRequest(index, task_addresses, host_hash))

    def all_task_addresses(self, index):
        resp = self._send(AllTaskAddressesRequest(index))
        return resp.all_task_addresses

    def register_task_to_task_addresses(self, index, task_addresses):
        self._send(RegisterTaskToTaskAddressesRequest(index, task_addresses))
