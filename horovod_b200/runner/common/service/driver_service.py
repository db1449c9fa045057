"""The registry the tasks of a job report to (reference horovod/runner/common/service/driver_service.py: `BasicDriverService`
:43-178, `BasicDriverClient` :196-215): which task index lives at which addresses on which host, plus the task-to-task
addresses found by the tasks probing each other.  The launcher's own NIC probe uses a leaner protocol
(`runner/driver/driver_service.py`); this is the general form the Spark-style integrations of the reference build on.
"""
import threading

from horovod_b200.runner.common.util import network


class RegisterTaskRequest(object):
    def __init__(self, index, task_addresses, host_hash):
        self.index, self.task_addresses, self.host_hash = index, task_addresses, host_hash


class RegisterTaskToTaskAddressesRequest(object):
    def __init__(self, index, task_addresses):
        self.index, self.task_addresses = index, task_addresses


class AllTaskAddressesRequest(object):
    def __init__(self, index):
        self.index = index


class AllTaskAddressesResponse(object):
    def __init__(self, all_task_addresses):
        self.all_task_addresses = all_task_addresses


class BasicDriverService(network.BasicService):
    def __init__(self, num_proc, name, key, nic=None):
        super(BasicDriverService, self).__init__(name, key, nic)
        self._num_proc = num_proc
        self._cond = threading.Condition()
        self._all_task_addresses = {}
        self._task_addresses_for_driver = {}
        self._task_addresses_for_tasks = {}
        self._task_index_host_hash = {}
        self._task_host_hash_indices = {}

    def _handle(self, req, client_address):
        if isinstance(req, RegisterTaskRequest):
            with self._cond:
                self._all_task_addresses[req.index] = req.task_addresses
                # the addresses through which the DRIVER can reach the task: those on the interface this request arrived by
                self._task_addresses_for_driver[req.index] = self._filter_by_ip(req.task_addresses, client_address[0])
                previous = self._task_index_host_hash.get(req.index)
                if previous is not None and previous != req.host_hash:      # a restarted task may land on another host
                    self._task_host_hash_indices[previous].remove(req.index)
                self._task_index_host_hash[req.index] = req.host_hash
                members = self._task_host_hash_indices.setdefault(req.host_hash, [])
                if req.index not in members:
                    members.append(req.index)
                    members.sort()
                self._cond.notify_all()
            return network.AckResponse()
        if isinstance(req, RegisterTaskToTaskAddressesRequest):
            self.register_task_to_task_addresses(req.index, req.task_addresses)
            return network.AckResponse()
        if isinstance(req, AllTaskAddressesRequest):
            return AllTaskAddressesResponse(self.all_task_addresses(req.index))
        return super(BasicDriverService, self)._handle(req, client_address)

    @staticmethod
    def _filter_by_ip(addresses, target_ip):
        kept = {intf: [a for a in addrs if a[0] == target_ip] for intf, addrs in addresses.items()}
        kept = {intf: addrs for intf, addrs in kept.items() if addrs}
        return kept or addresses             # NAT / loopback: nothing matches, keep everything

    def all_task_addresses(self, index):
        with self._cond:
            return dict(self._all_task_addresses[index])

    def task_addresses_for_driver(self, index):
        with self._cond:
            return dict(self._task_addresses_for_driver[index])

    def task_addresses_for_tasks(self, index):
        with self._cond:
            return dict(self._task_addresses_for_tasks[index])

    def register_task_to_task_addresses(self, index, task_addresses):
        with self._cond:
            self._task_addresses_for_tasks[index] = task_addresses
            self._cond.notify_all()

    def task_indices(self):
        with self._cond:
            return sorted(self._all_task_addresses)

    def task_host_hash_indices(self):
        """host hash -> sorted task indices on that host."""
        with self._cond:
            return {h: list(v) for h, v in self._task_host_hash_indices.items() if v}

    def task_index_host_hash(self, index):
        with self._cond:
            return self._task_index_host_hash[index]

    def _wait(self, done, timeout, activity):
        with self._cond:
            while not done():
                self._cond.wait(0.1)
                if hasattr(timeout, 'check_time_out_for'):
                    timeout.check_time_out_for(activity)

    def wait_for_initial_registration(self, timeout):
        self._wait(lambda: len(self._all_task_addresses) >= self._num_proc, timeout, 'tasks to start')

    def wait_for_task_to_task_address_updates(self, timeout):
        self._wait(lambda: len(self._task_addresses_for_tasks) >= self._num_proc, timeout, 'tasks to update task-to-task addresses')


class BasicDriverClient(network.BasicClient):
    def __init__(self, name, driver_addresses, key, verbose=0, match_intf=False):
        super(BasicDriverClient, self).__init__(name, driver_addresses, key, verbose, match_intf=match_intf)

    def register_task(self, index, task_addresses, host_hash):
        self._send(RegisterTaskRequest(index, task_addresses, host_hash))

    def all_task_addresses(self, index):
        return self._send(AllTaskAddressesRequest(index)).all_task_addresses

    def register_task_to_task_addresses(self, index, task_addresses):
        self._send(RegisterTaskToTaskAddressesRequest(index, task_addresses))
