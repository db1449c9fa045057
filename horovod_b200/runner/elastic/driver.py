"""Elastic driver: keeps a job running while hosts come and go.

 * a discovery thread polls the host source once per second and tells the running workers (through the notification
   service of the lowest-ranked registered worker first) that the host set changed;
 * `resume()` computes a new rank layout that keeps surviving hosts in front (so rank 0 lives on a host with valid
   state), publishes it to the rendezvous server and spawns processes only on slots that have none;
 * worker exits are recorded in the WorkerStateRegistry whose barrier action decides between stop / blacklist+resume.

Role parity: horovod/runner/elastic/driver.py (ElasticDriver, Results).
"""
import logging
import queue
import threading
import time
from collections import OrderedDict

from horovod_b200.runner.common.util import hosts as hosts_util
from horovod_b200.runner.common.util import timeout as timeout_util
from horovod_b200.runner.elastic import constants
from horovod_b200.runner.elastic.discovery import HostManager
from horovod_b200.runner.elastic.registration import RoundCoordinator
from horovod_b200.runner.elastic.worker import HostUpdateResult, WorkerNotificationClient


class JobResults(object):
    def __init__(self, error_message, worker_results):
        self.error_message = error_message
        self.worker_results = worker_results


class _ResultCollector(object):
    """Collects (exit_code, timestamp) per worker name from the worker threads."""

    def __init__(self):
        self._threads = queue.Queue()
        self._results = {}
        self._lock = threading.Lock()
        self._error = None

    def track(self, thread):
        self._threads.put(thread)

    def put(self, name, value):
        with self._lock:
            self._results[name] = value

    def fail(self, message):
        self._error = message

    def join_all(self):
        while not self._threads.empty():
            self._threads.get().join()
        return JobResults(self._error, dict(self._results))


class ElasticDriver(object):
    def __init__(self, rendezvous, discovery, min_np, max_np, timeout=None, reset_limit=None, cooldown_range=None, verbose=0):
        self._rendezvous = rendezvous
        self._hosts = HostManager(discovery, cooldown_range)
        self._min_np, self._max_np = min_np, max_np
        self._verbose = verbose
        self._timeout = timeout or constants.ELASTIC_TIMEOUT_SECS
        self._lock = threading.RLock()
        self._slots = OrderedDict()          # (host, local_rank) -> SlotInfo of the CURRENT round
        self._by_rank = {}
        self._world_size = 0
        self._slots_ready = threading.Event()
        self._notify_clients = {}            # (host, local_rank) -> WorkerNotificationClient
        self._spawn = None
        self._registry = RoundCoordinator(self, self._hosts, reset_limit=reset_limit)
        self._clean_stop = False
        self._collector = _ResultCollector()
        self._stop = threading.Event()
        self._poller = threading.Thread(target=self._poll_hosts, daemon=True)
        self._poller.start()

    # ---- public API used by the launcher -------------------------------------------------------------------
    def start(self, np, create_worker_fn):
        self._spawn = create_worker_fn
        self._activate(np)

    def resume(self):
        self._activate(self._min_np)

    def stop(self, error_message=None):
        if error_message:
            self._collector.fail(error_message)
        self._stop.set()

    def finished(self):
        return self._stop.is_set()

    def get_results(self):
        return self._collector.join_all()

    def world_size(self):
        return self._world_size

    def local_size(self, host):
        return len([1 for (h, _) in self._slots if h == host])

    def has_rank_assignment(self, host, slot):
        if self._hosts.is_blacklisted(host):
            return False
        return (host, slot) in self._slots

    def get_slot_info(self, host, slot):
        return self._slots.get((host, slot), hosts_util.INVALID_SLOT_INFO) if self.has_rank_assignment(host, slot) \
            else hosts_util.INVALID_SLOT_INFO

    def get_coordinator_info(self):
        return self._by_rank.get(0)

    def record_ready(self, host, slot):
        """A worker is (re-)initialising; blocks until the round it must join is published and returns its id."""
        return self._registry.worker_ready(host, slot)

    def register_worker_server(self, host, slot, addresses, secret_key):
        self._notify_clients[(host, slot)] = WorkerNotificationClient(addresses, secret_key, self._verbose)

    def get_worker_client(self, slot_info):
        return self._notify_clients.get((slot_info.hostname, slot_info.local_rank))

    def wait_for_available_slots(self, min_np, min_hosts=1):
        deadline = timeout_util.Timeout(self._timeout, message='Timed out waiting for {activity}. Make sure the host discovery '
                                        'script reports at least the minimum number of slots (--min-np).')
        self._slots_ready.clear()
        while True:
            cur = self._hosts.current_hosts
            if cur.count_available_slots() >= min_np and len(cur.available_hosts) >= min_hosts:
                return cur
            if self._stop.is_set():
                raise RuntimeError('Job has been shutdown, see above error messages for details.')
            deadline.check_time_out_for('minimum number of slots to become available')
            self._slots_ready.wait(timeout=0.25)
            self._slots_ready.clear()

    # ---- internals -------------------------------------------------------------------------------------------
    def _activate(self, min_np):
        current = self.wait_for_available_slots(min_np)
        new_slots = self._assign(current)
        for si in new_slots:
            self._launch(si)

    def _assign(self, current_hosts):
        """New layout; returns the slots that have no running process yet."""
        with self._lock:
            host_list = [hosts_util.HostInfo(h, current_hosts.get_slots(h)) for h in current_hosts.host_assignment_order]
            layout = hosts_util.get_host_assignments(host_list, self._min_np, self._max_np)
            previous = set(self._slots.keys())
            if previous and not any(h in current_hosts.available_hosts for (h, _) in previous):
                raise RuntimeError('No hosts from previous set remaining, unable to broadcast state.')
            self._slots = OrderedDict(((si.hostname, si.local_rank), si) for si in layout)
            self._by_rank = {si.rank: si for si in layout}
            self._world_size = len(layout)
            self._rendezvous.init(layout)
            fresh = [si for si in layout if (si.hostname, si.local_rank) not in previous]
            self._registry.begin_round(list(self._slots.keys()), [(si.hostname, si.local_rank) for si in fresh])
            return fresh

    def _launch(self, slot_info):
        def run():
            host_event = self._hosts.get_host_event(slot_info.hostname)
            rc, ts = self._spawn(slot_info, [self._stop, host_event])
            self._on_exit(slot_info, rc, ts)
        t = threading.Thread(target=run, daemon=True)
        t.start()
        self._collector.track(t)

    def _on_exit(self, slot_info, exit_code, timestamp):
        name = '{}[{}]'.format(slot_info.hostname, slot_info.local_rank)
        if not self.has_rank_assignment(slot_info.hostname, slot_info.local_rank):
            logging.info('worker %s was removed from the job; ignoring its exit', name)
            return
        if self._stop.is_set() and exit_code != 0 and self._clean_stop:
            return  # torn down by us after another rank finished successfully
        self._registry.worker_exited(slot_info.hostname, slot_info.local_rank, exit_code)
        if exit_code == 0:
            self._clean_stop = True
        if self.finished():
            self._collector.put(name, (exit_code, timestamp))

    def _poll_hosts(self):
        first = True
        while not self._stop.is_set():
            self._slots_ready.set()
            try:
                change = self._hosts.update_available_hosts()
                if change != HostUpdateResult.no_update:
                    self._notify_workers(change)
                    self._slots_ready.set()
            except RuntimeError as e:
                if first:
                    self._collector.fail(str(e))
                    self._stop.set()
                logging.warning('host discovery failed: %s', e)
            first = False
            self._stop.wait(constants.DISCOVER_HOSTS_FREQUENCY_SECS)

    def _notify_workers(self, change):
        """Tell one live worker (the lowest rank that has a registered notification service, normally rank 0) that
        hosts changed; it spreads the decision to the others inside state.check_host_updates()."""
        timestamp = time.time()
        for rank in sorted(self._by_rank):
            si = self._by_rank[rank]
            client = self.get_worker_client(si)
            if client is None:
                continue
            try:
                client.notify_hosts_updated(timestamp, change)
                return
            except Exception:
                if self._verbose >= 2:
                    logging.exception('failed to notify %s[%d] of host updates', si.hostname, si.local_rank)
