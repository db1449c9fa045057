"""Worker-side notification service: the elastic driver tells (rank-0's) workers that the host set changed.

Role parity: horovod/runner/elastic/worker.py (WorkerNotificationManager / Service / Client, HostsUpdatedRequest).
Transport: the HMAC-authenticated pickle RPC of runner/common/util/network.py.
"""
import os
import threading

from horovod_b200.runner.common.util import network, secret


class HostUpdateResult:
    no_update = 0
    removed = 1
    added = 2
    mixed = 3


class HostsUpdatedRequest(object):
    """Notifies worker that the set of available hosts/slots has changed."""

    def __init__(self, timestamp, res=HostUpdateResult.mixed):
        self.timestamp = timestamp
        self.res = res


class WorkerNotificationService(network.BasicService):
    NAME = 'worker notification service'

    def __init__(self, key, nic, manager):
        super(WorkerNotificationService, self).__init__(WorkerNotificationService.NAME, key, nic)
        self._manager = manager

    def _handle(self, req, client_address):
        if isinstance(req, HostsUpdatedRequest):
            self._manager.handle_hosts_updated(req.timestamp, req.res)
            return network.AckResponse()
        return super(WorkerNotificationService, self)._handle(req, client_address)


class WorkerNotificationClient(network.BasicClient):
    def __init__(self, addresses, key, verbose, match_intf=False):
        super(WorkerNotificationClient, self).__init__(WorkerNotificationService.NAME, addresses, key, verbose,
                                                       match_intf=match_intf)

    def notify_hosts_updated(self, timestamp, update_res):
        self._send(HostsUpdatedRequest(timestamp, update_res))


class WorkerNotificationManager(object):
    def __init__(self):
        self._lock = threading.Lock()
        self._service = None
        self._listeners = set()

    def init(self, rendezvous_addr=None, rendezvous_port=None, nic=None, hostname=None, local_rank=None):
        with self._lock:
            if self._service:
                return
            rendezvous_addr = rendezvous_addr or os.environ.get('HOROVOD_GLOO_RENDEZVOUS_ADDR')
            if not rendezvous_addr:
                return  # not launched by the elastic driver: nothing to listen to
            rendezvous_port = rendezvous_port if rendezvous_port is not None else int(os.environ.get('HOROVOD_GLOO_RENDEZVOUS_PORT'))
            nic = nic or os.environ.get('HOROVOD_GLOO_IFACE')
            hostname = hostname or os.environ.get('HOROVOD_HOSTNAME')
            local_rank = local_rank if local_rank is not None else os.environ.get('HOROVOD_LOCAL_RANK')
            secret_key = secret.make_secret_key() if not os.environ.get(secret.HOROVOD_SECRET_KEY) else \
                secret.decode_key(os.environ[secret.HOROVOD_SECRET_KEY])
            self._service = WorkerNotificationService(secret_key, nic, self)
            value = network.dumps_base64((self._service.addresses(), secret_key))
            from horovod_b200.runner.http.http_client import put_data_into_kvstore
            put_data_into_kvstore(rendezvous_addr, rendezvous_port, 'worker_addresses', f'{hostname}:{local_rank}', value)

    def register_listener(self, listener):
        self._listeners.add(listener)

    def remove_listener(self, listener):
        self._listeners.discard(listener)

    def handle_hosts_updated(self, timestamp, update_res):
        for listener in self._listeners:
            listener.on_hosts_updated(timestamp, update_res)
