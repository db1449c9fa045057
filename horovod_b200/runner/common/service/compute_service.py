"""Compute service: the meeting point of a training job and a side job of data-processing ("compute") workers.

Role parity: horovod/runner/common/service/compute_service.py (`ComputeService` / `ComputeClient`: register_dispatcher,
wait_for_dispatcher_registration, register_worker_for_dispatcher, wait_for_dispatcher_worker_registration, shutdown,
wait_for_shutdown).  It is framework agnostic; the TensorFlow data-service glue is horovod_b200/tensorflow/data.

Design: the service is a tiny *notice board* — `Post(key, value)` pins a value, `Await(keys, timeout)` blocks until every
key is pinned.  The dispatcher / worker / shutdown vocabulary of the reference is expressed as keys on that board
(`('dispatcher', i)`, `('worker', i, w)`, `('shutdown',)`) by `ComputeClient`, so new synchronisation points need no new
message types.  Posting is idempotent for an equal value (RPCs are retried) and an error for a different one.
"""
import threading
import time

from horovod_b200.runner.common.util import network
from horovod_b200.runner.common.util.timeout import TimeoutException


class Post:
    def __init__(self, key, value=True):
        self.key, self.value = key, value


class Await:
    def __init__(self, keys, timeout=None):
        self.keys, self.timeout = list(keys), timeout


class AwaitCount:
    """Blocks until at least `count` keys starting with `prefix` are pinned (e.g. any N workers of a dispatcher, whose ids
    are not known in advance)."""

    def __init__(self, prefix, count, timeout=None):
        self.prefix, self.count, self.timeout = tuple(prefix), count, timeout


class Board:
    def __init__(self, values):
        self.values = values


class ComputeService(network.BasicService):
    NAME = 'Compute service'

    def __init__(self, dispatchers, workers_per_dispatcher, key, nics=None):
        if dispatchers <= 0:
            raise ValueError('The number of dispatchers must be larger than 0: %s' % dispatchers)
        if workers_per_dispatcher <= 0:
            raise ValueError('The number of workers per dispatcher must be larger than 0: %s' % workers_per_dispatcher)
        self.dispatchers, self.workers_per_dispatcher = dispatchers, workers_per_dispatcher
        self._pinned = {('config', 'workers_per_dispatcher'): workers_per_dispatcher, ('config', 'dispatchers'): dispatchers}
        self._changed = threading.Condition()
        super().__init__(ComputeService.NAME, key, nics)

    # -- board rules --------------------------------------------------------------------------------------------------------------
    def _matching(self, prefix):
        return [k for k in self._pinned if k[:len(prefix)] == prefix]

    def _validate(self, key):
        """Called with the lock held."""
        if key[0] in ('dispatcher', 'worker') and not 0 <= key[1] < self.dispatchers:
            raise IndexError('Dispatcher id must be within [0..%d]: %s' % (self.dispatchers - 1, key[1]))
        if key[0] == 'worker' and len(key) == 3 and key not in self._pinned:
            if len(self._matching(key[:2])) >= self.workers_per_dispatcher:
                raise IndexError('Dispatcher %d already has its %d workers; cannot register worker %s'
                                 % (key[1], self.workers_per_dispatcher, key[2]))

    def _wait_until(self, ready, timeout, what):
        """Lock held.  Returns when `ready()`; TimeoutException after `timeout` seconds."""
        deadline = None if timeout is None else time.monotonic() + timeout
        while not ready():
            left = 1.0 if deadline is None else deadline - time.monotonic()
            if left <= 0:
                raise TimeoutException('Timed out after %s s waiting for %s' % (timeout, what()))
            self._changed.wait(left)

    # -- RPC ------------------------------------------------------------------------------------------------------------------------
    def _handle(self, req, client_address):
        if isinstance(req, Post):
            key = tuple(req.key)
            with self._changed:
                self._validate(key)
                if key in self._pinned and self._pinned[key] != req.value:
                    raise ValueError('%s is already registered as %r, cannot change it to %r' % (key, self._pinned[key], req.value))
                self._pinned[key] = req.value
                self._changed.notify_all()
            return network.AckResponse()
        if isinstance(req, Await):
            keys = [tuple(k) for k in req.keys]
            with self._changed:
                for k in keys:
                    self._validate(k)
                self._wait_until(lambda: all(k in self._pinned for k in keys), req.timeout,
                                 lambda: [k for k in keys if k not in self._pinned])
                return Board({k: self._pinned[k] for k in keys})
        if isinstance(req, AwaitCount):
            with self._changed:
                self._validate(req.prefix)
                self._wait_until(lambda: len(self._matching(req.prefix)) >= req.count, req.timeout,
                                 lambda: '%d x %s (have %d)' % (req.count, req.prefix, len(self._matching(req.prefix))))
                return Board({k: self._pinned[k] for k in self._matching(req.prefix)})
        return super()._handle(req, client_address)


class ComputeClient(network.BasicClient):
    def __init__(self, compute_addresses, key, verbose=1):
        super().__init__(ComputeService.NAME, compute_addresses, key, verbose)
        self._wpd = None

    def register_dispatcher(self, dispatcher_id, dispatcher_address):
        self._send(Post(('dispatcher', dispatcher_id), dispatcher_address))

    def wait_for_dispatcher_registration(self, dispatcher_id, timeout):
        """-> the dispatcher's address; TimeoutException after `timeout` seconds."""
        key = ('dispatcher', dispatcher_id)
        return self._send(Await([key], timeout)).values[key]

    def register_worker_for_dispatcher(self, dispatcher_id, worker_id):
        self._send(Post(('worker', dispatcher_id, worker_id)))

    def wait_for_dispatcher_worker_registration(self, dispatcher_id, timeout):
        self._send(AwaitCount(('worker', dispatcher_id), self._workers_per_dispatcher(), timeout))

    def _workers_per_dispatcher(self):
        if self._wpd is None:
            self._wpd = self._send(Await([('config', 'workers_per_dispatcher')], 10)).values[('config', 'workers_per_dispatcher')]
        return self._wpd

    def shutdown(self):
        self._send(Post(('shutdown',)))

    def wait_for_shutdown(self):
        # long poll in slices so that a vanished service surfaces as a connection error instead of a silent hang
        while True:
            try:
                self._send(Await([('shutdown',)], 30))
                return
            except TimeoutException:
                continue
