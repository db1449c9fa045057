"""Elastic training: committed in-memory state + the retry loop around the user's training function.

Public API parity with horovod/common/elastic.py (`State`, `ObjectState`, `run_fn`, `notification_manager`); the
implementation is organised differently:

* host-change notifications from the driver land in a small lock-protected log (`_HostEventLog`); `check_host_updates`
  folds everything newer than the last agreed timestamp into one decision and takes RANK 0's view of
  (previously agreed timestamp, newest timestamp, merged result) for every rank — a worker that joined later has a
  different local history and must not decide on its own;
* `ObjectState` snapshots its tracked attributes with `copy.deepcopy`, so committing a list/dict attribute and mutating
  it afterwards cannot corrupt the snapshot;
* the retry loop is an explicit little state machine (`_ElasticLoop`) instead of nested try blocks.
"""
import copy
import functools
import threading

from horovod_b200.common.exceptions import HorovodInternalError, HostsUpdatedInterrupt
from horovod_b200.runner.elastic.worker import HostUpdateResult, WorkerNotificationManager

notification_manager = WorkerNotificationManager()


class _HostEventLog:
    """(timestamp, HostUpdateResult) records appended by the notification thread, folded by the training thread."""

    def __init__(self):
        self._lock = threading.Lock()
        self._records = []

    def append(self, timestamp, result):
        with self._lock:
            self._records.append((timestamp, result))

    def clear(self):
        with self._lock:
            self._records = []

    def fold_newer_than(self, timestamp):
        """Removes every record; returns (newest timestamp seen or `timestamp`, OR of the results newer than it)."""
        with self._lock:
            records, self._records = self._records, []
        newest, merged = timestamp, HostUpdateResult.no_update
        for ts, res in records:
            if ts > timestamp:
                merged |= res
                newest = max(newest, ts)
        return newest, merged


class State:
    """What survives a reset.  Subclasses implement `save` / `restore` / `sync` (and optionally `reset`).

    bcast_object(obj) -> rank 0's obj on every rank; get_rank() -> this worker's current rank.
    """

    def __init__(self, bcast_object, get_rank):
        self._bcast_object = bcast_object
        self._rank = get_rank
        self._host_events = _HostEventLog()
        self._agreed_timestamp = 0
        self._after_reset = []

    # -- hooks the retry loop and the notification service call ---------------------------------------------------------
    def register_reset_callbacks(self, callbacks):
        """Functions to run after every reset, e.g. to rescale the learning rate to the new world size."""
        self._after_reset += list(callbacks)

    def on_reset(self):
        self._host_events.clear()
        self.reset()
        for fn in self._after_reset:
            fn()

    def on_hosts_updated(self, timestamp, update_res):
        self._host_events.append(timestamp, update_res)

    # -- user API -------------------------------------------------------------------------------------------------------------
    def commit(self):
        """Snapshot the tracked state, then give a pending host change the chance to interrupt training."""
        self.save()
        self.check_host_updates()

    def check_host_updates(self):
        """Raises HostsUpdatedInterrupt on ALL ranks in the same call when the driver announced a membership change."""
        newest, merged = self._host_events.fold_newer_than(self._agreed_timestamp)
        before, newest, merged = self._bcast_object((self._agreed_timestamp, newest, merged))
        self._agreed_timestamp = newest
        if newest > before:
            # nothing has to be re-synchronised when hosts were only removed: the survivors already agree
            raise HostsUpdatedInterrupt(skip_sync=(merged == HostUpdateResult.removed))

    def save(self):
        raise NotImplementedError

    def restore(self):
        raise NotImplementedError

    def sync(self):
        raise NotImplementedError

    def reset(self):
        pass


class ObjectState(State):
    """Tracks plain Python values: `ObjectState(bcast, rank, epoch=0, batch=0)` exposes `.epoch` / `.batch`."""

    def __init__(self, bcast_object, get_rank, **kwargs):
        super().__init__(bcast_object=bcast_object, get_rank=get_rank)
        self._tracked = list(kwargs)
        self._snapshot = {}
        self._install(kwargs)

    def _install(self, values):
        self._snapshot = {k: copy.deepcopy(values[k]) for k in self._tracked}
        for k in self._tracked:
            setattr(self, k, values[k])

    def save(self):
        self._snapshot = {k: copy.deepcopy(getattr(self, k)) for k in self._tracked}

    def restore(self):
        for k in self._tracked:
            setattr(self, k, copy.deepcopy(self._snapshot[k]))

    def sync(self):
        if self._tracked:
            self._install(self._bcast_object(self._snapshot))

    # kept for subclasses / callers that used the reference's attribute name
    @property
    def _saved_state(self):
        return self._snapshot


class _ElasticLoop:
    """run `func(state, ...)` until it returns; HorovodInternalError -> restore the last commit and re-sync,
    HostsUpdatedInterrupt -> keep the current state (re-sync unless only removals happened); both -> `reset()` the runtime
    (shutdown + re-init through a new rendezvous round) and run the reset callbacks."""

    def __init__(self, func, reset):
        self.func, self.reset = func, reset

    def __call__(self, state, *args, **kwargs):
        notification_manager.init()
        notification_manager.register_listener(state)
        need_sync = True
        try:
            while True:
                outcome, value = self._attempt(state, need_sync, args, kwargs)
                if outcome == 'done':
                    return value
                need_sync = value
                self.reset()
                state.on_reset()
        finally:
            notification_manager.remove_listener(state)

    def _attempt(self, state, need_sync, args, kwargs):
        try:
            if need_sync:
                state.sync()
            return 'done', self.func(state, *args, **kwargs)
        except HorovodInternalError:
            state.restore()
            return 'retry', True
        except HostsUpdatedInterrupt as e:
            return 'retry', not e.skip_sync


def run_fn(func, reset):
    """Wraps `func(state, ...)` in the elastic retry loop; `reset` re-initialises the runtime after a failure."""
    loop = _ElasticLoop(func, reset)

    @functools.wraps(func)
    def wrapper(state, *args, **kwargs):
        return loop(state, *args, **kwargs)
    return wrapper
