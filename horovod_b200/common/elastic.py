"""Elastic training state and the retry loop around the user's training function.

API parity: horovod/common/elastic.py (State, ObjectState, run_fn).  On
HorovodInternalError (a peer died, a collective failed) the last commit is
restored; on HostsUpdatedInterrupt (driver announced new hosts) training just
re-synchronises; in both cases the runtime is shut down and re-initialised
through a fresh rendezvous round.
"""
import functools
import queue

from horovod_b200.common.exceptions import HorovodInternalError, HostsUpdatedInterrupt
from horovod_b200.runner.elastic.worker import HostUpdateResult, WorkerNotificationManager

notification_manager = WorkerNotificationManager()


class State(object):
    """State representation used for tracking in memory state across workers.

    Args:
        bcast_object: Function used to broadcast a variable from rank 0 to the other workers.
        get_rank: Function that returns the current rank of this worker.
    """

    def __init__(self, bcast_object, get_rank):
        self._bcast_object = bcast_object
        self._rank = get_rank
        self._host_messages = queue.Queue()
        self._last_updated_timestamp = 0
        self._reset_callbacks = []

    def register_reset_callbacks(self, callbacks):
        """Callbacks run after a reset (e.g. to rescale the learning rate to the new world size)."""
        self._reset_callbacks.extend(callbacks)

    def on_reset(self):
        self._host_messages = queue.Queue()
        self.reset()
        for callback in self._reset_callbacks:
            callback()

    def on_hosts_updated(self, timestamp, update_res):
        self._host_messages.put((timestamp, update_res))

    def commit(self):
        """Commits all modifications to state tracked by this object to host memory, then checks for host changes."""
        self.save()
        self.check_host_updates()

    def check_host_updates(self):
        """Raises HostsUpdatedInterrupt on every rank at the same point when the driver reported a host change."""
        # Iterate through the update messages sent from the driver; only the latest timestamp matters
        last_updated_timestamp = prev_timestamp = self._last_updated_timestamp
        all_update = HostUpdateResult.no_update
        while not self._host_messages.empty():
            timestamp, update = self._host_messages.get()
            if timestamp > last_updated_timestamp:
                last_updated_timestamp = timestamp
                all_update |= update
        # make the decision rank-consistent: rank 0's view wins
        prev_timestamp, self._last_updated_timestamp, all_update = self._bcast_object(
            (prev_timestamp, last_updated_timestamp, all_update))
        if self._last_updated_timestamp > prev_timestamp:
            raise HostsUpdatedInterrupt(all_update == HostUpdateResult.removed)

    def save(self):
        raise NotImplementedError()

    def restore(self):
        raise NotImplementedError()

    def sync(self):
        raise NotImplementedError()

    def reset(self):
        pass


class ObjectState(State):
    """State for simple Python objects; every kwarg becomes an attribute that is committed / restored / synced."""

    def __init__(self, bcast_object, get_rank, **kwargs):
        self._bcast_object = bcast_object
        self._saved_state = kwargs
        self._set_attrs()
        super(ObjectState, self).__init__(bcast_object=bcast_object, get_rank=get_rank)

    def save(self):
        new_state = {}
        for attr in self._saved_state.keys():
            new_state[attr] = getattr(self, attr)
        self._saved_state = new_state

    def restore(self):
        self._set_attrs()

    def sync(self):
        if self._saved_state:
            self._saved_state = self._bcast_object(self._saved_state)
            self._set_attrs()

    def _set_attrs(self):
        for attr, value in self._saved_state.items():
            setattr(self, attr, value)


def run_fn(func, reset):
    @functools.wraps(func)
    def wrapper(state, *args, **kwargs):
        notification_manager.init()
        notification_manager.register_listener(state)
        skip_sync = False
        try:
            while True:
                try:
                    if not skip_sync:
                        state.sync()
                    return func(state, *args, **kwargs)
                except HorovodInternalError:
                    state.restore()
                    skip_sync = False
                except HostsUpdatedInterrupt as e:
                    skip_sync = e.skip_sync
                reset()
                state.on_reset()
        finally:
            notification_manager.remove_listener(state)
    return wrapper
