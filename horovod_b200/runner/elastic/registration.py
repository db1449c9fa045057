"""Round coordination for elastic jobs.

A *round* is one rank assignment.  While a round runs, workers report to the coordinator in three ways:
  ready    a worker re-initialises (after a failed collective or a host-change interrupt) and asks for its new rank
  success  a worker process exited with code 0
  failure  a worker process died / exited non-zero
Once every member of the round has reported, the coordinator decides: finish the job (someone succeeded or everybody
failed), or blacklist the failed hosts and ask the driver for the next round.  Callers of `worker_ready()` block until
that next round is published, so they come back with their new rank.  Workers that were spawned FOR the current round
get their rank immediately (the transport's own mesh rendezvous is what synchronises them with the survivors).

Capability parity: horovod/runner/elastic/registration.py (WorkerStateRegistry: READY/SUCCESS/FAILURE barrier and the
stop / blacklist / reset-limit / resume decision) — re-designed around a condition variable and explicit rounds
instead of a resettable threading.Barrier.
"""
import logging
import threading

READY, SUCCESS, FAILURE = 'ready', 'success', 'failure'


class RoundCoordinator(object):
    def __init__(self, driver, host_manager, reset_limit=None):
        self._driver = driver
        self._hosts = host_manager
        self._reset_limit = reset_limit
        self._resets = 0
        self._cv = threading.Condition()
        self._round = 0
        self._members = set()     # slots that must report before the round can be decided
        self._fresh = set()       # members spawned for this round that have not asked for their rank yet
        self._reports = {}        # slot -> READY | SUCCESS | FAILURE
        self._deciding = False

    # ---- driver side -----------------------------------------------------------------------------------------
    def begin_round(self, all_slots, new_slots):
        """Publishes a new assignment: `all_slots` are its members, `new_slots` the ones that get a fresh process."""
        with self._cv:
            self._round += 1
            self._members = set(all_slots)
            self._fresh = set(new_slots)
            self._reports = {}
            self._cv.notify_all()
            return self._round

    @property
    def round(self):
        return self._round

    def reported(self, kind):
        with self._cv:
            return {s for s, k in self._reports.items() if k == kind}

    # ---- worker side -------------------------------------------------------------------------------------------
    def worker_ready(self, host, slot, timeout=None):
        """Called on behalf of a worker that is (re-)initialising. Returns the round whose assignment it must use."""
        key = (host, slot)
        with self._cv:
            if self._driver.finished():
                return self._round
            if key in self._fresh:
                self._fresh.discard(key)
                return self._round
            if key not in self._members or self._hosts.is_blacklisted(host):
                return self._round  # not part of the job any more: the caller will be told rank -1
            seen = self._round
            self._reports[key] = READY
            self._maybe_decide()
            while self._round == seen and not self._driver.finished():
                if not self._cv.wait(timeout=timeout if timeout else 1.0) and timeout:
                    break
            return self._round

    def worker_exited(self, host, slot, exit_code):
        key = (host, slot)
        with self._cv:
            rnd = self._round
            if self._driver.finished() or key not in self._members:
                return rnd
            self._fresh.discard(key)
            self._reports[key] = SUCCESS if exit_code == 0 else FAILURE
            self._maybe_decide()
            return rnd

    # ---- decision ------------------------------------------------------------------------------------------------
    def _maybe_decide(self):
        # called with the lock held
        if self._deciding or not self._members or set(self._reports) != self._members:
            return
        self._deciding = True
        try:
            self._decide()
        finally:
            self._deciding = False

    def _decide(self):
        kinds = list(self._reports.values())
        logging.info('round %d complete: %d ready, %d succeeded, %d failed', self._round, kinds.count(READY),
                     kinds.count(SUCCESS), kinds.count(FAILURE))
        if SUCCESS in kinds:
            # training finished on some rank: the job is done, stragglers are torn down
            self._driver.stop()
            self._cv.notify_all()
            return
        if kinds.count(FAILURE) == len(kinds):
            logging.error('every worker of round %d failed -> stop', self._round)
            self._driver.stop()
            self._cv.notify_all()
            return
        for (host, _), kind in self._reports.items():
            if kind == FAILURE:
                self._hosts.blacklist(host)
        if all(self._hosts.is_blacklisted(h) for (h, _) in self._members):
            logging.error('all hosts of round %d are blacklisted -> stop', self._round)
            self._driver.stop()
            self._cv.notify_all()
            return
        if self._reset_limit is not None and self._resets >= self._reset_limit:
            self._driver.stop(error_message='Job has been reset {} times which exceeds the --reset-limit of {}'.format(
                self._resets, self._reset_limit))
            self._cv.notify_all()
            return
        self._resets += 1
        try:
            self._driver.resume()   # -> begin_round() -> wakes the waiting worker_ready() callers
        except Exception as e:
            logging.exception('could not start the next round')
            self._driver.stop(error_message=str(e))
            self._cv.notify_all()
