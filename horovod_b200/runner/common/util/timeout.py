"""Deadline helper (reference runner/common/util/timeout.py)."""
import time


class TimeoutException(Exception):
    pass


class Timeout(object):
    def __init__(self, timeout, message):
        self._timeout = timeout
        self._timeout_at = time.time() + timeout
        self._message = message

    def remaining(self):
        return max(0, self._timeout_at - time.time())

    def timed_out(self):
        return time.time() > self._timeout_at

    def check_time_out_for(self, activity):
        if self.timed_out():
            raise TimeoutException(self._message.format(activity=activity, timeout=self._timeout))
