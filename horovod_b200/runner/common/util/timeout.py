"""Deadline object handed around by the launcher ("all of this has to happen within N seconds").
API parity: horovod/runner/common/util/timeout.py (`Timeout(timeout, message).remaining() / timed_out() /
check_time_out_for(activity)`, `TimeoutException`).  Monotonic clock: an NTP step cannot fire or stall a deadline."""
import time


class TimeoutException(Exception):
    pass


class Timeout:
    def __init__(self, timeout, message):
        self._seconds = timeout
        self._deadline = time.monotonic() + timeout
        self._message = message

    def remaining(self):
        left = self._deadline - time.monotonic()
        return left if left > 0 else 0

    def timed_out(self):
        return time.monotonic() > self._deadline

    def check_time_out_for(self, activity):
        """Raises TimeoutException (message formatted with {activity} and {timeout}) once the deadline passed."""
        if self.timed_out():
            raise TimeoutException(self._message.format(activity=activity, timeout=self._seconds))
