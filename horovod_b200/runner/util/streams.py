"""An in-memory pipe between a producer thread and a consumer thread.

Role parity: horovod/runner/util/streams.py (`Pipe`, used to forward a worker's stdout/stderr through an RPC stream).  The
reference hands over one buffer at a time; this pipe queues up to `max_chunks` chunks so that a bursty writer is not
serialised behind a slow reader, keeps str / bytes as written, and reports end-of-stream as `None` once drained.
"""
import collections
import threading


class Pipe:
    def __init__(self, max_chunks=64):
        self._chunks = collections.deque()
        self._max = max(1, max_chunks)
        self._cond = threading.Condition()
        self._closed = False

    def write(self, buf):
        if not buf:
            return
        with self._cond:
            while len(self._chunks) >= self._max and not self._closed:
                self._cond.wait()
            if self._closed:
                raise RuntimeError('Pipe is closed')
            self._chunks.append(buf)
            self._cond.notify_all()

    def read(self, length=-1):
        """Blocks for data; returns at most `length` items of ONE written chunk (all of it for length <= 0), or None when
        the pipe is closed and empty."""
        with self._cond:
            while not self._chunks and not self._closed:
                self._cond.wait()
            if not self._chunks:
                return None
            head = self._chunks[0]
            if 0 < length < len(head):
                self._chunks[0] = head[length:]
                head = head[:length]
            else:
                self._chunks.popleft()
            self._cond.notify_all()
            return head

    def flush(self):
        pass

    def close(self):
        with self._cond:
            self._closed = True
            self._cond.notify_all()

    @property
    def closed(self):
        return self._closed

    def __iter__(self):
        while True:
            chunk = self.read()
            if chunk is None:
                return
            yield chunk
