"""Data-loader scaffolding: a minimal loader protocol and a background-thread prefetch mixin.

Capability parity: horovod/data/data_loader_base.py:20-171 (``BaseDataLoader`` with the ``_iterate`` /
``_process_batch`` hooks; ``AsyncDataLoaderMixin`` with ``async_loader_queue_size``, ``close_async_loader`` and
exception forwarding from the producer thread).  The implementation is new: one producer thread per *epoch request*
coordinated through typed sentinels instead of ``None`` markers, so ``None`` is a legal batch, a producer error is
re-raised exactly once at the position it happened, and closing never relies on draining heuristics.
"""
import queue
import threading


class BaseDataLoader:
    """Subclasses implement ``__len__`` and ``_iterate`` (a generator of raw batches)."""

    def __len__(self):
        raise NotImplementedError

    def _iterate(self):
        raise NotImplementedError

    def _process_batch(self, batch):
        """Trainer hook applied to every batch on the consumer side (identity by default)."""
        return batch

    def __iter__(self):
        for batch in self._iterate():
            yield self._process_batch(batch)


class _EndOfEpoch:
    __slots__ = ()


class _ProducerError:
    __slots__ = ('exc',)

    def __init__(self, exc):
        self.exc = exc


class AsyncDataLoaderMixin:
    """Mix in *before* a ``BaseDataLoader`` implementation::

        class AsyncLoader(AsyncDataLoaderMixin, MyLoader): ...

    A daemon thread walks ``self._iterate()`` epoch after epoch and parks up to ``async_loader_queue_size`` batches in
    a bounded queue; ``__iter__`` pops one epoch's worth.  ``async_loader_queue_size=0`` disables the thread (the
    loader then behaves exactly like the synchronous base class).
    """

    def __init__(self, *args, async_loader_queue_size=64, debug_data_loader=False, **kwargs):
        self.async_loader_queue_size = int(async_loader_queue_size)
        self.debug_data_loader = debug_data_loader
        super().__init__(*args, **kwargs)
        self._q = None
        self._producer = None
        self._stop = threading.Event()
        self._epochs_wanted = threading.Semaphore(0)

    # -- producer ---------------------------------------------------------------------------------------------------
    def _put(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _produce(self):
        while not self._stop.is_set():
            if not self._epochs_wanted.acquire(timeout=0.1):
                continue
            try:
                for batch in self._iterate():
                    if not self._put(batch):
                        return
            except BaseException as exc:  # forwarded to, and re-raised in, the consumer
                self._put(_ProducerError(exc))
            if not self._put(_EndOfEpoch()):
                return

    def _ensure_started(self):
        if self._producer is None or not self._producer.is_alive():
            self._stop.clear()
            self._q = queue.Queue(self.async_loader_queue_size)
            self._epochs_wanted = threading.Semaphore(0)
            self._producer = threading.Thread(target=self._produce, name='hvd-data-prefetch', daemon=True)
            self._producer.start()

    # -- consumer ---------------------------------------------------------------------------------------------------
    def __iter__(self):
        if self.async_loader_queue_size <= 0:
            yield from super().__iter__()
            return
        self._ensure_started()
        self._epochs_wanted.release()
        while True:
            item = self._q.get()
            if isinstance(item, _EndOfEpoch):
                return
            if isinstance(item, _ProducerError):
                # the producer still emits the end-of-epoch marker after an error; swallow it so the next epoch is clean
                nxt = self._q.get()
                assert isinstance(nxt, _EndOfEpoch)
                raise item.exc
            yield self._process_batch(item)

    def close_async_loader(self):
        """Stops the producer thread; safe to call more than once and from ``__del__``."""
        if self._producer is None:
            return
        self._stop.set()
        self._producer.join(timeout=10)
        self._producer = None
        self._q = None

    def __del__(self):
        try:
            self.close_async_loader()
        except Exception:
            pass
