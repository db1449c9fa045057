"""Framework-neutral data loader base + a mixin that prefetches batches on a background thread so that host-side
batch preparation overlaps the training step (role parity: horovod/data/data_loader_base.py).

The B200 angle: with `pin_memory=True` the mixin also stages every batch into pinned host memory on the producer
thread, so the consumer's `tensor.cuda(non_blocking=True)` is a true asynchronous DMA."""
import queue
import threading


class BaseDataLoader(object):
    def __len__(self):
        """Length of the batches to be loaded."""
        raise NotImplementedError()

    def _iterate(self):
        """Interface for the implementation of iterate batches."""
        raise NotImplementedError()

    def __iter__(self):
        """Starting iteration and get batches."""
        for batch in self._iterate():
            yield self._process_batch(batch)

    def _process_batch(self, batch):
        """Hook to modify a batch before it is yielded."""
        return batch


class AsyncDataLoaderMixin(object):
    """Mix in FIRST: `class MyLoader(AsyncDataLoaderMixin, BaseDataLoader)`.

    `async_loader_queue_size` batches are produced ahead of time by a daemon thread; 0 disables the thread. Exceptions
    raised by the producer are re-raised in the consumer. `close_async_loader()` stops the thread."""

    def __init__(self, async_loader_queue_size=64, pin_memory=False, *args, **kwargs):
        self.async_loader_queue_size = async_loader_queue_size
        self.pin_memory = pin_memory
        super().__init__(*args, **kwargs)
        self.started = False
        if self.async_loader_queue_size > 0:
            self.finished_event = threading.Event()
            self.queue = queue.Queue(self.async_loader_queue_size)
            self.thread = threading.Thread(target=self._async_worker, daemon=True)

    def close_async_loader(self):
        """Close the async data loader."""
        if self.async_loader_queue_size > 0 and self.started:
            self.finished_event.set()
            while True:
                try:
                    self.queue.get_nowait()  # unblock a producer stuck in put()
                except queue.Empty:
                    break
            self.thread.join(timeout=10)

    def _pin(self, batch):
        try:
            import torch
        except ImportError:
            return batch
        if torch.is_tensor(batch):
            return batch.pin_memory() if not batch.is_cuda and torch.cuda.is_available() else batch
        if isinstance(batch, (list, tuple)):
            return type(batch)(self._pin(b) for b in batch)
        if isinstance(batch, dict):
            return {k: self._pin(v) for k, v in batch.items()}
        return batch

    def _async_worker(self):
        """Producer: loops over the underlying loader forever (one epoch after another) until closed."""
        try:
            while not self.finished_event.is_set():
                for batch in self._iterate():
                    if self.finished_event.is_set():
                        break
                    self.queue.put(self._pin(batch) if self.pin_memory else batch)
                self.queue.put(None)  # end-of-epoch marker
        except Exception as ex:
            self.queue.put(ex)
            self.queue.put(None)
        finally:
            self.queue.put(None)

    def __iter__(self):
        """Override the __iter__() to iterate data asynchronously to produce batches."""
        if self.async_loader_queue_size > 0:
            if not self.started:
                self.started = True
                self.thread.start()
            while True:
                batch = self.queue.get()
                if batch is None:
                    break
                if isinstance(batch, Exception):
                    raise batch
                yield self._process_batch(batch)
        else:
            for batch in self._iterate():
                yield self._process_batch(batch)
