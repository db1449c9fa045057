"""Worker -> driver message channel (reference horovod/ray/ray_logger.py:14-28).

`RayExecutor.run(fn, callbacks=[...])` hands every worker a queue; inside the training function
`ray_logger.log({'loss': ...})` puts the dict on that queue and the driver feeds each item to the callbacks while it waits
for the workers.  Without a configured queue `log` is a no-op, so the same training function also runs under `hvdrun`."""
_queue = None


def configure(queue):
    """Installs the queue of this worker process (called by the executor before the training function)."""
    global _queue
    _queue = queue


def log(info_dict):
    """Sends `info_dict` to the driver's callbacks (no-op when the job was started without callbacks)."""
    if _queue is None:
        return False
    _queue.put(info_dict)
    return True
