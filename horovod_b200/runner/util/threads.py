"""Small threading helpers of the launcher (API parity: horovod/runner/util/threads.py: `execute_function_multithreaded`,
`in_thread`, `on_event`)."""
import threading
from concurrent.futures import ThreadPoolExecutor


def execute_function_multithreaded(fn, args_list, block_until_all_done=True, max_concurrent_executions=1000):
    """fn(*args) for every args list, concurrently.  Blocking: returns {position in args_list: result} and raises if any
    call raised.  Non-blocking: the calls keep running on daemon threads and None is returned."""
    calls = [tuple(a) for a in args_list]
    if not block_until_all_done:
        for c in calls:
            in_thread(fn, args=c, daemon=True)
        return None
    if not calls:
        return {}
    with ThreadPoolExecutor(max_workers=max(1, min(max_concurrent_executions, len(calls)))) as pool:
        futures = [pool.submit(fn, *c) for c in calls]
    failed = [f for f in futures if f.exception() is not None]
    if failed:
        raise RuntimeError('Some threads for func {func} did not complete successfully.'.format(func=getattr(fn, '__name__', fn))) \
            from failed[0].exception()
    return {i: f.result() for i, f in enumerate(futures)}


def _require_tuple(args):
    if not isinstance(args, tuple):
        raise ValueError('args must be a tuple, not {}, for a single argument use (arg,)'.format(type(args)))


def in_thread(target, args=(), name=None, daemon=True, silent=False):
    """Starts target(*args) on a new thread and returns the thread.  `silent` swallows exceptions of the target."""
    _require_tuple(args)

    def quiet(*a):
        try:
            target(*a)
        except Exception:  # noqa: BLE001 - by request of the caller
            pass
    t = threading.Thread(target=quiet if silent else target, args=args, name=name, daemon=daemon)
    t.start()
    return t


def on_event(event, func, args=(), stop=None, check_stop_interval_s=1.0, daemon=True, silent=False):
    """Runs func(*args) on a background thread once `event` is set; if `stop` is given and set first, never runs it."""
    if event is None:
        raise ValueError('Event must not be None')
    _require_tuple(args)

    def waiter():
        if stop is None:
            event.wait()
        else:
            while not (event.is_set() or stop.is_set()):
                event.wait(check_stop_interval_s)
            if stop.is_set() and not event.is_set():
                return
        func(*args)
    return in_thread(waiter, daemon=daemon, silent=silent)
