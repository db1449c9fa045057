"""Thread helpers (reference runner/util/threads.py)."""
import queue
import threading


def execute_function_multithreaded(fn, args_list, block_until_all_done=True, max_concurrent_executions=1000):
    """Runs fn(*args) for every args tuple in its own thread; returns {index: result}."""
    result_queue = queue.Queue()
    worker_queue = queue.Queue()
    for i, arg in enumerate(args_list):
        arg.append(i)
        worker_queue.put(arg)

    def fn_execute():
        while True:
            try:
                arg = worker_queue.get(block=False)
            except queue.Empty:
                return
            exec_index = arg[-1]
            res = fn(*arg[:-1])
            result_queue.put((exec_index, res))

    threads = []
    number_of_threads = min(max_concurrent_executions, len(args_list))
    for _ in range(number_of_threads):
        thread = in_thread(target=fn_execute, daemon=not block_until_all_done)
        threads.append(thread)
    # Returns the results only if block_until_all_done is set.
    results = None
    if block_until_all_done:
        for t in threads:
            t.join()
        results = {}
        while not result_queue.empty():
            item = result_queue.get()
            results[item[0]] = item[1]
        if len(results) != len(args_list):
            raise RuntimeError('Some threads for func {func} did not complete successfully.'.format(func=fn.__name__))
    return results


def in_thread(target, args=(), name=None, daemon=True, silent=False):
    """Executes the given function in background."""
    if not isinstance(args, tuple):
        raise ValueError('args must be a tuple, not {}, for a single argument use (arg,)'.format(type(args)))
    if silent:
        def fn(*args):
            try:
                target(*args)
            except Exception:
                pass
    else:
        fn = target
    bg = threading.Thread(target=fn, args=args, name=name)
    bg.daemon = daemon
    bg.start()
    return bg


def on_event(event, func, args=(), stop=None, check_stop_interval_s=1.0, daemon=True, silent=False):
    """Executes func(*args) once `event` is set (unless `stop` is set first)."""
    if event is None:
        raise ValueError('Event must not be None')
    if not isinstance(args, tuple):
        raise ValueError('args must be a tuple, not {}, for a single argument use (arg,)'.format(type(args)))
    if stop is None:
        def fn():
            event.wait()
            func(*args)
    else:
        def fn():
            while not event.is_set() and not stop.is_set():
                event.wait(timeout=check_stop_interval_s)
            if not stop.is_set():
                func(*args)
    return in_thread(fn, daemon=daemon, silent=silent)
