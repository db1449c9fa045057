"""Scheduler-agnostic job core shared by the Ray and Spark integrations.

Both integrations do the same four things (reference: ray/runner.py `Coordinator` :45-130 + `StaticAdapter` :424-660;
spark/runner.py `_make_spark_thread`/`_launch_job` + spark/driver/driver_service.py): (1) obtain N long-lived workers
from a cluster scheduler, (2) learn which worker sits on which host, (3) derive rank / local_rank / cross_rank so that
ranks on one host are contiguous, start the rendezvous KV server and hand every worker its environment, (4) run a
function on all workers and return the results in rank order.  Only step (1) is scheduler specific, so it is factored
behind `ActorBackend`; the Ray backend creates Ray actors, the Spark backend barrier-stage tasks, and
`LocalProcessBackend` plain subprocesses (used by the tests and handy for debugging without a cluster).
"""
import multiprocessing as mp
import os
import socket
import traceback
from collections import OrderedDict

from horovod_b200.runner.common.util.hosts import SlotInfo
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.mesh_run import create_run_env_vars, create_slot_env_vars


class WorkerActor:
    """What runs inside every worker (a Ray actor, a Spark task or a subprocess)."""

    def __init__(self, index=0):
        self.index = index

    def hostname(self):
        return socket.gethostname()

    def node_id(self):
        return os.environ.get('HVD_NODE_ID_OVERRIDE', socket.gethostname())

    def gpu_ids(self):
        v = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        return [x for x in v.split(',') if x]

    def update_env(self, env):
        os.environ.update({k: str(v) for k, v in env.items()})
        return True

    def env(self):
        return dict(os.environ)

    def execute(self, fn, args=(), kwargs=None):
        return fn(*args, **(kwargs or {}))


class ActorBackend:
    """Creates workers and calls methods on them.  `call` returns an opaque future; `get` resolves a list of them."""

    def create(self, index, env=None):
        raise NotImplementedError

    def call(self, handle, method, *args, **kwargs):
        raise NotImplementedError

    def get(self, futures, timeout=None):
        raise NotImplementedError

    def kill(self, handle):
        raise NotImplementedError

    # -- optional: needed only for run(..., callbacks=[...]) ------------------------------------------------------------
    def make_queue(self):
        """A queue object that can be pickled into the workers (worker -> driver messages)."""
        raise NotImplementedError('%s has no worker -> driver queue: run() cannot take callbacks' % type(self).__name__)

    def ready(self, futures, timeout=0.0):
        """True when every future is resolved; waits at most `timeout` seconds."""
        raise NotImplementedError


def _with_log_queue(queue, fn, args, kwargs):
    """Runs inside a worker: whatever `horovod_b200.ray.ray_logger.log(...)` receives travels to the driver's callbacks."""
    from horovod_b200.ray import ray_logger
    ray_logger.configure(queue)
    try:
        return fn(*args, **(kwargs or {}))
    finally:
        ray_logger.configure(None)


# ---- local subprocess backend ----------------------------------------------------------------------------------------
def _proc_main(index, conn, env):
    os.environ.update(env or {})
    actor = WorkerActor(index)
    while True:
        try:
            msg = conn.recv()
        except EOFError:
            return
        if msg is None:
            return
        import pickle
        method, args, kwargs = pickle.loads(msg)  # cloudpickle payload: lambdas / closures / __main__ functions work
        try:
            conn.send(('ok', getattr(actor, method)(*args, **kwargs)))
        except BaseException as e:  # reported to the driver, the worker keeps serving
            conn.send(('err', '%s: %s\n%s' % (type(e).__name__, e, traceback.format_exc())))


class _LocalFuture:
    def __init__(self, conn):
        self.conn = conn


class LocalProcessBackend(ActorBackend):
    """N subprocesses on this machine; `node_ids` fakes a multi-host layout (index -> node name)."""

    def __init__(self, node_ids=None, start_method='spawn'):
        self.ctx = mp.get_context(start_method)
        self.node_ids = node_ids or {}
        self._manager = None

    def make_queue(self):
        if self._manager is None:
            self._manager = self.ctx.Manager()
        return self._manager.Queue()

    def ready(self, futures, timeout=0.0):
        import time
        from multiprocessing.connection import wait
        deadline = time.monotonic() + (timeout or 0.0)
        pending = [f.conn for f in futures]
        while True:
            pending = [c for c in pending if not c.poll(0)]
            left = deadline - time.monotonic()
            if not pending or left <= 0:
                return not pending
            wait(pending, left)

    def shutdown(self):
        if self._manager is not None:
            self._manager.shutdown()
            self._manager = None

    def create(self, index, env=None):
        parent, child = self.ctx.Pipe()
        env = dict(env or {})
        if index in self.node_ids:
            env['HVD_NODE_ID_OVERRIDE'] = self.node_ids[index]
        p = self.ctx.Process(target=_proc_main, args=(index, child, env), daemon=True)
        p.start()
        child.close()
        return (p, parent)

    def call(self, handle, method, *args, **kwargs):
        import cloudpickle
        handle[1].send(cloudpickle.dumps((method, args, kwargs)))
        return _LocalFuture(handle[1])

    def get(self, futures, timeout=None):
        out, first_error = [], None
        for f in futures:  # every reply is consumed even after a failure, so the pipes stay in step with the calls
            if not f.conn.poll(timeout):
                raise TimeoutError('worker did not answer within %s s' % timeout)
            status, val = f.conn.recv()
            if status != 'ok' and first_error is None:
                first_error = val
            out.append(val)
        if first_error is not None:
            raise RuntimeError('worker raised:\n' + first_error)
        return out

    def kill(self, handle):
        p, conn = handle
        try:
            conn.send(None)
        except Exception:
            pass
        p.join(timeout=3)
        if p.is_alive():
            p.terminate()


# ---- rank assignment ---------------------------------------------------------------------------------------------------
def assign_slots(node_ids):
    """node_ids[i] = node of worker i.  Returns SlotInfo per worker such that the workers of a node get contiguous
    ranks (node order = first appearance, worker order within a node = creation order)."""
    by_node = OrderedDict()
    for i, n in enumerate(node_ids):
        by_node.setdefault(n, []).append(i)
    size = len(node_ids)
    local_sizes = {n: len(w) for n, w in by_node.items()}
    slots = [None] * size
    rank = 0
    for cross_rank, (n, workers) in enumerate(by_node.items()):
        for local_rank, w in enumerate(workers):
            # cross_size of a local_rank = number of nodes that have such a local rank (heterogeneous clusters)
            cross_size = sum(1 for m in by_node.values() if len(m) > local_rank)
            my_cross = sum(1 for m in list(by_node.values())[:cross_rank] if len(m) > local_rank)
            slots[w] = SlotInfo(hostname=n, rank=rank, local_rank=local_rank, cross_rank=my_cross, size=size,
                                local_size=local_sizes[n], cross_size=cross_size)
            rank += 1
    return slots


class ClusterJob:
    """Driver-side object: owns the workers, the rendezvous server and the rank table."""

    def __init__(self, backend, num_workers, env=None, nics=None, verbose=0, start_timeout=60):
        self.backend, self.num_workers = backend, num_workers
        self.env, self.nics, self.verbose, self.start_timeout = dict(env or {}), nics, verbose, start_timeout
        self.handles, self.slots, self.rendezvous = [], [], None

    def start(self, extra_env_fn=None):
        self.handles = [self.backend.create(i, self.env) for i in range(self.num_workers)]
        node_ids = self.backend.get([self.backend.call(h, 'node_id') for h in self.handles], self.start_timeout)
        self.slots = assign_slots(node_ids)
        self.rendezvous = RendezvousServer(self.verbose)
        port = self.rendezvous.start_server()
        self.rendezvous.init(self.slots)
        driver_ip = _routable_ip() if len(set(node_ids)) > 1 else '127.0.0.1'
        common = create_run_env_vars(driver_ip, port, nics=self.nics, elastic=False)
        futures = []
        for h, slot in zip(self.handles, self.slots):
            env = dict(common)
            env.update(create_slot_env_vars(slot))
            if extra_env_fn:
                env.update(extra_env_fn(slot) or {})
            futures.append(self.backend.call(h, 'update_env', env))
        self.backend.get(futures, self.start_timeout)
        return self

    def by_rank(self):
        order = sorted(range(self.num_workers), key=lambda i: self.slots[i].rank)
        return [self.handles[i] for i in order]

    def run(self, fn, args=(), kwargs=None, timeout=None, callbacks=None):
        """fn(*args, **kwargs) on every worker; results in rank order.  With `callbacks`, every dict a worker passes to
        `horovod_b200.ray.ray_logger.log` is handed to each callback on the driver while the workers run."""
        if not callbacks:
            futures = [self.backend.call(h, 'execute', fn, args, kwargs) for h in self.by_rank()]
            return self.backend.get(futures, timeout)
        q = self.backend.make_queue()
        futures = [self.backend.call(h, 'execute', _with_log_queue, (q, fn, tuple(args), kwargs)) for h in self.by_rank()]

        def drain():
            while not q.empty():
                item = q.get()
                for cb in callbacks:
                    cb(item)
        import time
        t0 = time.monotonic()
        while not self.backend.ready(futures, 0.1):
            drain()
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError('workers did not finish within %s s' % timeout)
        out = self.backend.get(futures, timeout)
        drain()
        return out

    def run_remote(self, fn, args=(), kwargs=None):
        return [self.backend.call(h, 'execute', fn, args, kwargs) for h in self.by_rank()]

    def run_single(self, fn, rank=0, args=(), kwargs=None, timeout=None):
        return self.backend.get([self.backend.call(self.by_rank()[rank], 'execute', fn, args, kwargs)], timeout)[0]

    def shutdown(self):
        for h in self.handles:
            try:
                self.backend.kill(h)
            except Exception:
                pass
        self.handles = []
        if self.rendezvous:
            self.rendezvous.stop()
            self.rendezvous = None


def _routable_ip():
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    try:
        s.connect(('10.255.255.255', 1))
        return s.getsockname()[0]
    except Exception:
        return '127.0.0.1'
    finally:
        s.close()


# ---- connect-back backend: workers are tasks launched by somebody else's scheduler -------------------------------------------
def serve_connect_back(index, address, authkey, env=None):
    """Body of a scheduler-launched task (e.g. a Spark barrier task): dial the driver, then serve actor calls until the
    driver hangs up.  Returns when the job is over."""
    from multiprocessing.connection import Client
    conn = Client(tuple(address), authkey=authkey)
    conn.send(('hello', index))
    _proc_main(index, conn, env or {})
    return index


def _task_entry(i, env=None, address=None, authkey=None):
    return serve_connect_back(i, address, authkey, env)


class ConnectBackBackend(ActorBackend):
    """`launch(n, task_main)` must start n tasks somewhere that each call `task_main(i)`; the tasks dial back to a
    Listener owned by the driver and from then on behave like actors."""

    def __init__(self, launch, num_workers, host=None, timeout=120):
        import secrets
        from multiprocessing.connection import Listener
        self.authkey = secrets.token_bytes(16)
        self.listener = Listener((host or _routable_ip(), 0), authkey=self.authkey)
        self.address = self.listener.address
        self.num_workers, self.timeout = num_workers, timeout
        self._conns = {}
        import functools
        task_main = functools.partial(_task_entry, address=tuple(self.address), authkey=self.authkey)  # picklable
        self._launch_handle = launch(num_workers, task_main)
        self._accept_all()

    def _accept_all(self):
        import threading
        err = []

        def accept():
            try:
                while len(self._conns) < self.num_workers:
                    c = self.listener.accept()
                    tag, idx = c.recv()
                    assert tag == 'hello'
                    self._conns[idx] = c
            except Exception as e:  # listener closed on timeout
                err.append(e)
        t = threading.Thread(target=accept, daemon=True)
        t.start()
        t.join(self.timeout)
        if t.is_alive() or len(self._conns) < self.num_workers:
            self.listener.close()
            raise TimeoutError('only %d of %d tasks connected back within %s s' % (len(self._conns), self.num_workers, self.timeout))

    def create(self, index, env=None):
        conn = self._conns[index]
        if env:
            self.get([self.call((None, conn), 'update_env', env)], self.timeout)
        return (None, conn)

    def call(self, handle, method, *args, **kwargs):
        import cloudpickle
        handle[1].send(cloudpickle.dumps((method, args, kwargs)))
        return _LocalFuture(handle[1])

    get = LocalProcessBackend.get

    def kill(self, handle):
        try:
            handle[1].send(None)
            handle[1].close()
        except Exception:
            pass

    def shutdown(self):
        try:
            self.listener.close()
        except Exception:
            pass


# ---- elastic flavour: tasks come and go -----------------------------------------------------------------------------------------
class ConnectBackPool:
    """Tasks launched by someone else's scheduler (Spark) dial back whenever they start — the first attempt of a task, or the
    attempt Spark schedules after a failure — and are handed to the elastic driver as slots: discovery reports how many live
    tasks sit on which node, "spawning a worker" means taking an idle task of the requested node.

    `launch(n, task_main)` starts n tasks like ConnectBackBackend's launcher; unlike there, connections are accepted for the
    whole life of the pool."""

    class Task:
        def __init__(self, index, conn, node):
            self.index, self.conn, self.node = index, conn, node
            self.busy, self.alive = False, True

    def __init__(self, launch, num_tasks, host=None, timeout=120):
        import functools
        import secrets
        import threading
        from multiprocessing.connection import Listener
        self.authkey = secrets.token_bytes(16)
        self.listener = Listener((host or _routable_ip(), 0), authkey=self.authkey)
        self.address = self.listener.address
        self.timeout = timeout
        self._lock = threading.Condition()
        self._tasks = []
        self._closed = False
        task_main = functools.partial(_task_entry, address=tuple(self.address), authkey=self.authkey)
        self._acceptor = threading.Thread(target=self._accept_loop, name='hvd-pool-accept', daemon=True)
        self._acceptor.start()
        self._launch_handle = launch(num_tasks, task_main)

    # -- connections ----------------------------------------------------------------------------------------------------------
    def _accept_loop(self):
        import cloudpickle
        while not self._closed:
            try:
                conn = self.listener.accept()
                tag, idx = conn.recv()
                assert tag == 'hello'
                conn.send(cloudpickle.dumps(('node_id', (), {})))
                if not conn.poll(self.timeout):
                    conn.close()
                    continue
                status, node = conn.recv()
                if status != 'ok':
                    conn.close()
                    continue
            except Exception:  # noqa: BLE001 - listener closed, or a task that died while saying hello
                if self._closed:
                    return
                continue
            with self._lock:
                self._tasks.append(ConnectBackPool.Task(idx, conn, node))
                self._lock.notify_all()

    def _sweep(self):
        """Marks idle tasks whose connection is gone (callers hold the lock)."""
        for t in self._tasks:
            if t.alive and not t.busy:
                try:
                    if t.conn.poll(0):          # an idle task never sends: readable means EOF (or garbage)
                        t.conn.recv()
                        t.alive = False
                except (EOFError, OSError):
                    t.alive = False

    def wait_for(self, count, timeout):
        import time
        deadline = time.monotonic() + timeout
        with self._lock:
            while True:
                self._sweep()
                alive = sum(1 for t in self._tasks if t.alive)
                if alive >= count:
                    return alive
                left = deadline - time.monotonic()
                if left <= 0:
                    raise TimeoutError('only %d of %d tasks connected back within %s s' % (alive, count, timeout))
                self._lock.wait(min(left, 0.2))

    def hosts_and_slots(self):
        with self._lock:
            self._sweep()
            out = OrderedDict()
            for t in self._tasks:
                if t.alive:
                    out[t.node] = out.get(t.node, 0) + 1
            return out

    # -- workers ----------------------------------------------------------------------------------------------------------------
    def actor_factory(self, hostname, env):
        """(hostname, env) -> object with execute(fn) / kill(): an idle task of that node, with `env` applied."""
        import time
        deadline = time.monotonic() + self.timeout
        with self._lock:
            while True:
                self._sweep()
                task = next((t for t in self._tasks if t.alive and not t.busy and t.node == hostname), None)
                if task is not None:
                    task.busy = True
                    break
                left = deadline - time.monotonic()
                if left <= 0:
                    raise RuntimeError('no idle task on %s' % hostname)
                self._lock.wait(min(left, 0.2))
        pool = self

        class Worker:
            def _call(self, method, *args):
                import cloudpickle
                try:
                    task.conn.send(cloudpickle.dumps((method, args, {})))
                    while not task.conn.poll(0.5):
                        if not task.alive:
                            raise RuntimeError('task %s on %s was stopped' % (task.index, task.node))
                    status, value = task.conn.recv()
                except (EOFError, OSError) as e:
                    task.alive = False
                    raise RuntimeError('task %s on %s died: %s' % (task.index, task.node, e or type(e).__name__))
                if status != 'ok':
                    raise RuntimeError('worker raised:\n' + str(value))
                return value

            def execute(self, fn):
                try:
                    self._call('update_env', env)
                    return self._call('execute', fn)
                finally:
                    with pool._lock:
                        task.busy = False
                        pool._lock.notify_all()

            def kill(self):
                task.alive = False
                try:
                    task.conn.close()            # the task's serve loop ends on EOF: the scheduler sees the task finish
                except Exception:  # noqa: BLE001
                    pass
        return Worker()

    def shutdown(self):
        self._closed = True
        try:
            self.listener.close()
        except Exception:  # noqa: BLE001
            pass
        with self._lock:
            for t in self._tasks:
                try:
                    if t.alive:
                        t.conn.send(None)
                    t.conn.close()
                except Exception:  # noqa: BLE001
                    pass
                t.alive = False
