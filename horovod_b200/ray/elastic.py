"""Elastic training on Ray (parity: horovod/ray/elastic.py: `RayHostDiscovery` :39-71, `ElasticRayExecutor` :150-470).

Discovery asks Ray's global state for alive nodes and converts their CPU/GPU resources into slots; the elastic driver
(`horovod_b200.runner.elastic.driver.ElasticDriver`) then treats Ray exactly like an ssh cluster whose "spawn a worker"
callback creates an actor on the requested node instead of running ssh.
"""
import os
import queue
import threading

from horovod_b200.runner.common.util import timeout as _timeout_mod
from horovod_b200.runner.elastic.discovery import HostDiscovery
from horovod_b200.runner.elastic.driver import ElasticDriver
from horovod_b200.runner.elastic.rendezvous import create_rendezvous_handler
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.mesh_run import create_run_env_vars, create_slot_env_vars


class RayHostDiscovery(HostDiscovery):
    """Hosts and slots from `ray.nodes()`."""

    def __init__(self, use_gpu=False, cpus_per_slot=1, gpus_per_slot=1, nodes_fn=None):
        self.use_gpu, self.cpus_per_slot, self.gpus_per_slot = use_gpu, cpus_per_slot, gpus_per_slot
        self._nodes_fn = nodes_fn

    def _nodes(self):
        if self._nodes_fn is not None:
            return self._nodes_fn()
        import ray
        return ray.nodes()

    def find_available_hosts_and_slots(self):
        mapping = {}
        for node in self._nodes():
            if not node.get('alive', node.get('Alive', False)):
                continue
            res = node.get('Resources', {})
            host = node.get('NodeManagerAddress') or node.get('NodeManagerHostname')
            slots = int(res.get('GPU', 0) // self.gpus_per_slot) if self.use_gpu else int(res.get('CPU', 0) // self.cpus_per_slot)
            if host and slots > 0:
                mapping[host] = slots
        return mapping


class ElasticRayExecutor:
    """Runs an elastic job: workers are (re)created as Ray actors whenever discovery reports a change.

    `run(worker_fn)` returns the list of values returned by the workers of the final, successful round.
    """

    @staticmethod
    def create_settings(min_num_proc=1, max_num_proc=None, reset_limit=None, elastic_timeout=600, timeout_s=30,
                        ssh_identity_file=None, nics=None, **kwargs):
        from horovod_b200.runner.elastic.settings import ElasticSettings
        start_timeout = _timeout_mod.Timeout(timeout_s, message='Timed out waiting for {activity}. Please check connectivity between servers.')
        return ElasticSettings(discovery=None, min_num_proc=min_num_proc, max_num_proc=max_num_proc, elastic_timeout=elastic_timeout,
                               reset_limit=reset_limit, num_proc=min_num_proc, ssh_identity_file=ssh_identity_file, nics=nics,
                               start_timeout=start_timeout, **kwargs)

    def __init__(self, settings, use_gpu=False, cpus_per_slot=1, gpus_per_slot=None, env_vars=None, override_discovery=True,
                 actor_factory=None, queue_factory=None):
        if gpus_per_slot and not use_gpu:
            raise ValueError('gpus_per_slot is set, but use_gpu is False. use_gpu must be True if gpus_per_slot is set.')
        gpus_per_slot = gpus_per_slot or 1
        if override_discovery:
            settings.discovery = RayHostDiscovery(use_gpu=use_gpu, cpus_per_slot=cpus_per_slot, gpus_per_slot=gpus_per_slot)
        self.settings, self.use_gpu = settings, use_gpu
        self.cpus_per_slot, self.gpus_per_slot = cpus_per_slot, gpus_per_slot
        self.env_vars = dict(env_vars or {})
        self.driver, self.rendezvous = None, None
        self._actor_factory = actor_factory  # (hostname, env) -> object with .execute(fn) [tests]; default: Ray actor
        self._queue_factory = queue_factory  # () -> picklable queue for run(callbacks=...); default: ray.util.queue.Queue

    def start(self):
        self.rendezvous = RendezvousServer(self.settings.verbose)
        self.driver = ElasticDriver(self.rendezvous, self.settings.discovery, self.settings.min_num_proc, self.settings.max_num_proc,
                                    timeout=self.settings.elastic_timeout, reset_limit=self.settings.reset_limit,
                                    cooldown_range=getattr(self.settings, 'cooldown_range', None), verbose=self.settings.verbose)
        port = self.rendezvous.start_server()
        create_rendezvous_handler(self.driver).install(self.rendezvous)
        self.driver.wait_for_available_slots(self.settings.min_num_proc)
        self._run_env = create_run_env_vars(_driver_ip(), port, nics=self.settings.nics, elastic=True)

    def _make_actor(self, hostname, env):
        if self._actor_factory is not None:
            return self._actor_factory(hostname, env)
        import ray
        from horovod_b200.runner.cluster_job import WorkerActor
        res = {f'node:{hostname}': 0.01}
        cls = ray.remote(num_cpus=self.cpus_per_slot, num_gpus=self.gpus_per_slot if self.use_gpu else 0, resources=res)(WorkerActor)
        actor = cls.remote()
        ray.get(actor.update_env.remote(env))

        class _H:
            def execute(self_inner, fn):
                return ray.get(actor.execute.remote(fn))

            def kill(self_inner):
                ray.kill(actor)
        return _H()

    def _log_queue(self):
        if self._queue_factory is not None:
            return self._queue_factory()
        from ray.util.queue import Queue
        return Queue()

    def run(self, worker_fn, callbacks=None):
        """`callbacks`: every dict a worker hands to `horovod_b200.ray.ray_logger.log` is passed to each of them on the
        driver while the job runs (reference elastic_v2.py `_process_calls`)."""
        results_q = queue.Queue()
        errors = []                              # (hostname, rank, exception) of every worker that failed, in order
        stop_drain = threading.Event()
        if callbacks:
            import functools
            from horovod_b200.runner.cluster_job import _with_log_queue
            log_q = self._log_queue()
            worker_fn = functools.partial(_with_log_queue, log_q, worker_fn, (), None)

            def drain_once():
                while not log_q.empty():
                    item = log_q.get()
                    for cb in callbacks:
                        cb(item)

            def drain_loop():
                while not stop_drain.wait(0.1):
                    drain_once()
                drain_once()
            drainer = threading.Thread(target=drain_loop, daemon=True)
            drainer.start()

        def spawn(slot_info, events):
            env = dict(self.env_vars)
            env.update(self._run_env)
            env.update(create_slot_env_vars(slot_info))
            env['HOROVOD_ELASTIC'] = '1'
            actor = self._make_actor(slot_info.hostname, env)
            done = threading.Event()
            box = {}

            def body():
                try:
                    box['value'] = actor.execute(worker_fn)
                    box['code'] = 0
                except Exception as e:  # a dead actor is a failed worker: the driver blacklists / re-plans
                    box['code'], box['error'] = 1, e
                    errors.append((slot_info.hostname, slot_info.rank, e))
                done.set()
            threading.Thread(target=body, daemon=True).start()
            while not done.wait(0.1):
                if any(e.is_set() for e in events):
                    if hasattr(actor, 'kill'):
                        actor.kill()
                    return 1, 0
            if box['code'] == 0:
                # the rank of a worker changes with every reset: key the result by the slot (host, local rank) it ran in and
                # order by the driver's FINAL rank table below
                results_q.put(((slot_info.hostname, slot_info.local_rank), slot_info.rank, box['value']))
            return box['code'], 0

        self.driver.start(self.settings.num_proc or self.settings.min_num_proc, spawn)
        res = self.driver.get_results()
        self.driver.stop()
        self.rendezvous.stop()
        if callbacks:
            stop_drain.set()
            drainer.join(timeout=10)
        def first_error():
            if not errors:
                return ''
            host, rank, e = errors[0]
            return '\nfirst worker failure (rank %s on %s): %s' % (rank, host, e)
        if res.error_message:
            raise RuntimeError(res.error_message + first_error())
        out = {}
        while not results_q.empty():
            (host, local_rank), spawn_rank, v = results_q.get()
            try:
                final = self.driver.get_slot_info(host, local_rank).rank if self.driver.has_rank_assignment(host, local_rank) else None
            except Exception:  # noqa: BLE001 - a slot that left the final layout
                final = None
            out[(0, final) if final is not None else (1, spawn_rank, host, local_rank)] = v
        if not out:
            raise RuntimeError('the elastic job ended without a single successful worker' + first_error())
        return [out[k] for k in sorted(out)]


def _driver_ip():
    from horovod_b200.runner.cluster_job import _routable_ip
    return os.environ.get('HVD_DRIVER_IP', _routable_ip())
