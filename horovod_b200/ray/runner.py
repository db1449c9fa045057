"""RayExecutor: run hvd jobs on a Ray cluster.

Parity: horovod/ray/runner.py (`RayExecutor.create_settings/start/run/run_remote/execute/execute_single/shutdown`
:168-420, `Coordinator` :45-130) and strategy.py (`ColocatedStrategy` :66-137: num_hosts x num_workers_per_host with a
STRICT_SPREAD placement group; `PGStrategy` :139-225: num_workers packed by a PACK placement group).

The scheduler-independent part (rank table, rendezvous, env hand-off, ordered results) is
`horovod_b200.runner.cluster_job.ClusterJob`; this file only maps "create a worker" onto Ray actors and placement
groups.  Pass `backend=` to run the same executor on another ActorBackend (the tests use LocalProcessBackend since Ray
is not installed in this image).
"""
from dataclasses import dataclass
from typing import Optional

from horovod_b200.ray.adapter import Adapter, BaseParams
from horovod_b200.ray.worker import BaseHorovodWorker
from horovod_b200.runner.cluster_job import ActorBackend, ClusterJob


class MiniSettings:
    """What create_settings returns (reference runner.py:25-42)."""

    def __init__(self, timeout_s=30, ssh_identity_file=None, ssh_str=None, placement_group_timeout_s=100, nics=None, verbose=0):
        self.timeout_s, self.ssh_identity_file, self.ssh_str = timeout_s, ssh_identity_file, ssh_str
        self.placement_group_timeout_s, self.nics, self.verbose = placement_group_timeout_s, nics, verbose


class RayBackend(ActorBackend):
    """Ray actors inside one placement group.  `bundles` is one resource dict per worker."""

    def __init__(self, bundles, strategy='PACK', pg_timeout_s=100, use_current_placement_group=True, worker_bundle=None,
                 worker_resources=None):
        import ray
        from ray.util.placement_group import get_current_placement_group, placement_group
        self.ray = ray
        self.bundles = bundles
        # worker i runs inside bundle worker_bundle[i] with worker_resources[i] (default: one bundle per worker)
        self.worker_bundle = worker_bundle or list(range(len(bundles)))
        self.worker_resources = worker_resources or [bundles[b] for b in self.worker_bundle]
        self.pg, self._own_pg = None, False
        if use_current_placement_group:
            self.pg = get_current_placement_group()
        if self.pg is None:
            self.pg = placement_group(bundles, strategy=strategy)
            self._own_pg = True
            ready, _ = ray.wait([self.pg.ready()], timeout=pg_timeout_s)
            if not ready:
                raise TimeoutError('Placement group creation timed out. Make sure your cluster either has enough resources or use an '
                                   'autoscaling cluster. Current resources available: %s, resources requested by the placement group: %s'
                                   % (ray.available_resources(), bundles))
        self._remote_cls = ray.remote(BaseHorovodWorker)

    def create(self, index, env=None):
        from ray.util.scheduling_strategies import PlacementGroupSchedulingStrategy
        b = self.worker_resources[index]
        opts = dict(num_cpus=b.get('CPU', 1), num_gpus=b.get('GPU', 0),
                    scheduling_strategy=PlacementGroupSchedulingStrategy(placement_group=self.pg,
                                                                         placement_group_bundle_index=self.worker_bundle[index]))
        actor = self._remote_cls.options(**opts).remote(index)
        if env:
            self.ray.get(actor.update_env.remote(env))
        return actor

    def call(self, handle, method, *args, **kwargs):
        return getattr(handle, method).remote(*args, **kwargs)

    def get(self, futures, timeout=None):
        return self.ray.get(futures, timeout=timeout)

    def kill(self, handle):
        self.ray.kill(handle)

    def make_queue(self):
        from ray.util.queue import Queue
        return Queue()

    def ready(self, futures, timeout=0.0):
        done, _ = self.ray.wait(list(futures), num_returns=len(futures), timeout=timeout)
        return len(done) == len(futures)

    def shutdown(self):
        if self._own_pg and self.pg is not None:
            from ray.util.placement_group import remove_placement_group
            remove_placement_group(self.pg)
            self.pg = None


_Settings = MiniSettings


@dataclass
class StaticParams(BaseParams):
    """A job of fixed size: `num_workers` packed wherever resources are, or `num_hosts` x `num_workers_per_host`."""
    num_workers: Optional[int] = None
    num_hosts: Optional[int] = None
    num_workers_per_host: int = 1
    use_current_placement_group: bool = True

    def __post_init__(self):
        super().__post_init__()
        if self.num_workers is None and self.num_hosts is None:
            raise ValueError('Either `num_workers` or `num_hosts` must be set.')
        if self.num_workers is not None and self.num_hosts is not None:
            raise ValueError('Only one of `num_workers` and `num_hosts` may be set.')

    @property
    def elastic(self):
        return False

    @property
    def adapter(self):
        return StaticAdapter

    @property
    def total_workers(self):
        return self.num_workers if self.num_workers is not None else self.num_hosts * self.num_workers_per_host

    def placement(self):
        """(bundles, strategy, worker -> bundle index, per-worker resources)"""
        from horovod_b200.ray import strategy
        if self.num_hosts is not None:
            plan = strategy.ColocatedStrategy(self.num_hosts, self.num_workers_per_host, self.cpus_per_worker, self.gpus_per_worker)
        else:
            plan = strategy.PackStrategy(self.num_workers, self.cpus_per_worker, self.gpus_per_worker)
        return plan.describe()


class StaticAdapter(Adapter):
    """Fixed set of workers inside one placement group (reference runner.py:424-660).  The scheduler-independent part is
    `ClusterJob`; `backend=` swaps Ray for another ActorBackend."""

    def __init__(self, settings, params, env_vars=None, backend=None):
        self.settings, self.params = settings, params
        self.env_vars = dict(env_vars or {})
        self.backend, self.job = backend, None

    def start(self, executable_cls=None, executable_args=None, executable_kwargs=None, extra_env_vars=None):
        if self.backend is None:
            bundles, strat, worker_bundle, worker_res = self.params.placement()
            self.backend = RayBackend(bundles, strat, self.settings.placement_group_timeout_s,
                                      self.params.use_current_placement_group, worker_bundle, worker_res)
        env = dict(self.env_vars)
        env.update(extra_env_vars or {})
        self.job = ClusterJob(self.backend, self.params.total_workers, env=env, nics=self.settings.nics,
                              verbose=getattr(self.settings, 'verbose', 0), start_timeout=max(self.settings.timeout_s, 30)).start()
        if executable_cls is not None:
            a, k = tuple(executable_args or ()), dict(executable_kwargs or {})

            def make():
                import builtins
                builtins._hvd_ray_executable = executable_cls(*a, **k)
                return True
            self.job.run(make)

    @staticmethod
    def _on_executable(fn):
        def call():
            import builtins
            return fn(getattr(builtins, '_hvd_ray_executable', None))
        return call

    def execute(self, fn, callbacks=None):
        return self.job.run(self._on_executable(fn), callbacks=callbacks)

    def run(self, fn, args=None, kwargs=None, callbacks=None):
        return self.job.run(fn, tuple(args or ()), dict(kwargs or {}), callbacks=callbacks)

    def run_remote(self, fn, args=None, kwargs=None, callbacks=None):
        return self.job.run_remote(fn, tuple(args or ()), dict(kwargs or {}))

    def execute_single(self, fn):
        return self.job.run_single(self._on_executable(fn), 0)

    def shutdown(self):
        if self.job:
            self.job.shutdown()
            self.job = None
        if hasattr(self.backend, 'shutdown'):
            self.backend.shutdown()


class RayExecutor:
    """Job class for hvd + Ray.

    Either `num_workers` (packed wherever resources are) or `num_hosts` x `num_workers_per_host` (one bundle group per
    host, spread strictly), or — elastic — `min_workers` / `max_workers`.  `use_gpu` gives every worker `gpus_per_worker`
    GPUs; CUDA_VISIBLE_DEVICES inside an actor is what Ray sets, and hvd's local_rank indexes into it.  The keyword
    arguments become a params object (`StaticParams` / `elastic_v2.ElasticParams`) whose adapter does the work.
    """

    @classmethod
    def create_settings(cls, timeout_s=30, ssh_identity_file=None, ssh_str=None, placement_group_timeout_s=100, nics=None):
        return MiniSettings(timeout_s, ssh_identity_file, ssh_str, placement_group_timeout_s, nics)

    def __init__(self, settings=None, num_workers=None, num_hosts=None, num_workers_per_host=1, cpus_per_worker=1, use_gpu=False,
                 gpus_per_worker=None, use_current_placement_group=True, backend=None, env_vars=None, min_workers=None,
                 max_workers=None, reset_limit=None, cooldown_range=None, elastic_timeout=600, override_discovery=True,
                 elastic_actor_factory=None, elastic_queue_factory=None):
        self.settings = settings or MiniSettings()
        self.env_vars = dict(env_vars or {})
        if min_workers is not None or max_workers is not None:
            if num_workers is not None or num_hosts is not None:
                raise ValueError('`num_workers` / `num_hosts` describe a static job; use `min_workers` / `max_workers` alone for an elastic one.')
            from horovod_b200.ray.elastic_v2 import ElasticParams
            self.params = ElasticParams(cpus_per_worker=cpus_per_worker, use_gpu=use_gpu, gpus_per_worker=gpus_per_worker,
                                        min_workers=min_workers, max_workers=max_workers, reset_limit=reset_limit,
                                        cooldown_range=cooldown_range, elastic_timeout=elastic_timeout,
                                        override_discovery=override_discovery)
            self.adapter = self.params.adapter(self.settings, self.params, env_vars=self.env_vars, actor_factory=elastic_actor_factory,
                                               queue_factory=elastic_queue_factory)
        else:
            self.params = StaticParams(cpus_per_worker=cpus_per_worker, use_gpu=use_gpu, gpus_per_worker=gpus_per_worker,
                                       num_workers=num_workers, num_hosts=num_hosts, num_workers_per_host=num_workers_per_host,
                                       use_current_placement_group=use_current_placement_group)
            self.adapter = self.params.adapter(self.settings, self.params, env_vars=self.env_vars, backend=backend)

    # -- what callers and tests read ------------------------------------------------------------------------------------------
    @property
    def elastic(self):
        return self.params.elastic

    @property
    def num_workers(self):
        return self.params.min_workers if self.elastic else self.params.total_workers

    @property
    def job(self):
        return getattr(self.adapter, 'job', None)

    def _placement(self):
        return self.params.placement()

    # -- delegations ----------------------------------------------------------------------------------------------------------
    def start(self, executable_cls=None, executable_args=None, executable_kwargs=None, extra_env_vars=None):
        """Creates the workers, assigns ranks and (optionally) instantiates `executable_cls` on each of them.  An elastic
        executor (`min_workers` / `max_workers`) starts discovery + rendezvous instead; its workers are created by `run`."""
        return self.adapter.start(executable_cls, executable_args, executable_kwargs, extra_env_vars)

    def execute(self, fn, callbacks=None):
        """fn(executable) on every worker (the object created by start(executable_cls=...)); results in rank order."""
        return self.adapter.execute(fn, callbacks=callbacks)

    def run(self, fn, args=None, kwargs=None, callbacks=None):
        return self.adapter.run(fn, args=args, kwargs=kwargs, callbacks=callbacks)

    def run_remote(self, fn, args=None, kwargs=None, callbacks=None):
        """Non-blocking: returns the backend's futures (Ray ObjectRefs) in rank order."""
        return self.adapter.run_remote(fn, args=args, kwargs=kwargs, callbacks=callbacks)

    def execute_single(self, fn):
        return self.adapter.execute_single(fn)

    def shutdown(self):
        return self.adapter.shutdown()
