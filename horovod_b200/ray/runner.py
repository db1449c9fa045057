"""RayExecutor: run hvd jobs on a Ray cluster.

Parity: horovod/ray/runner.py (`RayExecutor.create_settings/start/run/run_remote/execute/execute_single/shutdown`
:168-420, `Coordinator` :45-130) and strategy.py (`ColocatedStrategy` :66-137: num_hosts x num_workers_per_host with a
STRICT_SPREAD placement group; `PGStrategy` :139-225: num_workers packed by a PACK placement group).

The scheduler-independent part (rank table, rendezvous, env hand-off, ordered results) is
`horovod_b200.runner.cluster_job.ClusterJob`; this file only maps "create a worker" onto Ray actors and placement
groups.  Pass `backend=` to run the same executor on another ActorBackend (the tests use LocalProcessBackend since Ray
is not installed in this image).
"""
from horovod_b200.runner.cluster_job import ActorBackend, ClusterJob, WorkerActor


class _Settings:
    """What create_settings returns (reference MiniSettings)."""

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
        self._remote_cls = ray.remote(WorkerActor)

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

    def shutdown(self):
        if self._own_pg and self.pg is not None:
            from ray.util.placement_group import remove_placement_group
            remove_placement_group(self.pg)
            self.pg = None


class RayExecutor:
    """Job class for hvd + Ray.

    Either `num_workers` (packed wherever resources are) or `num_hosts` x `num_workers_per_host` (one bundle group per
    host, spread strictly).  `use_gpu` gives every worker `gpus_per_worker` GPUs; CUDA_VISIBLE_DEVICES inside an
    actor is what Ray sets, and hvd's local_rank indexes into it.
    """

    @classmethod
    def create_settings(cls, timeout_s=30, ssh_identity_file=None, ssh_str=None, placement_group_timeout_s=100, nics=None):
        return _Settings(timeout_s, ssh_identity_file, ssh_str, placement_group_timeout_s, nics)

    def __init__(self, settings=None, num_workers=None, num_hosts=None, num_workers_per_host=1, cpus_per_worker=1, use_gpu=False,
                 gpus_per_worker=None, use_current_placement_group=True, backend=None, env_vars=None, min_workers=None,
                 max_workers=None, reset_limit=None, cooldown_range=None, elastic_timeout=600, override_discovery=True,
                 elastic_actor_factory=None):
        self.elastic = min_workers is not None or max_workers is not None
        if self.elastic:
            if num_workers is not None or num_hosts is not None:
                raise ValueError('`num_workers` / `num_hosts` describe a static job; use `min_workers` / `max_workers` alone for an elastic one.')
            if min_workers is None or min_workers < 1:
                raise ValueError('`min_workers` must be provided (>= 1) for an elastic job.')
            if max_workers is not None and max_workers < min_workers:
                raise ValueError('`max_workers` (%s) must not be smaller than `min_workers` (%s).' % (max_workers, min_workers))
            self._elastic_args = dict(min_workers=min_workers, max_workers=max_workers, reset_limit=reset_limit,
                                      cooldown_range=cooldown_range, elastic_timeout=elastic_timeout,
                                      override_discovery=override_discovery, actor_factory=elastic_actor_factory)
            num_workers = min_workers
        if num_workers is None and num_hosts is None:
            raise ValueError('Either `num_workers` or `num_hosts` must be set.')
        if num_workers is not None and num_hosts is not None:
            raise ValueError('Only one of `num_workers` and `num_hosts` may be set.')
        if gpus_per_worker and not use_gpu:
            raise ValueError('gpus_per_worker is set, but use_gpu is False. use_gpu must be True if gpus_per_worker is set.')
        if use_gpu and isinstance(gpus_per_worker, int) and gpus_per_worker < 1:
            raise ValueError(f'gpus_per_worker must be >= 1: Got {gpus_per_worker}.')
        self.settings = settings or _Settings()
        self.colocated = num_hosts is not None
        self.num_workers = num_workers if num_workers is not None else num_hosts * num_workers_per_host
        self.num_hosts, self.num_workers_per_host = num_hosts, num_workers_per_host
        self.cpus_per_worker, self.use_gpu = cpus_per_worker, use_gpu
        self.gpus_per_worker = (gpus_per_worker or 1) if use_gpu else 0
        self.use_current_placement_group = use_current_placement_group
        self.env_vars = dict(env_vars or {})
        self._backend, self.job = backend, None

    def _placement(self):
        """(bundles, strategy, worker -> bundle index, per-worker resources)"""
        from horovod_b200.ray import strategy
        if self.colocated:
            plan = strategy.ColocatedStrategy(self.num_hosts, self.num_workers_per_host, self.cpus_per_worker, self.gpus_per_worker)
        else:
            plan = strategy.PackStrategy(self.num_workers, self.cpus_per_worker, self.gpus_per_worker)
        return plan.describe()

    def _start_elastic(self, extra_env_vars):
        from horovod_b200.ray.elastic import ElasticRayExecutor
        a = self._elastic_args
        settings = ElasticRayExecutor.create_settings(min_num_proc=a['min_workers'], max_num_proc=a['max_workers'],
                                                      reset_limit=a['reset_limit'], elastic_timeout=a['elastic_timeout'],
                                                      timeout_s=self.settings.timeout_s, nics=self.settings.nics,
                                                      **({'cooldown_range': a['cooldown_range']} if a['cooldown_range'] else {}))
        if not a['override_discovery']:
            settings.discovery = getattr(self.settings, 'discovery', None)
        env = dict(self.env_vars)
        env.update(extra_env_vars or {})
        self._elastic_executor = ElasticRayExecutor(settings, use_gpu=self.use_gpu, cpus_per_slot=self.cpus_per_worker,
                                                    gpus_per_slot=self.gpus_per_worker or None, env_vars=env,
                                                    override_discovery=a['override_discovery'], actor_factory=a['actor_factory'])
        self._elastic_executor.start()

    def start(self, executable_cls=None, executable_args=None, executable_kwargs=None, extra_env_vars=None):
        """Creates the workers, assigns ranks and (optionally) instantiates `executable_cls` on each of them.  An elastic
        executor (`min_workers` / `max_workers`) starts discovery + rendezvous instead; its workers are created by `run`."""
        if self.elastic:
            if executable_cls is not None:
                raise ValueError('executable_cls is not supported by the elastic executor: workers come and go between resets.')
            return self._start_elastic(extra_env_vars)
        backend = self._backend
        if backend is None:
            bundles, strat, worker_bundle, worker_res = self._placement()
            backend = RayBackend(bundles, strat, self.settings.placement_group_timeout_s, self.use_current_placement_group,
                                 worker_bundle, worker_res)
            self._backend = backend
        env = dict(self.env_vars)
        env.update(extra_env_vars or {})
        self.job = ClusterJob(backend, self.num_workers, env=env, nics=self.settings.nics, verbose=getattr(self.settings, 'verbose', 0),
                              start_timeout=max(self.settings.timeout_s, 30)).start()
        self._has_executable = executable_cls is not None
        if executable_cls is not None:
            a, k = tuple(executable_args or ()), dict(executable_kwargs or {})

            def make():
                import builtins
                builtins._hvd_ray_executable = executable_cls(*a, **k)
                return True
            self.job.run(make)

    def execute(self, fn):
        """fn(executable) on every worker (the object created by start(executable_cls=...)); results in rank order."""
        def call():
            import builtins
            return fn(getattr(builtins, '_hvd_ray_executable', None))
        return self.job.run(call)

    def run(self, fn, args=None, kwargs=None, callbacks=None):
        if self.elastic:
            import functools
            return self._elastic_executor.run(functools.partial(fn, *tuple(args or ()), **dict(kwargs or {})), callbacks=callbacks)
        return self.job.run(fn, tuple(args or ()), dict(kwargs or {}))

    def run_remote(self, fn, args=None, kwargs=None):
        """Non-blocking: returns the backend's futures (Ray ObjectRefs) in rank order."""
        return self.job.run_remote(fn, tuple(args or ()), dict(kwargs or {}))

    def execute_single(self, fn):
        def call():
            import builtins
            return fn(getattr(builtins, '_hvd_ray_executable', None))
        return self.job.run_single(call, 0)

    def shutdown(self):
        if self.job:
            self.job.shutdown()
            self.job = None
        if hasattr(self._backend, 'shutdown'):
            self._backend.shutdown()
