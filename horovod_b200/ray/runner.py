"""RayExecutor: start N Ray actors, give each its rank environment + the address of the rendezvous KV server hosted by
the driver, and run functions on all of them (role parity: horovod/ray/runner.py RayExecutor / Coordinator)."""
from horovod_b200.ray import strategy
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.util import network


class RayExecutor(object):
    def __init__(self, settings=None, num_workers=None, num_hosts=None, num_workers_per_host=1, cpus_per_worker=1,
                 use_gpu=False, gpus_per_worker=None):
        if num_workers is None and num_hosts is None:
            raise ValueError('Either `num_workers` or `num_hosts` must be specified.')
        self.settings = settings
        self.num_workers = num_workers
        self.num_hosts = num_hosts
        self.num_workers_per_host = num_workers_per_host
        self.cpus_per_worker = cpus_per_worker
        self.use_gpu = use_gpu
        self.gpus_per_worker = gpus_per_worker if gpus_per_worker is not None else (1 if use_gpu else 0)
        self.workers = []
        self._server = None
        self._pg = None

    def start(self, executable_cls=None, executable_args=None, executable_kwargs=None, extra_env_vars=None):
        import ray
        from ray.util.placement_group import placement_group

        if self.num_hosts:
            bundles, strat = strategy.colocated_bundles(self.num_hosts, self.num_workers_per_host, self.cpus_per_worker, self.gpus_per_worker)
            total = self.num_hosts * self.num_workers_per_host
        else:
            bundles, strat = strategy.pack_bundles(self.num_workers, self.cpus_per_worker, self.gpus_per_worker)
            total = self.num_workers
        self._pg = placement_group(bundles, strategy=strat)
        ray.get(self._pg.ready())

        @ray.remote(num_cpus=self.cpus_per_worker, num_gpus=self.gpus_per_worker)
        class _Worker(object):
            def __init__(self):
                self.executable = None

            def hostname(self):
                import socket
                return socket.gethostname()

            def update_env_vars(self, env):
                import os
                os.environ.update({k: str(v) for k, v in env.items()})

            def start_executable(self, cls, args, kwargs):
                self.executable = cls(*(args or []), **(kwargs or {}))

            def execute(self, fn):
                return fn(self.executable) if self.executable is not None else fn()

        self.workers = [_Worker.options(placement_group=self._pg).remote() for _ in range(total)]
        hostnames = ray.get([w.hostname.remote() for w in self.workers])
        envs = strategy.assign_ranks(hostnames)
        self._server = RendezvousServer()
        port = self._server.start_server()
        addr = network.get_driver_ip(None)
        for w, env in zip(self.workers, envs):
            env = dict(env, HOROVOD_GLOO_RENDEZVOUS_ADDR=addr, HOROVOD_GLOO_RENDEZVOUS_PORT=str(port), **(extra_env_vars or {}))
            ray.get(w.update_env_vars.remote(env))
        if executable_cls is not None:
            ray.get([w.start_executable.remote(executable_cls, executable_args, executable_kwargs) for w in self.workers])

    def execute(self, fn):
        import ray
        return ray.get([w.execute.remote(fn) for w in self.workers])

    def run(self, fn, args=None, kwargs=None):
        import ray
        args, kwargs = args or [], kwargs or {}
        return ray.get([w.execute.remote(lambda _=None: fn(*args, **kwargs)) for w in self.workers])

    def shutdown(self):
        import ray
        for w in self.workers:
            ray.kill(w)
        self.workers = []
        if self._server:
            self._server.stop()
            self._server = None
