"""Elastic jobs through the unified `RayExecutor(min_workers=..., max_workers=...)` API (reference horovod/ray/elastic_v2.py:
`RayHostDiscovery` :40, `TestDiscovery` :74, `ElasticParams` :151, `ElasticAdapter` :197-560).

The machinery — discovery loop, rendezvous, re-planning after a failure — is `horovod_b200.ray.elastic.ElasticRayExecutor`
on top of `runner.elastic.driver.ElasticDriver`; the adapter maps the Adapter interface onto it."""
import functools
import logging
import random
import time
from dataclasses import dataclass
from typing import List, Optional

from horovod_b200.ray.adapter import Adapter, BaseParams
from horovod_b200.ray.elastic import ElasticRayExecutor, RayHostDiscovery  # noqa: F401

logger = logging.getLogger(__name__)


class TestDiscovery(RayHostDiscovery):
    """Discovery for fault-injection runs: starts from what Ray reports and then adds / removes hosts at random, so a
    training script can be exercised against a changing cluster without touching the cluster (reference :74-148)."""
    __test__ = False          # not a pytest class

    def __init__(self, min_hosts, max_hosts, change_frequency_s, use_gpu=False, cpus_per_worker=1, gpus_per_worker=1,
                 verbose=True, _graceful=True, nodes_fn=None, seed=None):
        super().__init__(use_gpu=use_gpu, cpus_per_slot=cpus_per_worker, gpus_per_slot=gpus_per_worker, nodes_fn=nodes_fn)
        self._min_hosts, self._max_hosts = min_hosts, max_hosts
        self._change_frequency_s = change_frequency_s
        self._graceful, self.verbose = _graceful, verbose
        self._last_reset_t = None
        self._removed_hosts = set()
        self._rng = random.Random(seed)

    def add_host(self, hosts):
        available = self._removed_hosts & set(hosts)
        if available:
            host = self._rng.choice(sorted(available))
            self._removed_hosts.remove(host)
            self._log('adding host %s' % host)
        else:
            self._log('unable to add a host: none is removed')

    def remove_host(self, hosts):
        good = [h for h in hosts if h not in self._removed_hosts]
        if good:
            host = self._rng.choice(sorted(good))
            self._removed_hosts.add(host)
            self._log('removing host %s' % host)

    def change_hosts(self, hosts):
        for h in list(self._removed_hosts):
            if h not in hosts:
                self._removed_hosts.remove(h)
        current = len(hosts) - len(self._removed_hosts)
        if current <= self._min_hosts:
            self.add_host(hosts)
        elif current >= self._max_hosts:
            self.remove_host(hosts)
        elif self._rng.random() < 0.5:
            self.add_host(hosts)
        else:
            self.remove_host(hosts)

    def find_available_hosts_and_slots(self):
        now = time.time()
        if self._last_reset_t is None:
            self._last_reset_t = now
        hosts = super().find_available_hosts_and_slots()
        if now - self._last_reset_t >= self._change_frequency_s:
            self.change_hosts(hosts)
            self._last_reset_t = now
        self._log('total hosts %d, removed %s' % (len(hosts), sorted(self._removed_hosts)))
        return {h: s for h, s in hosts.items() if h not in self._removed_hosts}

    def _log(self, msg):
        if self.verbose:
            logger.info('TestDiscovery: %s', msg)


@dataclass
class ElasticParams(BaseParams):
    """`min_workers` .. `max_workers` workers; the job restarts from the last commit whenever the set changes."""
    min_workers: int = 1
    max_workers: Optional[int] = None
    reset_limit: Optional[int] = None
    cooldown_range: Optional[List[int]] = None
    elastic_timeout: int = 600
    override_discovery: bool = True

    def __post_init__(self):
        super().__post_init__()
        if self.min_workers is None or self.min_workers < 1:
            raise ValueError('`min_workers` must be provided (>= 1) for an elastic job.')
        if self.max_workers is not None and self.max_workers < self.min_workers:
            raise ValueError('`max_workers` (%s) must not be smaller than `min_workers` (%s).' % (self.max_workers, self.min_workers))

    @property
    def elastic(self):
        return True

    @property
    def adapter(self):
        return ElasticAdapter


class ElasticAdapter(Adapter):
    def __init__(self, settings, params, env_vars=None, actor_factory=None, queue_factory=None):
        self.settings, self.params = settings, params
        self.env_vars = dict(env_vars or {})
        self._actor_factory, self._queue_factory = actor_factory, queue_factory
        self.executor = None

    def start(self, executable_cls=None, executable_args=None, executable_kwargs=None, extra_env_vars=None):
        if executable_cls is not None:
            raise ValueError('executable_cls is not supported by the elastic executor: workers come and go between resets.')
        p = self.params
        extra = {'cooldown_range': p.cooldown_range} if p.cooldown_range else {}
        settings = ElasticRayExecutor.create_settings(min_num_proc=p.min_workers, max_num_proc=p.max_workers, reset_limit=p.reset_limit,
                                                      elastic_timeout=p.elastic_timeout, timeout_s=self.settings.timeout_s,
                                                      nics=self.settings.nics, **extra)
        if not p.override_discovery:
            settings.discovery = getattr(self.settings, 'discovery', None)
        env = dict(self.env_vars)
        env.update(extra_env_vars or {})
        self.executor = ElasticRayExecutor(settings, use_gpu=p.use_gpu, cpus_per_slot=p.cpus_per_worker,
                                           gpus_per_slot=p.gpus_per_worker or None, env_vars=env,
                                           override_discovery=p.override_discovery, actor_factory=self._actor_factory,
                                           queue_factory=self._queue_factory)
        self.executor.start()

    def run(self, fn, args=None, kwargs=None, callbacks=None):
        return self.executor.run(functools.partial(fn, *tuple(args or ()), **dict(kwargs or {})), callbacks=callbacks)

    def run_remote(self, fn, args=None, kwargs=None, callbacks=None):
        raise NotImplementedError('Elastic jobs have no fixed worker set to hand futures out for: use run().')

    def execute(self, fn, callbacks=None):
        raise NotImplementedError('Elastic jobs keep no executable on their workers: use run().')

    def execute_single(self, fn):
        raise NotImplementedError('Elastic jobs keep no executable on their workers: use run().')

    def shutdown(self):
        ex, self.executor = self.executor, None
        if ex is not None:
            for part in (ex.driver, ex.rendezvous):
                try:
                    if part is not None:
                        part.stop()
                except Exception:  # noqa: BLE001 - already stopped by run()
                    pass
