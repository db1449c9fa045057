"""Ray / Spark integrations exercised on the scheduler-independent core with the local subprocess backend (neither Ray
nor PySpark is installed here).  Reference coverage model: test/single/test_ray.py (RayExecutor: rank table, env,
run/execute/execute_single/run_remote, train end-to-end) and test/integration/test_spark.py (`horovod.spark.run`)."""
import os
import time

import pytest

from horovod_b200.runner.cluster_job import ClusterJob, LocalProcessBackend, assign_slots


def test_assign_slots_contiguous_per_node_and_heterogeneous():
    slots = assign_slots(['a', 'b', 'a', 'b', 'a'])
    assert [s.rank for s in slots] == [0, 3, 1, 4, 2]
    assert [s.local_rank for s in slots] == [0, 0, 1, 1, 2]
    assert [s.local_size for s in slots] == [3, 2, 3, 2, 3]
    assert [s.cross_rank for s in slots] == [0, 1, 0, 1, 0]
    assert [s.cross_size for s in slots] == [2, 2, 2, 2, 1]
    assert all(s.size == 5 for s in slots)


def _train(scale):
    import torch
    import horovod_b200.torch as hvd
    hvd.init()
    out = hvd.allreduce(torch.ones(3) * (hvd.rank() + 1) * scale, op=hvd.Sum).tolist()
    res = (hvd.rank(), hvd.size(), hvd.local_rank(), hvd.local_size(), hvd.cross_rank(), out)
    hvd.shutdown()
    return res


def test_cluster_job_runs_hvd_on_fake_two_node_layout(native_built):
    backend = LocalProcessBackend(node_ids={0: 'n0', 1: 'n1', 2: 'n0', 3: 'n1'})
    job = ClusterJob(backend, 4, env={'HOROVOD_LOG_LEVEL': 'warning', 'OMP_NUM_THREADS': '1'}).start()
    try:
        res = job.run(_train, args=(2.0,), timeout=120)
    finally:
        job.shutdown()
    assert [r[0] for r in res] == [0, 1, 2, 3] and all(r[1] == 4 for r in res)
    assert [r[2] for r in res] == [0, 1, 0, 1] and [r[4] for r in res] == [0, 0, 1, 1]
    assert all(r[3] == 2 for r in res)
    assert all(r[5] == [20.0] * 3 for r in res)


def test_ray_executor_api_over_local_backend(native_built):
    from horovod_b200.ray import RayExecutor

    class Trainer:
        def __init__(self, base):
            self.base = base

        def rank_plus(self):
            return self.base + int(os.environ['HOROVOD_RANK'])

    settings = RayExecutor.create_settings(timeout_s=30)
    with pytest.raises(ValueError):
        RayExecutor(settings)
    with pytest.raises(ValueError):
        RayExecutor(settings, num_workers=2, num_hosts=1)
    with pytest.raises(ValueError):
        RayExecutor(settings, num_workers=2, gpus_per_worker=1)
    ex = RayExecutor(settings, num_workers=2, backend=LocalProcessBackend(), env_vars={'HVD_TEST_FLAG': 'x', 'OMP_NUM_THREADS': '1'})
    ex.start(executable_cls=Trainer, executable_args=[100])
    try:
        assert ex.execute(lambda t: t.rank_plus()) == [100, 101]
        assert ex.execute_single(lambda t: t.base) == 100
        assert ex.run(lambda: os.environ['HVD_TEST_FLAG']) == ['x', 'x']
        res = ex.run(_train, args=[1.0])
        assert [r[0] for r in res] == [0, 1] and all(r[5] == [3.0] * 3 for r in res)
        futs = ex.run_remote(lambda a, b=0: a + b, args=[1], kwargs={'b': 2})
        assert ex.job.backend.get(futs, 30) == [3, 3]
        with pytest.raises(RuntimeError, match='ZeroDivisionError'):
            ex.run(lambda: 1 / 0)
        assert ex.run(lambda: 'still alive') == ['still alive'] * 2
    finally:
        ex.shutdown()


def test_ray_host_discovery_from_node_table():
    from horovod_b200.ray import RayHostDiscovery
    nodes = [{'alive': True, 'NodeManagerAddress': '10.0.0.1', 'Resources': {'CPU': 8.0, 'GPU': 4.0}},
             {'alive': False, 'NodeManagerAddress': '10.0.0.2', 'Resources': {'CPU': 8.0, 'GPU': 4.0}},
             {'alive': True, 'NodeManagerAddress': '10.0.0.3', 'Resources': {'CPU': 3.0}}]
    assert RayHostDiscovery(use_gpu=True, gpus_per_slot=2, nodes_fn=lambda: nodes).find_available_hosts_and_slots() == {'10.0.0.1': 2}
    assert RayHostDiscovery(cpus_per_slot=2, nodes_fn=lambda: nodes).find_available_hosts_and_slots() == {'10.0.0.1': 4, '10.0.0.3': 1}


def _mp_launch(n, task_main):
    """Stands in for a Spark barrier stage: n subprocesses that each run the task body."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=task_main, args=(i, {'OMP_NUM_THREADS': '1', 'HOROVOD_LOG_LEVEL': 'warning'}), daemon=True) for i in range(n)]
    for p in procs:
        p.start()
    return procs


def test_spark_run_over_connect_back_tasks(native_built):
    import horovod_b200.spark as hvd_spark
    res = hvd_spark.run(_train, args=(1.0,), num_proc=3, _launch=_mp_launch, start_timeout=120, verbose=0)
    assert [r[0] for r in res] == [0, 1, 2] and all(r[1] == 3 for r in res)
    assert all(r[5] == [6.0] * 3 for r in res)
    with pytest.raises(ValueError):
        hvd_spark.run(_train, use_mpi=True, num_proc=1, _launch=_mp_launch)


def test_spark_run_without_pyspark_raises():
    import importlib.util
    if importlib.util.find_spec('pyspark') is not None:
        pytest.skip('PySpark is installed')
    import horovod_b200.spark as hvd_spark
    with pytest.raises(ImportError, match='PySpark'):
        hvd_spark.run(lambda: 0)


def _elastic_worker():
    import torch
    import horovod_b200.torch as hvd
    hvd.init()
    out = hvd.allreduce(torch.ones(2) * (hvd.rank() + 1), op=hvd.Sum).tolist()
    res = (hvd.rank(), hvd.size(), out)
    hvd.shutdown()
    return res


def test_elastic_ray_executor_with_local_actors(native_built):
    """ElasticRayExecutor driven by a fixed discovery and subprocess 'actors' (Ray itself is not installed): the elastic
    driver plans 2 slots on localhost, spawns the workers through the actor factory and collects their results."""
    from horovod_b200.ray import ElasticRayExecutor
    from horovod_b200.runner.elastic.discovery import FixedHosts

    backend = LocalProcessBackend()
    created = []

    def actor_factory(hostname, env):
        h = backend.create(len(created), dict(env, OMP_NUM_THREADS='1', HOROVOD_LOG_LEVEL='warning'))
        created.append(h)

        class Actor:
            def execute(self, fn):
                return backend.get([backend.call(h, 'execute', fn)], 120)[0]

            def kill(self):
                backend.kill(h)
        return Actor()

    settings = ElasticRayExecutor.create_settings(min_num_proc=2, max_num_proc=2, elastic_timeout=60, timeout_s=30)
    settings.discovery = FixedHosts({'localhost': 2})
    ex = ElasticRayExecutor(settings, override_discovery=False, actor_factory=actor_factory)
    ex.start()
    try:
        res = ex.run(_elastic_worker)
    finally:
        for h in created:
            backend.kill(h)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] == 2 and r[2] == [3.0, 3.0] for r in res), res


def test_unified_ray_executor_elastic_mode(native_built):
    """RayExecutor(min_workers=..., max_workers=...) (the reference's v2 API) delegates to the elastic executor."""
    from horovod_b200.ray import RayExecutor
    from horovod_b200.runner.elastic.discovery import FixedHosts

    backend = LocalProcessBackend()
    created = []

    def actor_factory(hostname, env):
        h = backend.create(len(created), dict(env, OMP_NUM_THREADS='1', HOROVOD_LOG_LEVEL='warning'))
        created.append(h)

        class Actor:
            def execute(self, fn):
                return backend.get([backend.call(h, 'execute', fn)], 120)[0]

            def kill(self):
                backend.kill(h)
        return Actor()

    settings = RayExecutor.create_settings(timeout_s=30)
    settings.discovery = FixedHosts({'localhost': 2})
    ex = RayExecutor(settings, min_workers=2, max_workers=2, elastic_timeout=60, override_discovery=False,
                     elastic_actor_factory=actor_factory, cpus_per_worker=1)
    ex.start()
    try:
        res = ex.run(_elastic_worker)
    finally:
        for h in created:
            backend.kill(h)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] == 2 and r[2] == [3.0, 3.0] for r in res), res


def _mp_launch_on_nodes(nodes, respawn_marker=None):
    """Launcher that stands in for a non-barrier Spark stage: task i runs on fake node nodes[i]; with `respawn_marker`, a task
    that exits before the marker file exists is started again once (Spark's task retry)."""
    def launch(n, task_main):
        import multiprocessing as mp
        import threading
        ctx = mp.get_context('spawn')

        def start(i):
            env = {'OMP_NUM_THREADS': '1', 'HOROVOD_LOG_LEVEL': 'warning', 'HVD_NODE_ID_OVERRIDE': nodes[i]}
            p = ctx.Process(target=task_main, args=(i, env), daemon=True)
            p.start()
            return p
        procs = [start(i) for i in range(n)]
        if respawn_marker is not None:
            def watch():
                retried = set()
                while not os.path.exists(respawn_marker):
                    for i, p in enumerate(procs):
                        if not p.is_alive() and p.exitcode not in (0, None) and i not in retried:
                            retried.add(i)
                            procs.append(start(i))
                    time.sleep(0.2)
            threading.Thread(target=watch, daemon=True).start()
        return procs
    return launch


def _elastic_train(marker_dir, batches=12, die_on_node=None, die_at=4, step_sleep=0.0):
    """An ordinary elastic training function: commits every 2 batches; the task on `die_on_node` kills its process at batch
    `die_at` of its first life."""
    import torch
    import horovod_b200.torch as hvd
    hvd.init()
    w = torch.zeros(1)
    state = hvd.elastic.TorchState(batch=0, total=0.0, sizes=[])
    node = os.environ.get('HVD_NODE_ID_OVERRIDE')

    @hvd.elastic.run
    def train(state):
        while state.batch < batches:
            if node == die_on_node and state.batch == die_at and not os.path.exists(os.path.join(marker_dir, 'died')):
                open(os.path.join(marker_dir, 'died'), 'w').close()
                os._exit(17)
            s = hvd.allreduce(torch.ones(1), op=hvd.Sum, name='elastic.sum').item()
            time.sleep(step_sleep)
            state.total += s
            state.sizes.append(int(s))
            state.batch += 1
            if state.batch % 2 == 0:
                state.commit()
        return state.batch, state.total, list(state.sizes)
    out = train(state)
    res = (hvd.rank(), hvd.size(), node) + tuple(out)
    hvd.shutdown()
    return res


def test_spark_run_elastic_survives_a_dying_task(native_built, tmp_path):
    """run_elastic over a pool of connect-back tasks (what Spark tasks are): 3 tasks on 3 'hosts', the one on host c kills its
    process mid-training; its host is dropped, the two survivors roll back to their last commit and finish as a 2-rank job."""
    import horovod_b200.spark as hvd_spark
    res = hvd_spark.run_elastic(_elastic_train, args=(str(tmp_path),), kwargs={'die_on_node': 'c'}, num_proc=3, min_num_proc=2,
                                _launch=_mp_launch_on_nodes(['a', 'b', 'c']), start_timeout=120, elastic_timeout=60, verbose=0)
    assert len(res) == 2 and sorted(r[0] for r in res) == [0, 1] and all(r[1] == 2 for r in res), res
    assert sorted(r[2] for r in res) == ['a', 'b'] and os.path.exists(tmp_path / 'died')
    for _, _, _, batch, total, sizes in res:
        assert batch == 12 and sizes[:4] == [3, 3, 3, 3] and sizes[-1] == 2 and set(sizes) == {2, 3}, sizes
        assert total == float(sum(sizes))
    assert res[0][3:] == res[1][3:]                                      # both survivors replayed the same history


def test_spark_run_elastic_picks_up_the_retried_task(native_built, tmp_path):
    """The attempt the scheduler starts for the failed task dials back as a fresh slot on another host and joins at the next
    reset: the job ends with 3 ranks again."""
    import horovod_b200.spark as hvd_spark
    marker = str(tmp_path / 'done')
    try:
        res = hvd_spark.run_elastic(_elastic_train, args=(str(tmp_path),), kwargs={'die_on_node': 'c', 'batches': 400, 'die_at': 4, 'step_sleep': 0.03},
                                    num_proc=3, min_num_proc=2, max_num_proc=3, cooldown_range=[1, 2],
                                    _launch=_mp_launch_on_nodes(['a', 'b', 'c'], respawn_marker=marker), start_timeout=120,
                                    elastic_timeout=60, verbose=0)
    finally:
        open(marker, 'w').close()
    assert len(res) == 3 and sorted(r[0] for r in res) == [0, 1, 2] and all(r[1] == 3 for r in res), res
    sizes = res[0][5]
    assert sizes[:4] == [3, 3, 3, 3] and 2 in sizes and sizes[-1] == 3, (sizes[:10], sizes[-5:])


def test_spark_run_elastic_argument_checks():
    import horovod_b200.spark as hvd_spark
    with pytest.raises(ValueError, match='num_proc is required'):
        hvd_spark.run_elastic(lambda: 0, _launch=lambda n, f: [])
    with pytest.raises(ValueError, match='min_num_proc <= num_proc <= max_num_proc'):
        hvd_spark.run_elastic(lambda: 0, num_proc=2, min_num_proc=3, _launch=lambda n, f: [])


def _logging_worker(scale):
    from horovod_b200.ray import ray_logger
    rank = int(os.environ['HOROVOD_RANK'])
    for step in range(3):
        assert ray_logger.log({'rank': rank, 'step': step, 'loss': scale * step})
    return rank


def test_ray_executor_callbacks_receive_worker_logs(native_built):
    """run(fn, callbacks=[...]): dicts passed to ray_logger.log inside the workers reach the driver's callbacks (reference
    test_ray.py::test_horovod_train with callbacks); without callbacks `log` is a no-op that returns False."""
    from horovod_b200.ray import RayExecutor, ray_logger
    assert ray_logger.log({'x': 1}) is False
    ex = RayExecutor(RayExecutor.create_settings(timeout_s=30), num_workers=2, backend=LocalProcessBackend(), env_vars={'OMP_NUM_THREADS': '1'})
    ex.start()
    seen = []
    try:
        assert ex.run(_logging_worker, args=[0.5], callbacks=[seen.append]) == [0, 1]
        assert ex.run(lambda: __import__('horovod_b200.ray.ray_logger', fromlist=['x']).log({'late': 1})) == [False, False]
    finally:
        ex.shutdown()
    assert sorted((d['rank'], d['step']) for d in seen) == [(r, s) for r in (0, 1) for s in range(3)]
    assert all(d['loss'] == 0.5 * d['step'] for d in seen)


def test_unified_elastic_executor_callbacks_and_unsupported_calls(native_built):
    """Elastic flavour of the unified executor: callbacks travel through the queue factory; the fixed-worker-set calls
    (execute / run_remote) say why they do not exist."""
    import multiprocessing as mp
    from horovod_b200.ray import RayExecutor
    from horovod_b200.ray.elastic_v2 import ElasticAdapter, ElasticParams
    from horovod_b200.runner.elastic.discovery import FixedHosts

    backend = LocalProcessBackend()
    manager = mp.get_context('spawn').Manager()
    created = []

    def actor_factory(hostname, env):
        h = backend.create(len(created), dict(env, OMP_NUM_THREADS='1', HOROVOD_LOG_LEVEL='warning'))
        created.append(h)

        class Actor:
            def execute(self, fn):
                return backend.get([backend.call(h, 'execute', fn)], 120)[0]

            def kill(self):
                backend.kill(h)
        return Actor()

    settings = RayExecutor.create_settings(timeout_s=30)
    settings.discovery = FixedHosts({'localhost': 2})
    ex = RayExecutor(settings, min_workers=2, max_workers=2, override_discovery=False, elastic_actor_factory=actor_factory,
                     elastic_queue_factory=manager.Queue, env_vars={'HVD_TEST_FLAG': 'y'})
    assert isinstance(ex.params, ElasticParams) and isinstance(ex.adapter, ElasticAdapter) and ex.elastic
    ex.start()
    seen = []
    try:
        with pytest.raises(NotImplementedError):
            ex.execute(lambda t: t)
        with pytest.raises(NotImplementedError):
            ex.run_remote(lambda: 1)
        res = ex.run(_logging_worker, args=[2.0], callbacks=[seen.append])
    finally:
        for h in created:
            backend.kill(h)
        ex.shutdown()
        manager.shutdown()
    assert sorted(res) == [0, 1]
    assert sorted((d['rank'], d['step']) for d in seen) == [(r, s) for r in (0, 1) for s in range(3)]


def test_spark_conf_for_elastic_jobs():
    from horovod_b200.spark import conf
    c = conf.elastic_conf()
    assert c['spark.task.maxFailures'] == conf.SPARK_CONF_MAX_INT and c['spark.blacklist.enabled'] == 'false'
    strict = conf.elastic_conf(reuse_failed_executors=False)
    assert strict['spark.blacklist.enabled'] == 'true' and strict['spark.blacklist.stage.maxFailedTasksPerExecutor'] == '1'
    assert set(k for k, _ in (conf.SPARK_CONF_REUSE_FAILING_NODE, conf.SPARK_CONF_REUSE_NODE_ONCE_FOR_SAME_TASK)) <= set(conf.SPARK_CONF_DEFAULT_VALUES)
    warned = []
    assert conf.check_elastic_conf(lambda k, d: d, warn=warned.append) == {'spark.task.maxFailures': '4'} and 'maxFailures' in warned[0]
    assert conf.check_elastic_conf(lambda k, d: c.get(k, d)) == {}
