"""Placement strategies (pure) and RayBackend against an in-process stand-in for Ray (tests/fakes/ray).
Reference coverage model: test/single/test_ray.py (colocated / pack placement, resource requests, pg timeout, gpu args)."""
import os
import subprocess
import sys

import pytest

from horovod_b200.ray import strategy

HERE = os.path.dirname(os.path.abspath(__file__))


def test_colocated_strategy_layout():
    s = strategy.ColocatedStrategy(num_hosts=2, num_workers_per_host=4, cpus_per_worker=2, gpus_per_worker=1)
    assert s.num_workers == 8 and s.placement == 'STRICT_SPREAD'
    assert s.bundles == [{'CPU': 8, 'GPU': 4}, {'CPU': 8, 'GPU': 4}]
    assert s.worker_bundle == [0, 0, 0, 0, 1, 1, 1, 1] and s.worker_resources == [{'CPU': 2, 'GPU': 1}] * 8
    assert s.total_resources() == {'CPU': 16, 'GPU': 8}
    assert strategy.colocated_bundles(2, 4, 2, 1) == (s.bundles, 'STRICT_SPREAD')
    for bad in ((0, 1), (1, 0)):
        with pytest.raises(ValueError):
            strategy.ColocatedStrategy(*bad)


def test_pack_strategy_layout():
    s = strategy.PackStrategy(3, cpus_per_worker=1)
    assert s.bundles == [{'CPU': 1}] * 3 and s.placement == 'PACK' and s.worker_bundle == [0, 1, 2]
    assert 'GPU' not in s.worker_resources[0] and strategy.PGStrategy is strategy.PackStrategy
    with pytest.raises(ValueError):
        strategy.PackStrategy(0)
    with pytest.raises(ValueError):
        strategy.PackStrategy(2, gpus_per_worker=-1)


def test_executor_argument_validation():
    from horovod_b200.ray import RayExecutor
    with pytest.raises(ValueError):
        RayExecutor()
    with pytest.raises(ValueError):
        RayExecutor(num_workers=2, num_hosts=1)
    with pytest.raises(ValueError):
        RayExecutor(num_workers=2, gpus_per_worker=1)                    # gpus without use_gpu
    with pytest.raises(ValueError):
        RayExecutor(num_workers=2, use_gpu=True, gpus_per_worker=0)
    with pytest.raises(ValueError):
        RayExecutor(num_workers=2, min_workers=1)                        # static and elastic sizes mixed
    with pytest.raises(ValueError):
        RayExecutor(max_workers=4)                                       # elastic needs min_workers
    with pytest.raises(ValueError):
        RayExecutor(min_workers=4, max_workers=2)
    ex = RayExecutor(min_workers=1, max_workers=3)
    assert ex.elastic and ex.num_workers == 1
    with pytest.raises(ValueError, match='executable_cls'):
        ex.start(executable_cls=object)
    ex = RayExecutor(num_hosts=2, num_workers_per_host=2, use_gpu=True)
    bundles, placement, wb, wr = ex._placement()
    assert placement == 'STRICT_SPREAD' and bundles == [{'CPU': 2, 'GPU': 2}] * 2 and wb == [0, 0, 1, 1] and wr[0] == {'CPU': 1, 'GPU': 1}


SCRIPT = r'''
import ray
from horovod_b200.ray.runner import RayBackend
from horovod_b200.ray import strategy

plan = strategy.ColocatedStrategy(2, 2, cpus_per_worker=1)
b = RayBackend(plan.bundles, plan.placement, pg_timeout_s=2, use_current_placement_group=True, worker_bundle=plan.worker_bundle,
               worker_resources=plan.worker_resources)
pg = ray._state['groups'][-1]
assert pg.strategy == 'STRICT_SPREAD' and pg.bundle_specs == [{'CPU': 2}, {'CPU': 2}]
actors = [b.create(i, env={'HVD_FAKE_RAY_TEST': str(i)}) for i in range(4)]
assert [a.options['scheduling_strategy'].placement_group_bundle_index for a in actors] == [0, 0, 1, 1]
assert all(a.options['num_cpus'] == 1 and a.options['num_gpus'] == 0 and a.options['scheduling_strategy'].placement_group is pg for a in actors)
outs = b.get([b.call(a, 'execute', (lambda i=i: i * 10)) for i, a in enumerate(actors)], timeout=10)
assert outs == [0, 10, 20, 30], outs
import os
assert os.environ.get('HVD_FAKE_RAY_TEST') == '3'          # update_env ran on the (in-process) actors
b.kill(actors[0])
assert ray._state['killed'] == [actors[0]]
b.shutdown()
assert pg.removed

# a placement group that is already current is reused and NOT removed on shutdown
ray._state['current_pg'] = ray.util.placement_group.placement_group([{'CPU': 1}], 'PACK')
n = len(ray._state['groups'])
b2 = RayBackend([{'CPU': 1}], 'PACK', use_current_placement_group=True)
assert len(ray._state['groups']) == n and b2.pg is ray._state['current_pg']
b2.shutdown()
assert not ray._state['current_pg'].removed
ray._state['current_pg'] = None

# more than the cluster has: times out with a message that names both sides
try:
    RayBackend([{'CPU': 64}], 'PACK', pg_timeout_s=0.3)
    raise SystemExit('expected TimeoutError')
except TimeoutError as e:
    assert 'Placement group creation timed out' in str(e) and "'CPU': 64" in str(e) and "'CPU': 8" in str(e), str(e)

# discovery through ray.nodes()
from horovod_b200.ray import RayHostDiscovery
assert RayHostDiscovery(cpus_per_slot=2).find_available_hosts_and_slots() == {'10.0.0.1': 4}
assert RayHostDiscovery(use_gpu=True).find_available_hosts_and_slots() == {}
print('RAY BACKEND OK')
'''


def test_ray_backend_against_stand_in(tmp_path):
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(HERE, 'fakes'), os.path.dirname(HERE), env.get('PYTHONPATH', '')])
    script = tmp_path / 'ray_backend_check.py'
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'RAY BACKEND OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_params_and_adapter_interfaces():
    from horovod_b200.ray.adapter import Adapter, BaseParams
    from horovod_b200.ray.elastic_v2 import ElasticAdapter, ElasticParams
    from horovod_b200.ray.runner import StaticAdapter, StaticParams
    p = StaticParams(num_hosts=2, num_workers_per_host=3, use_gpu=True)
    assert not p.elastic and p.adapter is StaticAdapter and p.total_workers == 6 and p.gpus_per_worker == 1
    assert StaticParams(num_workers=4).gpus_per_worker == 0
    with pytest.raises(ValueError):
        StaticParams()
    with pytest.raises(ValueError):
        BaseParams(gpus_per_worker=2)
    e = ElasticParams(min_workers=2, max_workers=4, cooldown_range=[1, 2])
    assert e.elastic and e.adapter is ElasticAdapter
    with pytest.raises(ValueError):
        ElasticParams(min_workers=0)
    assert issubclass(StaticAdapter, Adapter) and issubclass(ElasticAdapter, Adapter)
    with pytest.raises(TypeError):
        Adapter()                                                        # abstract


def test_base_worker_and_utils(monkeypatch):
    from horovod_b200.ray import BaseHorovodWorker, utils
    for k in ('HOROVOD_HOSTNAME', 'HOROVOD_RANK', 'HOROVOD_SIZE', 'HVD_FAKE_VAR'):
        monkeypatch.setenv(k, 'restored-at-teardown')       # the worker writes these into os.environ (it owns its process)
    w = BaseHorovodWorker(world_rank=3, world_size=8)
    assert os.environ['HOROVOD_RANK'] == '3' and os.environ['HOROVOD_SIZE'] == '8'
    assert w.update_env_vars({'HVD_FAKE_VAR': 7}) and w.env_vars()['HVD_FAKE_VAR'] == '7'
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '2,5')
    assert w.get_gpu_ids() == ['2', '5']
    assert w.execute(lambda a, b=1: a + b, (1,), {'b': 2}) == 3

    class Exe:
        def __init__(self, k, scale=1):
            self.v = k * scale
    w.start_executable(Exe, [3], {'scale': 2})
    assert w.execute(lambda exe: exe.v) == 6
    assert utils.nics_to_env_var({'eth1', 'eth0'}) == {'HOROVOD_GLOO_IFACE': 'eth0', 'NCCL_SOCKET_IFNAME': 'eth0,eth1'}

    class S:
        nics = None
    s = S()
    s.nics = ['ib0']
    assert utils.detect_nics(s, ['a', 'b']) == {'ib0'}
    s.nics = None
    assert utils.detect_nics(s, ['a']) and isinstance(utils.detect_nics(s, ['a']), set)      # one host: whatever is local
    tables = {'w1': {'lo': ['127.0.0.1'], 'eth0': ['10.0.0.1'], 'ib0': ['10.1.0.1']}, 'w2': {'lo': ['127.0.0.1'], 'eth0': ['10.0.0.2']}}
    assert utils.detect_nics(s, ['a', 'b'], ['w1', 'w2'], call=lambda w, fn: tables[w]) == {'eth0'}
    tables['w2'] = {'lo': ['127.0.0.1'], 'enp1': ['10.0.0.2']}
    with pytest.raises(RuntimeError, match='common'):
        utils.detect_nics(s, ['a', 'b'], ['w1', 'w2'], call=lambda w, fn: tables[w])


def test_test_discovery_adds_and_removes_hosts(monkeypatch):
    from horovod_b200.ray.elastic_v2 import TestDiscovery
    nodes = [{'alive': True, 'NodeManagerAddress': '10.0.0.%d' % i, 'Resources': {'CPU': 4.0}} for i in range(1, 5)]
    clock = [1000.0]
    monkeypatch.setattr('horovod_b200.ray.elastic_v2.time.time', lambda: clock[0])
    d = TestDiscovery(min_hosts=2, max_hosts=3, change_frequency_s=10, nodes_fn=lambda: nodes, verbose=False, seed=1)
    assert len(d.find_available_hosts_and_slots()) == 4            # nothing changes before the first period is over
    sizes = []
    for _ in range(40):
        clock[0] += 11
        hosts = d.find_available_hosts_and_slots()
        assert all(s == 4 for s in hosts.values())
        sizes.append(len(hosts))
    assert sizes[0] == 3 and min(sizes) >= 2 and max(sizes[1:]) <= 3 and len(set(sizes)) > 1, sizes
    nodes.pop()                                                    # a removed host that really disappeared is forgotten
    clock[0] += 11
    assert set(d.find_available_hosts_and_slots()) <= {'10.0.0.1', '10.0.0.2', '10.0.0.3'}
