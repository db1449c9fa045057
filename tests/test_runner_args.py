"""Launcher argument / config / controller-selection cases (families of the reference's test/single/test_run.py:
test_params_args, test_autotune_args, test_autotuning_with_fixed_param, test_timeline_args, test_stall_check_args,
test_library_args, test_library_env_override, test_logging_args, test_config_file(_override_args),
test_validate_config_args, test_hash, test_host_hash, test_get_mpi_implementation, test_run_controller,
test_mpi_run_*, test_horovodrun_hostfile, test_get_host_assignments_*)."""
import itertools
import os
import sys
from unittest import mock

import pytest

from horovod_b200.runner import launch, mpi_run as mpi_mod
from horovod_b200.runner.common.util import config_parser, hosts, settings as hvd_settings
from horovod_b200.runner.common.util.host_hash import host_hash


def parse(*argv):
    with mock.patch.object(sys, 'argv', ['hvdrun'] + list(argv)):
        return launch.parse_args()


def env_of(*argv, env=None):
    env = {} if env is None else env
    config_parser.set_env_from_args(env, parse(*argv))
    return env


def test_params_args():
    env = env_of('-np', '2', '--fusion-threshold-mb', '10', '--cycle-time-ms', '20', '--cache-capacity', '512',
                 '--hierarchical-allreduce', '--hierarchical-allgather', '--thread-affinity', '0,1', '--num-nccl-streams', '2')
    assert env['HOROVOD_FUSION_THRESHOLD'] == str(10 * 1024 * 1024)
    assert env['HOROVOD_CYCLE_TIME'] == '20.0' and env['HOROVOD_CACHE_CAPACITY'] == '512'
    assert env['HOROVOD_HIERARCHICAL_ALLREDUCE'] == '1' and env['HOROVOD_HIERARCHICAL_ALLGATHER'] == '1'
    assert env['HOROVOD_THREAD_AFFINITY'] == '0,1' and env['HOROVOD_NUM_NCCL_STREAMS'] == '2'
    env = env_of('-np', '2', '--no-hierarchical-allreduce')
    assert env['HOROVOD_HIERARCHICAL_ALLREDUCE'] == '0' and 'HOROVOD_FUSION_THRESHOLD' not in env
    with pytest.raises(SystemExit):
        parse('-np', '2', '--hierarchical-allreduce', '--no-hierarchical-allreduce')


def test_autotune_args():
    env = env_of('-np', '2', '--autotune', '--autotune-log-file', '/tmp/a.csv', '--autotune-warmup-samples', '1',
                 '--autotune-steps-per-sample', '5', '--autotune-bayes-opt-max-samples', '10', '--autotune-gaussian-process-noise', '0.2')
    assert env['HOROVOD_AUTOTUNE'] == '1' and env['HOROVOD_AUTOTUNE_LOG'] == '/tmp/a.csv'
    assert env['HOROVOD_AUTOTUNE_WARMUP_SAMPLES'] == '1' and env['HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE'] == '5'
    assert env['HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES'] == '10' and env['HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE'] == '0.2'
    # tuning knobs without --autotune are not exported
    assert not [k for k in env_of('-np', '2', '--autotune-warmup-samples', '1') if 'AUTOTUNE' in k]
    # a fixed parameter is still exported next to autotune (the runtime pins it and tunes the rest)
    env = env_of('-np', '2', '--autotune', '--cache-capacity', '1024', '--no-hierarchical-allgather')
    assert env['HOROVOD_AUTOTUNE'] == '1' and env['HOROVOD_CACHE_CAPACITY'] == '1024' and env['HOROVOD_HIERARCHICAL_ALLGATHER'] == '0'


def test_timeline_stall_library_logging_args():
    env = env_of('-np', '2', '--timeline-filename', '/tmp/t.json', '--timeline-mark-cycles')
    assert env['HOROVOD_TIMELINE'] == '/tmp/t.json' and env['HOROVOD_TIMELINE_MARK_CYCLES'] == '1'
    assert 'HOROVOD_TIMELINE_MARK_CYCLES' not in env_of('-np', '2', '--timeline-mark-cycles')
    env = env_of('-np', '2', '--no-stall-check')
    assert env['HOROVOD_STALL_CHECK_DISABLE'] == '1'
    env = env_of('-np', '2', '--stall-check-warning-time-seconds', '10', '--stall-check-shutdown-time-seconds', '20')
    assert env['HOROVOD_STALL_CHECK_TIME_SECONDS'] == '10' and env['HOROVOD_STALL_SHUTDOWN_TIME_SECONDS'] == '20'
    assert 'HOROVOD_STALL_CHECK_DISABLE' not in env
    env = env_of('-np', '2', '--mpi-threads-disable', '--gpu-backend', 'nccl', '--allreduce-variant', 'twoshot', '--wire-dtype', 'bf16',
                 '--comm-ctas', '64')
    assert env['HOROVOD_MPI_THREADS_DISABLE'] == '1' and env['HVD_GPU_BACKEND'] == 'nccl'
    assert env['HVD_ALLREDUCE_VARIANT'] == 'twoshot' and env['HVD_WIRE_DTYPE'] == 'bf16' and env['HVD_COMM_CTAS'] == '64'
    env = env_of('-np', '2', '--log-level', 'INFO', '--log-hide-timestamp')
    assert env['HOROVOD_LOG_LEVEL'] == 'INFO' and env['HOROVOD_LOG_HIDE_TIME'] == '1'
    # existing environment entries that no option touches survive
    env = env_of('-np', '2', env={'HOROVOD_GLOO_TIMEOUT_SECONDS': '1800'})
    assert env == {'HOROVOD_GLOO_TIMEOUT_SECONDS': '1800'}
    with pytest.raises(SystemExit):
        parse('-np', '2', '--gpu-backend', 'rocm')


CONFIG = """
controller: gloo
params:
  fusion_threshold_mb: 32
  cycle_time_ms: 10
  cache_capacity: 2048
  hierarchical_allreduce: true
  hierarchical_allgather: true
autotune:
  enabled: true
  log_file: autotune_log.csv
  warmup_samples: 5
  steps_per_sample: 20
  bayes_opt_max_samples: 50
  gaussian_process_noise: 0.9
timeline:
  filename: timeline.json
  mark_cycles: true
stall_check:
  enabled: false
  warning_time_seconds: 120
  shutdown_time_seconds: 240
library_options:
  mpi_threads_disable: true
  gpu_backend: p2p
  wire_dtype: fp16
logging:
  level: INFO
  hide_timestamp: true
"""


def test_config_file_and_cli_override(tmp_path):
    cfg = tmp_path / 'c.yaml'
    cfg.write_text(CONFIG)
    a = parse('-np', '2', '--config-file', str(cfg))
    assert a.use_gloo and a.fusion_threshold_mb == 32 and a.cycle_time_ms == 10 and a.cache_capacity == 2048
    assert a.hierarchical_allreduce and a.hierarchical_allgather and a.autotune and a.autotune_log_file == 'autotune_log.csv'
    assert (a.autotune_warmup_samples, a.autotune_steps_per_sample, a.autotune_bayes_opt_max_samples) == (5, 20, 50)
    assert a.autotune_gaussian_process_noise == 0.9 and a.timeline_filename == 'timeline.json' and a.timeline_mark_cycles
    assert a.no_stall_check and a.stall_check_warning_time_seconds == 120 and a.stall_check_shutdown_time_seconds == 240
    assert a.mpi_threads_disable and a.gpu_backend == 'p2p' and a.wire_dtype == 'fp16' and a.log_level == 'INFO' and a.log_hide_timestamp
    b = parse('-np', '2', '--fusion-threshold-mb', '128', '--config-file', str(cfg), '--cycle-time-ms', '20', '--no-autotune',
              '--stall-check', '--log-level', 'DEBUG')
    assert b.fusion_threshold_mb == 128 and b.cycle_time_ms == 20 and b.cache_capacity == 2048      # CLI wins, file fills the rest
    assert not b.autotune and not b.no_stall_check and b.log_level == 'DEBUG'
    env = {}
    config_parser.set_env_from_args(env, b)
    assert env['HOROVOD_FUSION_THRESHOLD'] == str(128 << 20) and 'HOROVOD_AUTOTUNE' not in env and env['HOROVOD_STALL_CHECK_DISABLE'] == '0'


@pytest.mark.parametrize('flag', ['--fusion-threshold-mb', '--cycle-time-ms', '--cache-capacity', '--autotune-warmup-samples',
                                  '--autotune-steps-per-sample', '--autotune-bayes-opt-max-samples',
                                  '--stall-check-warning-time-seconds', '--stall-check-shutdown-time-seconds'])
def test_validate_config_args_rejects_negative(flag):
    with pytest.raises(ValueError, match='must be >= 0'):
        parse('-np', '2', flag, '-1')


def test_validate_gaussian_noise_range():
    with pytest.raises(ValueError):
        parse('-np', '2', '--autotune', '--autotune-gaussian-process-noise', '1.5')
    assert parse('-np', '2', '--autotune', '--autotune-gaussian-process-noise', '1').autotune_gaussian_process_noise == 1.0


def test_host_hash_is_stable_and_salted():
    h = host_hash()
    assert h == host_hash() and isinstance(h, str) and len(h) > 8
    assert host_hash('salt-a') != host_hash('salt-b') and host_hash('salt-a') != h
    with mock.patch('socket.gethostname', return_value='some-other-host'):
        assert host_hash() != h


@pytest.mark.parametrize('banner,expected', [
    ('mpirun (Open MPI) 4.1.4', mpi_mod._OMPI_IMPL), ('OpenRTE 2.0', mpi_mod._OMPI_IMPL),
    ('mpirun (IBM Spectrum MPI) 10.3', mpi_mod._SMPI_IMPL), ('HYDRA build details: MPICH 3.3', mpi_mod._MPICH_IMPL),
    ('Intel(R) MPI Library for Linux* OS, Version 2019', mpi_mod._IMPI_IMPL), ('something else', mpi_mod._UNKNOWN_IMPL)])
def test_get_mpi_implementation(banner, expected):
    with mock.patch('horovod_b200.runner.mpi_run.tiny_shell_exec.execute', return_value=(banner, 0)):
        assert mpi_mod._get_mpi_implementation() == expected


def test_get_mpi_implementation_missing():
    with mock.patch('horovod_b200.runner.mpi_run.tiny_shell_exec.execute', return_value=None):
        assert mpi_mod._get_mpi_implementation() == mpi_mod._MISSING_IMPL
    with mock.patch('horovod_b200.runner.mpi_run.tiny_shell_exec.execute', return_value=('not found', 127)):
        assert mpi_mod._get_mpi_implementation() == mpi_mod._MISSING_IMPL


def test_run_controller_matrix():
    for use_gloo, use_mpi, use_js, mpi_there, in_lsf in itertools.product([None, False, True], [None, False, True], [None, False, True],
                                                                         [False, True], [False, True]):
        g, m, j = mock.MagicMock(), mock.MagicMock(), mock.MagicMock()
        with mock.patch('horovod_b200.runner.mpi_run.mpi_available', return_value=mpi_there), \
                mock.patch('horovod_b200.runner.launch.lsf.LSFUtils.using_lsf', return_value=in_lsf):
            if use_gloo:
                expect = 'g'
            elif use_mpi:
                expect = 'm' if mpi_there else ValueError
            elif use_js:
                expect = 'j' if in_lsf else ValueError
            else:
                expect = 'j' if (in_lsf and mpi_there) else 'g'
            if expect is ValueError:
                with pytest.raises(ValueError):
                    launch.run_controller(use_gloo, g, use_mpi, m, use_js, j, 2)
                assert not (g.called or m.called or j.called)
                continue
            launch.run_controller(use_gloo, g, use_mpi, m, use_js, j, 2)
            assert (g.call_count, m.call_count, j.call_count) == {'g': (1, 0, 0), 'm': (0, 1, 0), 'j': (0, 0, 1)}[expect]
        assert launch.is_gloo_used(use_gloo, use_mpi, use_js) == bool(use_gloo or (not use_mpi and not use_js))


def _settings(**kw):
    base = dict(num_proc=2, hosts='localhost:2', verbose=0)
    base.update(kw)
    return hvd_settings.Settings(**base)


def test_mpi_command_minimal_and_full():
    flags, binding, impl = ['-mca pml ob1', '-mca btl ^openib'], ['-bind-to none', '-map-by slot'], mpi_mod._OMPI_IMPL
    cmd = mpi_mod.build_mpi_command(_settings(), None, {}, ['cmd'], flags, binding, impl)
    assert cmd == 'mpirun --allow-run-as-root --tag-output -np 2 -H localhost:2 -bind-to none -map-by slot -mca pml ob1 -mca btl ^openib cmd'
    env = {'PATH': '/bin', 'PYTHONPATH': '/p', 'SSH_CONNECTION': 'x', 'MY_SECRET_KEY': 'k', 'BASH_FUNC_f%%': '() {}'}
    s = _settings(num_proc=4, hosts='h1:2,h2:2', ssh_port=1022, ssh_identity_file='/id', extra_mpi_args='>mpi-extra args go here<',
                  binding_args='>binding args go here<', output_filename='>output filename goes here<')
    cmd = mpi_mod.build_mpi_command(s, ['eth0', 'eth1'], env, ['cmd', 'arg1', 'a b'], flags, binding, impl, '1.2.3.4', 4242)
    assert '-np 4 -H h1:2,h2:2 >binding args go here<' in cmd and '-mca plm_rsh_args "-p 1022 -i /id"' in cmd
    assert '-mca btl_tcp_if_include eth0,eth1 -x NCCL_SOCKET_IFNAME=eth0,eth1' in cmd
    assert '--output-filename >output filename goes here<' in cmd and '-x PATH -x PYTHONPATH' in cmd
    assert 'SSH_CONNECTION' not in cmd and 'MY_SECRET_KEY' not in cmd and 'BASH_FUNC' not in cmd
    assert '-x HOROVOD_GLOO_RENDEZVOUS_ADDR=1.2.3.4 -x HOROVOD_GLOO_RENDEZVOUS_PORT=4242' in cmd
    assert cmd.endswith(">mpi-extra args go here< cmd arg1 'a b'")


def test_mpi_command_large_cluster_and_other_impls():
    many = ','.join('host-%d:1' % i for i in range(mpi_mod._LARGE_CLUSTER_THRESHOLD))
    cmd = mpi_mod.build_mpi_command(_settings(num_proc=mpi_mod._LARGE_CLUSTER_THRESHOLD, hosts=many), None, {}, 'cmd', [], [], mpi_mod._OMPI_IMPL)
    assert '-mca plm_rsh_no_tree_spawn true' in cmd and '-mca plm_rsh_num_concurrent %d' % mpi_mod._LARGE_CLUSTER_THRESHOLD in cmd
    cmd = mpi_mod.build_mpi_command(_settings(hosts='a:1,b:1', ssh_port=2222), None, {'PATH': '/bin'}, 'cmd', [], [], mpi_mod._MPICH_IMPL,
                                    '1.2.3.4', 99)
    assert cmd.startswith('mpirun -l -np 2 -hosts a,b') and '-x PATH' not in cmd and '-bootstrap=ssh -bootstrap-exec-args "-p 2222"' in cmd
    assert '-genv HOROVOD_GLOO_RENDEZVOUS_ADDR 1.2.3.4 -genv HOROVOD_GLOO_RENDEZVOUS_PORT 99' in cmd
    cmd = mpi_mod.build_mpi_command(_settings(hosts='a:1,b:1', binding_args='-x'), None, {}, 'cmd', [], [], mpi_mod._IMPI_IMPL)
    assert '-hosts' not in cmd and ' -H ' not in cmd and ' -x ' not in cmd


def test_mpi_run_raises_without_mpi_and_on_nonzero_exit():
    with mock.patch('horovod_b200.runner.mpi_run._get_mpi_implementation_flags', return_value=(None, None, None)):
        with pytest.raises(Exception, match='MPI'):
            mpi_mod.mpi_run(_settings(), None, {}, 'cmd')


def test_hostfile_and_assignments(tmp_path):
    f = tmp_path / 'hosts'
    f.write_text('# cluster\n172.31.32.7 slots=8\n172.31.33.9 slots=8   # second\nnode3:4\nnode4\n\n')
    assert hosts.parse_host_files(str(f)) == '172.31.32.7:8,172.31.33.9:8,node3:4,node4:1'
    a = parse('-np', '2', '--hostfile', str(f))
    assert a.hostfile == str(f)
    with pytest.raises(SystemExit):
        parse('-np', '2', '-H', 'a:1', '--hostfile', str(f))
    slots = hosts.get_host_assignments(hosts.parse_hosts('worker-0:2,worker-1:2'), 4)
    assert [(s.hostname, s.rank, s.local_rank, s.cross_rank, s.size, s.local_size, s.cross_size) for s in slots] == [
        ('worker-0', 0, 0, 0, 4, 2, 2), ('worker-0', 1, 1, 0, 4, 2, 2), ('worker-1', 2, 0, 1, 4, 2, 2), ('worker-1', 3, 1, 1, 4, 2, 2)]
    # elastic: max_np caps the layout, min_np is the requirement
    slots = hosts.get_host_assignments(hosts.parse_hosts('worker-0:2,worker-1:2'), 1, 3)
    assert len(slots) == 3 and slots[2].hostname == 'worker-1' and slots[2].local_size == 1 and slots[1].cross_size == 1
    with pytest.raises(ValueError, match='Requested more processes'):
        hosts.get_host_assignments(hosts.parse_hosts('a:1,b:1'), 3)
    # heterogeneous
    slots = hosts.get_host_assignments(hosts.parse_hosts('w0:1,w1:2,w2:3'), 6)
    assert [s.cross_size for s in slots] == [3, 3, 2, 3, 2, 1] and [s.cross_rank for s in slots] == [0, 1, 0, 2, 1, 0]
    with pytest.raises(ValueError, match='Invalid host input'):
        hosts.parse_hosts_and_slots('host1,host2:2')
    assert hosts.parse_hosts_and_slots('[::1]:2,h-2.x:4') == (['[::1]', 'h-2.x'], {'[::1]': 2, 'h-2.x': 4})
