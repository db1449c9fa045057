"""`hvdrun` end to end on localhost (reference coverage model: test/integration/test_static_run.py): output capture
files, rank prefixes, timestamps, exit-code propagation, slot validation, env forwarding, run-func API."""
import os
import subprocess
import sys

import pytest

from conftest import REPO


def hvdrun(*args, timeout=300, env=None):
    e = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', ''), HOROVOD_LOG_LEVEL='warning')
    e.update(env or {})
    p = subprocess.run([sys.executable, '-m', 'horovod_b200.runner.launch', *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=e, timeout=timeout, cwd=REPO)
    return p.returncode, p.stdout.decode(errors='replace')


PRINT_RANK = "import os; print('hello from', os.environ['HOROVOD_RANK'], 'of', os.environ['HOROVOD_SIZE'], os.environ.get('HVD_T_FORWARD', '-'))"


def test_rank_prefixes_and_env_forwarding(native_built):
    rc, out = hvdrun('-np', '2', sys.executable, '-c', PRINT_RANK, env={'HVD_T_FORWARD': 'forwarded'})
    assert rc == 0, out
    assert '[0]<stdout>:hello from 0 of 2 forwarded' in out and '[1]<stdout>:hello from 1 of 2 forwarded' in out, out


def test_output_filename_writes_per_rank_files(native_built, tmp_path):
    d = tmp_path / 'logs'
    code = "import os, sys; print('to stdout', os.environ['HOROVOD_RANK']); print('to stderr', file=sys.stderr)"
    rc, out = hvdrun('-np', '2', '--output-filename', str(d), sys.executable, '-c', code)
    assert rc == 0, out
    for r in (0, 1):
        so = (d / f'rank.{r}' / 'stdout').read_text()
        se = (d / f'rank.{r}' / 'stderr').read_text()
        assert f'to stdout {r}' in so and 'to stderr' in se, (so, se)


def test_timestamp_prefix(native_built):
    import time
    rc, out = hvdrun('-np', '1', '--prefix-output-with-timestamp', sys.executable, '-c', "print('stamped')")
    assert rc == 0 and 'stamped' in out
    line = [l for l in out.splitlines() if 'stamped' in l][0]
    assert time.strftime('%Y') in line and '[0]<stdout>:stamped' in line, line


def test_nonzero_exit_code_is_reported(native_built):
    code = "import os, sys, time; r = int(os.environ['HOROVOD_RANK']); time.sleep(0.5 if r else 30) if r == 0 else None; sys.exit(3 if r == 1 else 0)"
    rc, out = hvdrun('-np', '2', sys.executable, '-c', code, timeout=300)
    assert rc != 0
    assert 'exited with non-zero status' in out and 'Exit code: 3' in out and 'Process name: 1' in out, out


def test_more_processes_than_slots_is_rejected(native_built):
    rc, out = hvdrun('-np', '3', '-H', 'localhost:2', sys.executable, '-c', 'print(1)')
    assert rc != 0 and 'Requested more processes (3) than there are available slots (2)' in out, out
    rc, out = hvdrun('-np', '2', '-H', 'localhost', sys.executable, '-c', 'print(1)')
    assert rc != 0 and 'Invalid host input' in out, out


def test_version_and_missing_command(native_built):
    rc, out = hvdrun('--version')
    assert rc == 0 and out.strip()
    rc, out = hvdrun('-np', '1')
    assert rc != 0 and 'command' in out.lower(), out


def _fn(a, b=0):
    import horovod_b200.torch as hvd
    hvd.init()
    r = hvd.rank()
    hvd.shutdown()
    return (r, a + b)


def _elastic_fn():
    import torch
    import horovod_b200.torch as hvd
    hvd.init()
    out = hvd.allreduce(torch.ones(1), op=hvd.Sum, name='e').item()
    res = (hvd.rank(), hvd.size(), out)
    hvd.shutdown()
    return res


def test_run_func_api_returns_rank_ordered_results(native_built):
    import horovod_b200
    res = horovod_b200.run(_fn, args=(10,), kwargs={'b': 5}, num_proc=2)
    assert res == [(0, 15), (1, 15)]
    with pytest.warns(DeprecationWarning, match='np is deprecated'):          # the old spelling still works
        assert horovod_b200.run(_fn, args=(1,), np=2, network_interfaces=['lo']) == [(0, 1), (1, 1)]
    with pytest.raises(ValueError, match='deprecated'):
        horovod_b200.run(_fn, args=(1,), num_proc=2, np=3)
    with pytest.raises(ValueError, match='network_interface'):
        horovod_b200.run(_fn, args=(1,), num_proc=1, network_interface='lo', network_interfaces='lo')
    with pytest.raises(Exception):
        horovod_b200.run(lambda: 1 / 0, num_proc=2)


def test_run_func_api_elastic_with_discovery_script(native_built, tmp_path):
    """run(..., host_discovery_script=...) is the programmatic form of `hvdrun --host-discovery-script`."""
    import horovod_b200
    script = tmp_path / 'discover.sh'
    script.write_text('#!/bin/sh\necho localhost:2\n')
    script.chmod(0o755)
    res = horovod_b200.run(_elastic_fn, num_proc=2, min_num_proc=2, max_num_proc=2, host_discovery_script=str(script))
    assert sorted(res) == [(0, 2, 2.0), (1, 2, 2.0)]


def test_hard_crash_of_one_worker_terminates_the_job(native_built):
    """kill -9 of one rank while the other sits in a collective: the launcher tears the job down and reports the signal."""
    import time
    code = ("import os, signal, time, torch\n"
            "import horovod_b200.torch as hvd\n"
            "hvd.init()\n"
            "hvd.allreduce(torch.ones(2))\n"
            "if hvd.rank() == 1:\n"
            "    os.kill(os.getpid(), signal.SIGKILL)\n"
            "hvd.allreduce(torch.ones(2), name='never')\n"
            "time.sleep(60)\n")
    t0 = time.time()
    rc, out = hvdrun('-np', '2', sys.executable, '-c', code, timeout=90)
    assert rc != 0 and time.time() - t0 < 60, out
    assert 'Process name: 1' in out and ('Exit code: 137' in out or 'Exit code: -9' in out), out


def test_workers_do_not_outlive_a_killed_launcher(native_built, tmp_path):
    """SIGKILL the launcher (it cannot run any cleanup): the per-worker supervisor processes notice the parent is gone and
    take the workers down — nothing keeps spinning in a collective forever."""
    import signal
    import time
    script = tmp_path / 'sleeper.py'
    script.write_text(
        "import horovod_b200.torch as hvd, torch, time, os\n"
        "hvd.init()\n"
        "open(%r + '/pid.%%d' %% hvd.rank(), 'w').write(str(os.getpid()))\n"
        "for i in range(3000):\n"
        "    hvd.allreduce(torch.ones(4), name='s')\n"
        "    time.sleep(0.05)\n" % str(tmp_path))
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', ''), HOROVOD_LOG_LEVEL='warning')
    launcher = subprocess.Popen([sys.executable, '-m', 'horovod_b200.runner.launch', '-np', '2', sys.executable, str(script)],
                                env=env, cwd=REPO, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        deadline = time.time() + 60
        while time.time() < deadline and not all((tmp_path / ('pid.%d' % r)).exists() for r in range(2)):
            time.sleep(0.2)
        pids = [int((tmp_path / ('pid.%d' % r)).read_text()) for r in range(2)]
        launcher.send_signal(signal.SIGKILL)
        launcher.wait(10)

        def alive(pid):
            try:
                os.kill(pid, 0)
            except ProcessLookupError:
                return False
            with open('/proc/%d/stat' % pid) as f:            # a zombie waiting to be reaped by init is dead for our purposes
                return f.read().rsplit(')', 1)[1].split()[0] not in ('Z', 'X')
        deadline = time.time() + 30
        while time.time() < deadline and any(alive(p) for p in pids):
            time.sleep(0.5)
        assert not any(alive(p) for p in pids), 'workers %s survived their launcher' % [p for p in pids if alive(p)]
    finally:
        if launcher.poll() is None:
            launcher.kill()


def test_reference_program_runs_under_the_horovod_namespace(native_built):
    """`import horovod.torch as hvd` + `horovodrun -np 2`: the drop-in namespace resolves to this framework's modules."""
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', ''), HOROVOD_LOG_LEVEL='warning')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bin', 'horovodrun'), '-np', '2', sys.executable,
                        os.path.join(REPO, 'tests', 'parallel', 'compat_namespace_worker.py')], env=env, cwd=REPO,
                       capture_output=True, text=True, timeout=200)
    assert r.returncode == 0 and 'COMPAT OK' in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
