"""Generic task / driver services (runner/common/service/{task_service,driver_service}.py).
Reference coverage model: test/single/test_service.py and test_task_service.py (run a command through the service with an
environment, stream its output, abort it, exit codes; task registration and address bookkeeping on the driver side)."""
import io
import sys
import threading
import time

import pytest

from horovod_b200.runner.common.service import driver_service, task_service
from horovod_b200.runner.common.util import secret
from horovod_b200.runner.common.util.timeout import Timeout, TimeoutException


@pytest.fixture
def task():
    key = secret.make_secret_key()
    svc = task_service.BasicTaskService('test task service', 0, key, command_env={'HVD_TASK_DEFAULT': 'from-service', 'HVD_TASK_DROP': None})
    client = task_service.BasicTaskClient('test task service', svc.addresses(), key, attempts=1)
    yield svc, client
    svc.shutdown()


def test_run_command_streams_output_and_reports_exit_code(task, monkeypatch):
    svc, client = task
    monkeypatch.setenv('HVD_TASK_DROP', 'inherited')
    assert client.command_result() == (False, None) and not svc.check_for_command_start(0.05)
    cmd = ("%s -c \"import os, sys; print('out', os.environ['HVD_TASK_DEFAULT'], os.environ['HVD_TASK_EXTRA'], "
           "os.environ.get('HVD_TASK_DROP')); print('err line', file=sys.stderr); sys.exit(7)\"" % sys.executable)
    client.run_command(cmd, {'HVD_TASK_EXTRA': 'from-request'}, capture_stdout=True, capture_stderr=True)
    client.run_command('echo must not run twice', {})                       # a retried / second request is acknowledged and ignored
    out, err = io.StringIO(), io.StringIO()
    threads = client.stream_command_output(out, err)
    assert client.wait_for_command_exit_code(delay=0.2) == 7
    for t in threads:
        t.join(10)
    assert client.command_result() == (True, 7) and client.command_terminated() and svc.command_exit_code() == 7
    assert out.getvalue() == '[0]<stdout>:out from-service from-request None\n', out.getvalue()
    assert err.getvalue() == '[0]<stderr>:err line\n' and 'twice' not in out.getvalue()
    svc.wait_for_command_termination()


def test_abort_kills_the_command_tree(task):
    svc, client = task
    client.run_command('sleep 600', {}, capture_stdout=True)
    svc.wait_for_command_start(timeout=10)
    t0 = time.time()
    client.abort_command()
    code = client.wait_for_command_exit_code(delay=0.5)
    assert code != 0 and time.time() - t0 < 20
    with pytest.raises(task_service.CommandOutputNotCaptured):
        client._send(task_service.StreamCommandStdErrRequest(0))


def test_registration_signal_and_result_hand_back(task):
    svc, client = task
    with pytest.raises(TimeoutException):
        svc.wait_for_initial_registration(Timeout(0.3, 'Timed out waiting for {activity}.'))
    waiter = threading.Thread(target=svc.wait_for_initial_registration, args=(Timeout(10, 'Timed out waiting for {activity}.'),))
    waiter.start()
    client.notify_initial_registration_complete()
    waiter.join(10)
    assert not waiter.is_alive()
    assert svc.fn_result() is None
    client.register_code_result({'rank': 3, 'loss': 0.25})
    assert svc.fn_result() == {'rank': 3, 'loss': 0.25}
    with pytest.raises(TimeoutError):
        svc.wait_for_command_start(timeout=0.1)


def test_wrong_key_is_rejected(task):
    svc, _ = task
    from horovod_b200.runner.common.util import network
    with pytest.raises(network.NoValidAddressesFound):
        task_service.BasicTaskClient('test task service', svc.addresses(), secret.make_secret_key(), attempts=1)


def test_driver_service_registry_and_waits():
    key = secret.make_secret_key()
    svc = driver_service.BasicDriverService(3, 'test driver service', key)
    try:
        client = driver_service.BasicDriverClient('test driver service', svc.addresses(), key)
        with pytest.raises(TimeoutException):
            svc.wait_for_initial_registration(Timeout(0.3, 'Timed out waiting for {activity}.'))
        addr = {'lo': [('127.0.0.1', 5000)], 'eth0': [('10.1.2.3', 5000)]}
        client.register_task(2, addr, 'host-b')
        client.register_task(0, {'lo': [('127.0.0.1', 5001)]}, 'host-a')
        client.register_task(1, {'lo': [('127.0.0.1', 5002)]}, 'host-a')
        svc.wait_for_initial_registration(Timeout(5, 'Timed out waiting for {activity}.'))
        assert svc.task_indices() == [0, 1, 2] and svc.task_host_hash_indices() == {'host-a': [0, 1], 'host-b': [2]}
        assert svc.task_index_host_hash(2) == 'host-b' and client.all_task_addresses(2) == addr
        assert svc.task_addresses_for_driver(2) == {'lo': [('127.0.0.1', 5000)]}      # the interface the request arrived through
        client.register_task(1, {'lo': [('127.0.0.1', 5003)]}, 'host-b')             # task 1 restarted on the other host
        assert svc.task_host_hash_indices() == {'host-a': [0], 'host-b': [1, 2]} and client.all_task_addresses(1) == {'lo': [('127.0.0.1', 5003)]}
        with pytest.raises(TimeoutException):
            svc.wait_for_task_to_task_address_updates(Timeout(0.3, 'Timed out waiting for {activity}.'))
        for i in range(3):
            client.register_task_to_task_addresses(i, {'lo': [('127.0.0.1', 6000 + i)]})
        svc.wait_for_task_to_task_address_updates(Timeout(5, 'Timed out waiting for {activity}.'))
        assert svc.task_addresses_for_tasks(1) == {'lo': [('127.0.0.1', 6001)]}
    finally:
        svc.shutdown()
