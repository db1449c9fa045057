"""Compute service (notice board RPC) and the tf.data-service orchestration with stand-in servers.
Reference coverage model: test/single/test_compute_service.py, test/integration/test_tensorflow2_keras... compute parts."""
import os
import threading
import time

import pytest

from horovod_b200.runner.common.service.compute_service import ComputeClient, ComputeService
from horovod_b200.runner.common.util import secret
from horovod_b200.runner.common.util.timeout import TimeoutException


@pytest.fixture
def service():
    key = secret.make_secret_key()
    svc = ComputeService(2, 2, key)
    yield svc, key
    svc.shutdown()


def test_constructor_validation():
    key = secret.make_secret_key()
    for d, w in ((0, 1), (1, 0), (-1, 2)):
        with pytest.raises(ValueError):
            ComputeService(d, w, key)


def test_dispatcher_registration_and_waits(service):
    svc, key = service
    c = ComputeClient(svc.addresses(), key)
    with pytest.raises(TimeoutException):
        c.wait_for_dispatcher_registration(0, 0.3)
    got = {}
    t = threading.Thread(target=lambda: got.setdefault('addr', c.wait_for_dispatcher_registration(1, 10)))
    t.start()
    time.sleep(0.2)
    c.register_dispatcher(1, 'grpc://host:1234')
    t.join(10)
    assert got['addr'] == 'grpc://host:1234'
    c.register_dispatcher(1, 'grpc://host:1234')                       # idempotent (RPCs are retried)
    with pytest.raises(ValueError, match='already registered'):
        c.register_dispatcher(1, 'grpc://other:1')
    for bad in (-1, 2):
        with pytest.raises(IndexError):
            c.register_dispatcher(bad, 'grpc://x:1')
        with pytest.raises(IndexError):
            c.wait_for_dispatcher_registration(bad, 0.1)


def test_worker_registration_counts_and_shutdown(service):
    svc, key = service
    c = ComputeClient(svc.addresses(), key)
    c.register_worker_for_dispatcher(0, 0)
    with pytest.raises(TimeoutException):
        c.wait_for_dispatcher_worker_registration(0, 0.3)               # 1 of 2
    c.register_worker_for_dispatcher(0, 0)                              # same worker again: still 1
    with pytest.raises(TimeoutException):
        c.wait_for_dispatcher_worker_registration(0, 0.2)
    c.register_worker_for_dispatcher(0, 7)
    c.wait_for_dispatcher_worker_registration(0, 5)
    with pytest.raises(IndexError, match='already has'):
        c.register_worker_for_dispatcher(0, 8)
    with pytest.raises(IndexError):
        c.register_worker_for_dispatcher(5, 0)
    done = threading.Event()
    threading.Thread(target=lambda: (c.wait_for_shutdown(), done.set()), daemon=True).start()
    time.sleep(0.2)
    assert not done.is_set()
    ComputeClient(svc.addresses(), key).shutdown()
    assert done.wait(10)


def test_tf_data_service_orchestration_with_stand_in_servers():
    """TfDataServiceConfig, tf_data_service, send_to_data_service, compute_worker_fn: run in a subprocess whose
    `tensorflow` is tests/fakes/tensorflow (TensorFlow is not installed here)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(here, 'fakes'), os.path.dirname(here), env.get('PYTHONPATH', '')])
    r = subprocess.run([sys.executable, os.path.join(here, 'parallel', 'tf_data_service_check.py')], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and 'TF DATA SERVICE OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
