import os, threading, time, pickle
import pytest
from horovod_b200.runner.common.service.compute_service import ComputeClient, ComputeService
from horovod_b200.runner.common.util import secret
from horovod_b200.tensorflow.data.compute_service import TfDataServiceConfig, compute_worker_fn, send_to_data_service, tf_data_service


def test_config_round_trip(tmp_path):
    key = secret.make_secret_key()
    cfg = TfDataServiceConfig(2, 3, 'compute', {'lo': [('127.0.0.1', 1234)]}, key, timeout=17)
    path = str(tmp_path / 'compute.json')
    threading.Timer(0.3, cfg.write, args=(path,)).start()
    back = TfDataServiceConfig.read(path, wait_for_file_creation=True, poll_seconds=0.05, timeout=10)
    assert back == cfg and back.key == key and back.addresses == {'lo': [('127.0.0.1', 1234)]} and back.timeout == 17
    assert [f for f in os.listdir(tmp_path)] == ['compute.json']         # the temporary file was renamed, not left behind
    with pytest.raises(AttributeError):
        cfg.timeout = 3
    with pytest.raises(ValueError):
        TfDataServiceConfig(1, 1, 'elsewhere', {}, key)
    with pytest.raises(TimeoutError):
        TfDataServiceConfig.read(str(tmp_path / 'never.json'), wait_for_file_creation=True, poll_seconds=0.05, timeout=0.2)
    import pickle
    assert pickle.loads(pickle.dumps(cfg)) == cfg


class FakeServers:
    """Stand-in for tf.data.experimental.service."""
    log = []

    class DispatchServer:
        _n = 0

        def __init__(self):
            type(self)._n += 1
            self.target = 'grpc://localhost:%d' % (5000 + type(self)._n)
            self.stopped = False

        def _stop(self):
            self.stopped = True

        def join(self):
            pass

    class WorkerConfig:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class WorkerServer:
        def __init__(self, config):
            self.config, self.started, self.stopped = config, False, False

        def start(self):
            self.started = True

        def _stop(self):
            self.stopped = True

        def join(self):
            pass

    @staticmethod
    def distribute(**kw):
        return ('distribute', kw)


class FakeDataset:
    def apply(self, transformation):
        return ('applied', transformation)


def test_compute_side_dispatchers_end_to_end():
    key = secret.make_secret_key()
    svc = ComputeService(2, 2, key)
    try:
        cfg = TfDataServiceConfig(2, 2, 'compute', svc.addresses(), key, timeout=10)
        results = {}

        def worker(rank):
            results[rank] = compute_worker_fn(cfg, rank=rank, servers=FakeServers)
        threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(4)]
        for t in threads:
            t.start()
        # training rank 1 -> dispatcher 1, shared job with deterministic round robin
        out = send_to_data_service(FakeDataset(), cfg, rank=1, size=2, reuse_dataset=True, round_robin=True, servers=FakeServers)
        kind, (name, kw) = out
        assert kind == 'applied' and name == 'distribute' and kw['job_name'] == 'job' and kw['consumer_index'] == 1 and kw['num_consumers'] == 2
        assert kw['service'].startswith('grpc://localhost:') and kw['processing_mode'] == 'distributed_epoch'
        with tf_data_service(cfg, 0, servers=FakeServers) as addr0:
            assert addr0 != kw['service']                                 # rank 0 talks to dispatcher 0
        cfg.compute_client().shutdown()
        for t in threads:
            t.join(20)
        assert sorted(results) == [0, 1, 2, 3]
        assert [results[r][1] is not None for r in range(4)] == [True, False, True, False]   # first worker of a group hosts the dispatcher
        assert all(w.started and w.stopped for w, _ in results.values())
        assert results[1][0].config.dispatcher_address == results[0][1].target.split('://')[1]
    finally:
        svc.shutdown()


def test_training_side_dispatcher_and_restrictions():
    key = secret.make_secret_key()
    svc = ComputeService(1, 1, key)
    try:
        cfg = TfDataServiceConfig(1, 1, 'training', svc.addresses(), key, timeout=10)
        with pytest.raises(RuntimeError, match='training side dispatcher'):
            send_to_data_service(FakeDataset(), cfg, rank=0, servers=FakeServers)
        t = threading.Thread(target=compute_worker_fn, args=(cfg,), kwargs=dict(rank=0, servers=FakeServers), daemon=True)
        t.start()
        with tf_data_service(cfg, 0, servers=FakeServers) as addr:          # rank 0 starts THE dispatcher, the worker attaches
            assert addr.startswith('grpc://')
        cfg.compute_client().shutdown()
        t.join(20)
        assert not t.is_alive()
    finally:
        svc.shutdown()


def test_spark_compute_worker_driver(tmp):
    """spark/tensorflow/compute_worker.main with an injected `run`: the driver hosts the ComputeService, writes the config and
    starts one worker per task; the plan rejects task counts that do not divide by the dispatchers."""
    from horovod_b200.spark.tensorflow import compute_worker as scw
    assert scw.plan(4, 2) == 2
    with pytest.raises(ValueError, match='multiple'):
        scw.plan(3, 2)
    cfg_file = str(tmp / 'compute.json')
    seen = {}

    def fake_run(fn, args=(), num_proc=None, verbose=None, **kw):
        cfg = args[0]
        assert fn is compute_worker_fn and num_proc == 2 and cfg.dispatchers == 1 and cfg.workers_per_dispatcher == 2
        on_disk = TfDataServiceConfig.read(cfg_file)
        assert on_disk.addresses == cfg.addresses and on_disk.key == cfg.key and on_disk.dispatcher_side == 'compute'
        threads = [threading.Thread(target=fn, args=(cfg,), kwargs=dict(rank=r, servers=FakeServers), daemon=True) for r in range(num_proc)]
        for t in threads:
            t.start()
        addr = cfg.compute_client().wait_for_dispatcher_registration(0, 10)
        cfg.compute_client().wait_for_dispatcher_worker_registration(0, 10)
        seen['addr'] = addr
        cfg.compute_client().shutdown()
        for t in threads:
            t.join(20)
            assert not t.is_alive()
        return [None] * num_proc
    assert scw.main(cfg_file, dispatchers=1, timeout=10, run=fake_run, workers=2) == [None, None]
    assert seen['addr'].startswith('grpc://')
    a = scw.parse_args(['cfg.json', '--dispatchers', '2', '--dispatcher-side', 'training'])
    assert (a.configfile, a.dispatchers, a.dispatcher_side, a.timeout) == ('cfg.json', 2, 'training', 60)


if __name__ == '__main__':
    import pathlib, tempfile
    with tempfile.TemporaryDirectory() as d:
        test_spark_compute_worker_driver(pathlib.Path(d))
    with tempfile.TemporaryDirectory() as d:
        test_config_round_trip(pathlib.Path(d))
    test_compute_side_dispatchers_end_to_end()
    test_training_side_dispatcher_and_restrictions()
    print('TF DATA SERVICE OK')
