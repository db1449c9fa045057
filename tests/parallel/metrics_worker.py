"""hvd.metrics(): counters per collective type, Interval deltas, Prometheus endpoint."""
import urllib.request

import torch

import horovod_b200.torch as hvd
from horovod_b200.utils import metrics

hvd.init()
rank, size = hvd.rank(), hvd.size()
base = hvd.metrics()
assert 'runtime' in base and base['runtime']['cycles'] >= 0

with metrics.Interval() as m:
    for i in range(5):
        hvd.allreduce(torch.ones(1000), name='m.ar.%d' % i)                    # 5 x 4000 B
    hs = [hvd.allreduce_async(torch.ones(10), name='m.fused.%d' % i) for i in range(8)]
    for h in hs:
        hvd.synchronize(h)
    hvd.allgather(torch.ones(rank + 1, 2), name='m.ag')
    hvd.broadcast(torch.ones(7, dtype=torch.float64), root_rank=0, name='m.bc')
    try:
        hvd.allreduce(torch.ones(3 + rank), name='m.bad')                       # shape mismatch -> ERROR response
        raise SystemExit('expected an error')
    except hvd.HorovodInternalError:
        pass
d = m.delta
assert d['allreduce']['tensors'] == 13, d
assert 5 <= d['allreduce']['responses'] <= 13 and d['allreduce']['bytes'] == 5 * 4000 + 8 * 40, d     # fusion: fewer responses than tensors
assert d['allgather']['tensors'] == 1 and d['allgather']['bytes'] == (rank + 1) * 2 * 4, d
assert d['broadcast']['bytes'] == 56 and d['broadcast']['on_gpu'] == 0, d
assert d['error']['responses'] == 1, d
assert m.seconds > 0 and m.rate('allreduce', 'bytes') > 0
assert d['runtime']['responses'] >= 8 and d['runtime']['cycles'] > 0

assert hvd.metrics()['host_paths']['shared_memory'] > 0 and hvd.metrics()['host_paths']['two_level'] == 0
flat = metrics.flatten(hvd.metrics())
assert flat['hvd_allreduce_tensors'] >= 13 and 'hvd_runtime_cycles' in flat

server = metrics.start_prometheus(0, addr='127.0.0.1')
port = server.server_address[1]
text = urllib.request.urlopen('http://127.0.0.1:%d/metrics' % port, timeout=10).read().decode()
assert '# TYPE hvd_allreduce_bytes counter' in text and 'hvd_allreduce_bytes{local_rank="%d",rank="%d",size="%d"}' % (hvd.local_rank(), rank, size) in text, text
try:
    urllib.request.urlopen('http://127.0.0.1:%d/nope' % port, timeout=10)
    raise SystemExit('expected 404')
except urllib.error.HTTPError as e:
    assert e.code == 404
server.shutdown()

import logging
records = []
handler = logging.Handler()
handler.emit = lambda rec: records.append(rec.getMessage())
log = logging.getLogger('metrics-test')
log.setLevel(logging.INFO)
log.addHandler(handler)
stop = metrics.log_every(0.2, logger=log)
hvd.allreduce(torch.ones(256), name='m.logged')
import time
time.sleep(0.7)
stop.set()
assert any('allreduce: 1 ops' in r for r in records), records

hvd.barrier()
if rank == 0:
    print('METRICS OK')
hvd.shutdown()
