export PYTHONPATH=.
mkdir -p gpurun_out
timeout 100 python - <<'PY'
import json, torch
import horovod_b200.torch as hvd
hvd.init()
hvd.start_timeline('/tmp/tl.json', mark_cycles=True)
x = torch.ones(1 << 22, device='cuda')
for i in range(5):
    hvd.allreduce_(x, op=hvd.Sum, name='tl.x', prescale_factor=0.5)
    y = hvd.allgather(torch.ones(1024, device='cuda'), name='tl.ag')
torch.cuda.synchronize()
import time; time.sleep(0.3)
live = json.loads(open('/tmp/tl.json').read())
hvd.stop_timeline()
hvd.shutdown()
ev = json.loads(open('/tmp/tl.json').read())
gpu = [e for e in ev if isinstance(e, dict) and e.get('tid') == 1 and e.get('ph') == 'X']
names = sorted({e['name'] for e in gpu})
print('live events', len(live), 'final', len(ev), 'device-timed spans', len(gpu), names, 'dur us', [e['dur'] for e in gpu[:4]])
assert any('GPU ALLREDUCE' == n for n in names), names
open('gpurun_out/timeline_1gpu_sample.json', 'w').write(json.dumps(ev[:400]))
print('TIMELINE GPU OK')
PY
