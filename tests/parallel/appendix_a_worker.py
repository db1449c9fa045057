"""Behavioural details of the reference that programs rely on: legacy `average=` kwarg, default Average,\nobservable async-ness, truncating integer average, join divisor, error type, duplicate names."""
import warnings, torch, horovod_b200.torch as hvd
hvd.init()
r, n = hvd.rank(), hvd.size()
t = torch.ones(4) * (r + 1)
try:
    hvd.allreduce(t, average=True, op=hvd.Sum); raise SystemExit('both average and op must raise')
except ValueError: pass
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    a = hvd.allreduce(t, average=False, name='legacy')
    assert any(issubclass(x.category, DeprecationWarning) for x in w), [x.category for x in w]
assert torch.all(a == sum(range(1, n + 1)))
assert torch.allclose(hvd.allreduce(t, name='default'), torch.ones(4) * sum(range(1, n + 1)) / n)   # default Average
# async-ness observable
big = torch.ones(1 << 22)
seen_false = False
for i in range(20):
    h = hvd.allreduce_async(big, name='poll.%d' % i)
    if not hvd.poll(h): seen_false = True
    hvd.synchronize(h)
assert seen_false
# integer average truncates
ia = hvd.allreduce(torch.tensor([r + 1, 7], dtype=torch.int32), op=hvd.Average, name='iavg')
assert ia.tolist() == [sum(range(1, n + 1)) // n, 7], ia
# join: divisor stays the full size
if r == n - 1:
    last = hvd.join()
else:
    out = hvd.allreduce(torch.ones(3), op=hvd.Average, name='joined.avg')
    assert torch.allclose(out, torch.ones(3) * (n - 1) / n), out
    last = hvd.join()
agreed = hvd.allgather(torch.tensor([last]), name="join.last").tolist()
assert len(set(agreed)) == 1 and (n == 1 or agreed[0] != n - 1), agreed   # the LAST rank to join, identical everywhere
# HorovodInternalError is a RuntimeError
assert issubclass(hvd.HorovodInternalError, RuntimeError)
# duplicate in-flight name (deterministic form: rank 0 submits late, see ops_worker.py 'errors')
import time
dup_err = True
if r == 0:
    time.sleep(0.5)
    hvd.synchronize(hvd.allreduce_async(torch.ones(8), name='dup'))
else:
    h1 = hvd.allreduce_async(torch.ones(8), name='dup')
    try:
        hvd.allreduce_async(torch.ones(8), name='dup')
        dup_err = False
    except (hvd.HorovodInternalError, ValueError) as e:
        dup_err = 'dup' in str(e)
    hvd.synchronize(h1)
dup_err = all(hvd.allgather_object(dup_err))
hvd.barrier()
if r == 0: print('APPENDIX A OK dup_err=%s' % dup_err)
hvd.shutdown()
