import torch, horovod_b200.torch as hvd
hvd.init()
r, n = hvd.rank(), hvd.size()
for _ in range(4):                      # put the response in the cache
    out = hvd.allreduce(torch.ones(5) * (r + 1), op=hvd.Sum, name='jc')
    assert out[0].item() == n * (n + 1) / 2
if r == 0:
    for i in range(3):                  # rank 0 keeps going through the cache fast path while the others have joined
        out = hvd.allreduce(torch.ones(5) * 10, op=hvd.Sum, name='jc')
        assert out.tolist() == [10.0] * 5, out
        avg = hvd.allreduce(torch.ones(5) * 10, name='jc.avg')          # Average divides by the full size
        assert abs(avg[0].item() - 10.0 / n) < 1e-6, avg
last = hvd.join()
assert last == 0, last
# after join everybody is back: the cached response is still valid
out = hvd.allreduce(torch.ones(5) * (r + 1), op=hvd.Sum, name='jc')
assert out[0].item() == n * (n + 1) / 2
if r == 0: print('JOIN CACHED OK')
hvd.shutdown()
