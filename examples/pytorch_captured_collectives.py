"""Collectives as CUDA-graph nodes (`hvd.captured_allreduce_`), outside of DistributedOptimizer.

    bin/hvdrun -np 2 python examples/pytorch_captured_collectives.py        (needs >= 2 GPUs on one NVLink box)

A tensor allocated with `hvd.symm_empty` lives in peer-mapped memory; `hvd.captured_allreduce_` reduces it in place with ONE
kernel on the current CUDA stream — no handle, no negotiation round, no host synchronisation — so it can be captured into a
CUDA graph together with the kernels that produce and consume the tensor.  The kernel's own cross-GPU flag barrier is the only
synchronisation; every rank must capture the same sequence of collectives.  `hvd.GraphedStep` uses exactly this to make a whole
data-parallel training step one graph launch (see pytorch_graphed_step.py).

The reference has no counterpart: each of its collectives is negotiated by the background thread every time
(horovod/common/controller.cc:209-252) and cannot be recorded into a CUDA graph.
"""
import torch

import horovod_b200.torch as hvd

hvd.init()
torch.cuda.set_device(hvd.local_rank())
rank, size = hvd.rank(), hvd.size()
if size < 2:
    raise SystemExit('run with at least 2 ranks (one GPU each)')

n = 4 << 20
stats = hvd.symm_empty(n, dtype=torch.float32)          # COLLECTIVE allocation in registered symmetric memory
x = torch.zeros(n, device='cuda')
y = torch.zeros(n, device='cuda')

# warm-up of the same sequence outside capture (every rank issues the same collectives in the same order)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    stats.copy_(x)
    hvd.captured_allreduce_(stats, op=hvd.Average)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()

graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    stats.copy_(x * x)                              # producer
    hvd.captured_allreduce_(stats, op=hvd.Average)  # the collective: one kernel node of this graph
    y.copy_(stats.sqrt())                           # consumer

for step in range(5):
    x.fill_(float(rank + step))
    graph.replay()                                  # compute + communication, no host work besides this launch
    torch.cuda.synchronize()
    expected = (sum((r + step) ** 2 for r in range(size)) / size) ** 0.5
    assert abs(float(y[0]) - expected) < 1e-4 and abs(float(y[-1]) - expected) < 1e-4, (float(y[0]), expected)
if rank == 0:
    print('captured allreduce inside a CUDA graph: 5 replays correct on %d GPUs; host launches of collectives: %d'
          % (size, hvd.runtime_stats()['captured_collectives']))
hvd.shutdown()
