"""Collectives as CUDA-graph nodes, on >= 2 GPUs (run under hvdrun):

1. hvd.captured_allreduce_ captured into a torch.cuda.graph together with the kernels that produce and consume the
   tensor; replayed with changing inputs, checked against closed forms (Sum / Average / Max, fp32 and bf16).
2. hvd.GraphedStep with the gradient buckets reduced INSIDE the graph: trajectory equal to a plain PyTorch reference
   that averages gradients with ordinary allreduces; parameters bit-identical across ranks; no negotiated allreduce
   per step (engine metrics).
3. the reference idiom model.zero_grad() (set_to_none=True) with zero-copy buckets in eager mode.
4. hvd.join() while a zero-copy bucket response is cached: the joined rank must contribute zeros (no hang, no stale data).
"""
import copy
import sys

import torch

import horovod_b200.torch as hvd

hvd.init()
rank, size = hvd.rank(), hvd.size()
torch.cuda.set_device(hvd.local_rank())
dev = torch.device('cuda', hvd.local_rank())


def check_captured_allreduce():
    n = (1 << 20) + 64
    for dtype, op in ((torch.float32, hvd.Sum), (torch.float32, hvd.Average), (torch.bfloat16, hvd.Sum), (torch.float32, hvd.Max)):
        buf = hvd.symm_empty(n, dtype=dtype)
        src = torch.zeros(n, device=dev, dtype=dtype)
        out = torch.zeros(n, device=dev, dtype=torch.float32)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):  # warm-up outside capture (same sequence on every rank)
            buf.copy_(src)
            hvd.captured_allreduce_(buf, op=op)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            buf.copy_(src)                       # producer
            hvd.captured_allreduce_(buf, op=op)  # collective = one kernel node
            out.copy_(buf.float() * 2.0)         # consumer
        for it in range(4):
            src.fill_(float(rank + 1 + it))
            g.replay()
            torch.cuda.synchronize()
            vals = [float(r + 1 + it) for r in range(size)]
            exp = {hvd.Sum: sum(vals), hvd.Average: sum(vals) / size, hvd.Max: max(vals)}[op] * 2.0
            assert torch.allclose(out, torch.full_like(out, exp), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6), (dtype, op, it, out[:4], exp)
    st = hvd.runtime_stats()
    assert st['captured_collectives'] >= 8, st
    # not registered memory -> a clear error, not a hang
    try:
        hvd.captured_allreduce_(torch.zeros(1024, device=dev))
        raise SystemExit('captured_allreduce_ accepted an unregistered tensor')
    except hvd.HorovodInternalError:
        pass
    print('[ok] captured_allreduce', flush=True)


def _mlp():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                               torch.nn.Linear(256, 10)).to(dev)


def check_graphed_step_comm_in_graph(wire=None):
    steps = 6
    xs = [torch.randn(32, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * rank + i)) for i in range(steps)]
    ys = [torch.randint(0, 10, (32,), device=dev, generator=torch.Generator(device=dev).manual_seed(300 * rank + i)) for i in range(steps)]
    # reference: plain torch, gradients averaged with ordinary (negotiated) allreduces
    ref = _mlp()
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    for x, y in zip(xs, ys):
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        for i, p in enumerate(ref.parameters()):
            p.grad.copy_(hvd.allreduce(p.grad, op=hvd.Average, name='ref.g%d' % i))
        ropt.step()
    model = _mlp()
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9), named_parameters=model.named_parameters(),
                                   fused=True, bucket_cap_mb=0.05, bucket_wire_dtype=wire)  # several small buckets
    assert opt._zero_copy and len(opt._buckets) >= 2, (opt._zero_copy, len(opt._buckets))
    step = hvd.GraphedStep(lambda x, y: torch.nn.functional.cross_entropy(model(x), y), opt, (xs[0], ys[0]), warmup_iters=2)
    assert step.captured, step.fallback_reason
    assert step.comm_in_graph, 'the gradient allreduces were not captured into the graph'
    m0 = hvd.metrics().get('allreduce', {}).get('responses', 0)
    c0 = hvd.runtime_stats()['captured_collectives']
    losses = [step(x, y) for x, y in zip(xs, ys)]
    torch.cuda.synchronize()
    m1 = hvd.metrics().get('allreduce', {}).get('responses', 0)
    assert m1 == m0, ('graph replays must not negotiate allreduces', m0, m1)
    assert hvd.runtime_stats()['captured_collectives'] == c0, 'replays launch no new kernels from the host'
    tol = dict(rtol=2e-4, atol=2e-5) if wire is None else dict(rtol=5e-2, atol=5e-3)
    for (n_, a), (_, b) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(a, b, msg=lambda m, n_=n_: f'{n_}: {m}', **tol)
    # identical parameters on every rank (bit for bit: every rank applies the same reduced gradient)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    allp = hvd.allgather(flat.unsqueeze(0), name='graph.params.%s' % wire)
    for r in range(size):
        assert torch.equal(allp[r], allp[0]), 'parameters differ between ranks %d and 0' % r
    print('[ok] graphed_step_comm_in_graph wire=%s (%d buckets)' % (wire, len(opt._buckets)), flush=True)


def check_model_zero_grad_idiom():
    model = _mlp()
    ref = copy.deepcopy(model)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=model.named_parameters(), bucket_cap_mb=0.05)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    assert opt._zero_copy
    for it in range(3):
        x = torch.randn(16, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(1000 + 10 * rank + it))
        model.zero_grad()  # torch default set_to_none=True drops the bucket views
        model(x).square().mean().backward()
        opt.step()
        ref.zero_grad()
        ref(x).square().mean().backward()
        for i, p in enumerate(ref.parameters()):
            p.grad.copy_(hvd.allreduce(p.grad, op=hvd.Average, name='zg.g%d' % i))
        ropt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    print('[ok] model_zero_grad_idiom', flush=True)


def check_join_with_cached_bucket():
    """Rank 0 runs one more step than the others; the others hvd.join().  The bucket response is cached from the earlier
    steps (zero-copy, symm_key >= 0): while a rank is joined it must go through the fusion-buffer path with zeros."""
    n = 1 << 18
    buf = hvd.symm_empty(n, dtype=torch.float32)
    for it in range(3):  # populate the response cache with the zero-copy response
        buf.fill_(float(rank + 1))
        hvd.allreduce_(buf, op=hvd.Sum, name='join.bucket')
        torch.cuda.synchronize()
        assert float(buf[0]) == sum(range(1, size + 1)), buf[:4]
    if rank == 0:
        buf.fill_(5.0)
        hvd.allreduce_(buf, op=hvd.Sum, name='join.bucket')
        torch.cuda.synchronize()
        assert float(buf[0]) == 5.0 and float(buf[-1]) == 5.0, ('joined ranks must contribute zeros', buf[:4])
    last = hvd.join(hvd.local_rank())
    assert last == 0, last
    # and the zero-copy path is back afterwards
    buf.fill_(1.0)
    hvd.allreduce_(buf, op=hvd.Sum, name='join.bucket')
    torch.cuda.synchronize()
    assert float(buf[0]) == float(size)
    print('[ok] join_with_cached_bucket', flush=True)


def check_ipc_registration():
    """Plain (cudaMalloc) tensors reduced in place: peers map each other's allocation over CUDA IPC and the zero-copy
    kernel runs on them; a tensor that moves to another allocation is re-registered; MIN/MAX and out-of-place calls and
    small tensors keep working (packed path)."""
    import os
    n = 3 << 20  # 12 MiB fp32
    x = torch.empty(n, device=dev)
    c0 = hvd.runtime_stats()['ipc_zero_copy_allreduces']
    for it in range(4):
        x.copy_(torch.arange(n, device=dev) % 7 + rank + it)
        hvd.allreduce_(x, op=hvd.Sum, name='ipc.t')
        exp = (torch.arange(n, device=dev) % 7) * size + sum(range(size)) + it * size
        assert torch.equal(x, exp.float()), (it, x[:4], exp[:4])
    used = hvd.runtime_stats()['ipc_zero_copy_allreduces'] - c0
    expect_ipc = size <= int(os.environ.get('HVD_IPC_MAX_RANKS', '4')) and os.environ.get('HVD_IPC_REGISTRATION', '1') != '0'
    assert (used == 4) == expect_ipc, (used, expect_ipc)
    # the same name on a NEW allocation (the old one stays alive so the address differs): renegotiated, re-registered
    keep = x
    y = torch.empty(n + 1024, device=dev)[:n]
    y.fill_(float(rank + 1))
    hvd.allreduce_(y, op=hvd.Sum, name='ipc.t')
    assert float(y[0]) == sum(range(1, size + 1)) and float(y[-1]) == sum(range(1, size + 1))
    # averaged, bf16, and a view at an offset inside a larger allocation
    big = torch.empty(4 * n, device=dev, dtype=torch.bfloat16)
    v = big[n:3 * n]
    v.fill_(float(rank + 1))
    hvd.allreduce_(v, op=hvd.Average, name='ipc.view')
    assert torch.allclose(v.float(), torch.full_like(v, (size + 1) / 2.0).float(), rtol=1e-2)
    # MAX goes through the in-place kernel's P2P path on any team size
    m = torch.full((n,), float(rank), device=dev)
    hvd.allreduce_(m, op=hvd.Max, name='ipc.max')
    assert float(m[0]) == size - 1 and float(m[-1]) == size - 1
    # out-of-place and small tensors: packed path, same values
    z = hvd.allreduce(torch.full((n,), 2.0, device=dev), op=hvd.Sum, name='ipc.oop')
    assert float(z[0]) == 2.0 * size
    del keep
    print('[ok] ipc_registration (%d zero-copy launches on plain tensors)' % (hvd.runtime_stats()['ipc_zero_copy_allreduces'] - c0), flush=True)


def check_latency_lane():
    """A small allreduce issued right after a very large one must not wait for it: small responses run on their own
    stream / barrier channel / buffer tail (HVD_LATENCY_LANE_BYTES)."""
    import os
    import time
    big = torch.ones(96 << 20, device=dev)      # 384 MiB fp32: milliseconds on the wire
    small = torch.full((1024,), float(rank + 1), device=dev)
    exp_small = float(sum(range(1, size + 1)))
    for _ in range(2):  # warm-up: negotiation, caches, IPC registration of `big`
        hvd.allreduce_(big, op=hvd.Sum, name='lane.big')
        hvd.allreduce_(small.clone(), op=hvd.Sum, name='lane.small')
    torch.cuda.synchronize()
    hvd.barrier()
    ratios = []
    for it in range(3):
        big.fill_(1.0)
        s1 = small.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hb = hvd.allreduce_async_(big, op=hvd.Sum, name='lane.big')
        hs = hvd.allreduce_async_(s1, op=hvd.Sum, name='lane.small')
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            hvd.synchronize(hs)          # the side stream waits for the small collective only
            side.synchronize()
            t_small = time.perf_counter() - t0
        hvd.synchronize(hb)
        torch.cuda.current_stream().synchronize()
        t_big = time.perf_counter() - t0
        assert float(s1[0]) == exp_small and float(big[0]) == float(size) and float(big[-1]) == float(size)
        ratios.append(t_small / t_big)
    lane_on = int(os.environ.get('HVD_LATENCY_LANE_BYTES', str(256 << 10))) > 0
    best = min(ratios)
    print('[info] latency lane %s: small/big completion time ratio %.2f' % ('on' if lane_on else 'off', best), flush=True)
    if lane_on:
        assert best < 0.6, ('the small allreduce waited for the large one', ratios)
    print('[ok] latency_lane', flush=True)


only = set(sys.argv[1].split(',')) if len(sys.argv) > 1 else None
checks = [('captured', check_captured_allreduce), ('graphed', check_graphed_step_comm_in_graph),
          ('graphed_bf16', lambda: check_graphed_step_comm_in_graph(torch.bfloat16)),
          ('zero_grad', check_model_zero_grad_idiom), ('join', check_join_with_cached_bucket), ('ipc', check_ipc_registration), ('lane', check_latency_lane)]
for name, fn in checks:
    if only is None or name in only:
        fn()
hvd.barrier()
print('GRAPH COMM OK', flush=True)
hvd.shutdown()
