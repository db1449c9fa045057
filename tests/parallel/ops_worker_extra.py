"""Second batch of distributed checks, one per remaining case family of the reference's test/parallel/test_torch.py
(grad variants with process sets, grouped allgather / reducescatter, per-op error texts, sparse gradients, optimizer
corner cases, join with non-allreduce ops, barriers mixed with collectives).  Same harness as ops_worker.py."""
import argparse
import sys
import time
import traceback
import warnings

import torch

import horovod_b200.torch as hvd
from horovod_b200.common.exceptions import HorovodInternalError

p = argparse.ArgumentParser()
p.add_argument('--device', default='cpu')
p.add_argument('--only', default='')
args = p.parse_args()
hvd.init()
rank, size = hvd.rank(), hvd.size()
DEV = torch.device('cuda', hvd.local_rank()) if args.device == 'cuda' else torch.device('cpu')
if DEV.type == 'cuda':
    torch.cuda.set_device(DEV)
CHECKS = []


def check(name):
    def deco(fn):
        fn._check_name = name
        CHECKS.append(fn)
        return fn
    return deco


def raises(exc, fn, needle=None):
    try:
        fn()
    except exc as e:
        if needle is not None:
            assert needle.lower() in str(e).lower(), (needle, str(e))
        return
    raise AssertionError('expected %s' % (exc,))


def even_set():
    return hvd.add_process_set([r for r in range(size) if r % 2 == 0]) if size >= 2 else hvd.global_process_set


@check('allreduce_grad_process_sets')
def _():
    ps = even_set()
    try:
        if ps.included():
            x = torch.ones(5, device=DEV, requires_grad=True)
            y = hvd.allreduce(x, op=hvd.Sum, process_set=ps, name='g.ps')
            y.backward(torch.ones(5, device=DEV) * (rank + 1))
            members = [r for r in range(size) if r % 2 == 0]
            assert torch.allclose(x.grad, torch.full((5,), float(sum(r + 1 for r in members)), device=DEV)), x.grad
            x2 = torch.ones(5, device=DEV, requires_grad=True)
            hvd.allreduce(x2, op=hvd.Average, process_set=ps, name='g.ps.avg').sum().backward()
            assert torch.allclose(x2.grad, torch.ones(5, device=DEV))
    finally:
        if ps is not hvd.global_process_set:
            hvd.remove_process_set(ps)


@check('grouped_allreduce_grad')
def _():
    xs = [torch.ones(4, device=DEV, requires_grad=True), torch.ones(2, 3, device=DEV, requires_grad=True)]
    ys = hvd.grouped_allreduce(xs, op=hvd.Average, name='gg')
    (ys[0].sum() * 2 + ys[1].sum() * 3).backward()
    assert torch.allclose(xs[0].grad, torch.full((4,), 2.0, device=DEV)) and torch.allclose(xs[1].grad, torch.full((2, 3), 3.0, device=DEV))
    zs = [torch.ones(3, device=DEV) * (rank + 1) for _ in range(3)]
    outs = hvd.grouped_allreduce_(zs, op=hvd.Sum, name='gg.inplace')
    tot = float(size * (size + 1) // 2)
    assert all(o is z for o, z in zip(outs, zs)) and all(torch.allclose(z, torch.full((3,), tot, device=DEV)) for z in zs)


@check('allgather_grad_and_variable')
def _():
    rows = rank + 1
    x = torch.ones(rows, 3, device=DEV, requires_grad=True)
    y = hvd.allgather(x, name='ag.grad')
    assert y.shape[0] == size * (size + 1) // 2
    w = torch.cat([torch.full((r + 1, 3), float(r + 1)) for r in range(size)]).to(DEV)
    (y * w).sum().backward()
    # the upstream gradients are AVERAGED over ranks, then sliced (reference torch/mpi_ops.py:642): every rank weighted my
    # rows with rank + 1
    assert torch.allclose(x.grad, torch.full((rows, 3), float(rank + 1), device=DEV)), x.grad
    hs = [hvd.allgather_async(torch.full((2, 2), float(rank + i), device=DEV), name=f'ag.async.{i}') for i in range(5)]
    for i, h in enumerate(hs):
        out = hvd.synchronize(h)
        assert out.shape == (2 * size, 2) and torch.allclose(out[2 * (size - 1)], torch.full((2,), float(size - 1 + i), device=DEV))


@check('grouped_allgather_and_grad')
def _():
    xs = [torch.full((rank + 1, 2), float(rank), device=DEV, requires_grad=True), torch.ones(2, device=DEV, requires_grad=True)]
    ys = hvd.grouped_allgather(xs, name='gag')
    assert ys[0].shape[0] == size * (size + 1) // 2 and ys[1].shape[0] == 2 * size
    (ys[0].sum() + 2 * ys[1].sum()).backward()
    assert torch.allclose(xs[0].grad, torch.ones(rank + 1, 2, device=DEV))
    assert torch.allclose(xs[1].grad, torch.full((2,), 2.0, device=DEV))
    ps = even_set()
    try:
        if ps.included() and ps is not hvd.global_process_set:
            outs = hvd.grouped_allgather([torch.ones(1, device=DEV) * rank, torch.ones(2, device=DEV)], process_set=ps, name='gag.ps')
            members = [r for r in range(size) if r % 2 == 0]
            assert outs[0].tolist() == [float(r) for r in members] and outs[1].numel() == 2 * len(members)
    finally:
        if ps is not hvd.global_process_set:
            hvd.remove_process_set(ps)


@check('broadcast_variants')
def _():
    for root in range(size):
        t = torch.full((3, 2), float(rank), device=DEV)
        out = hvd.broadcast(t, root_rank=root, name=f'bc.{root}')
        assert torch.allclose(out, torch.full((3, 2), float(root), device=DEV)) and torch.allclose(t, torch.full((3, 2), float(rank), device=DEV))
        hvd.broadcast_(t, root_rank=root, name=f'bc_.{root}')
        assert torch.allclose(t, torch.full((3, 2), float(root), device=DEV))
    x = torch.ones(4, device=DEV, requires_grad=True)
    y = hvd.broadcast(x, root_rank=0, name='bc.grad')
    y.backward(torch.ones(4, device=DEV))
    exp = 1.0 if rank == 0 else 0.0  # averaged upstream gradient on the root, zero elsewhere (reference torch/mpi_ops.py:824-828)
    assert torch.allclose(x.grad, torch.full((4,), exp, device=DEV)), x.grad
    if size >= 2:
        raises((HorovodInternalError, ValueError), lambda: hvd.broadcast(torch.ones(2, device=DEV), root_rank=size + 3, name='bc.badroot'))
    # the library is still usable
    assert hvd.allreduce(torch.ones(1, device=DEV), op=hvd.Sum).item() == size


@check('alltoall_variants')
def _():
    # equal split, explicit splits (also as a tensor on the compute device), grads, errors
    x = (torch.arange(size * 2, dtype=torch.float32) + 100 * rank).to(DEV)
    out = hvd.alltoall(x, name='a2a.eq')
    assert out.tolist() == [100.0 * q + 2 * rank + j for q in range(size) for j in range(2)]
    splits = torch.tensor([r + 1 for r in range(size)], dtype=torch.int32, device=DEV)
    src = torch.cat([torch.full((r + 1, 2), float(rank * 10 + r)) for r in range(size)]).to(DEV)
    out, rs = hvd.alltoall(src, splits=splits, name='a2a.splits.dev')
    assert rs.tolist() == [rank + 1] * size and out.shape == (size * (rank + 1), 2)
    assert torch.allclose(out[:rank + 1], torch.full((rank + 1, 2), float(rank), device=DEV))
    xg = torch.ones(size * 2, device=DEV, requires_grad=True)
    hvd.alltoall(xg, name='a2a.grad').sum().backward()
    assert torch.allclose(xg.grad, torch.ones(size * 2, device=DEV))
    xg2 = torch.ones(size * (size + 1) // 2, device=DEV, requires_grad=True)
    o2, _ = hvd.alltoall(xg2, splits=[r + 1 for r in range(size)], name='a2a.grad.splits')
    (o2 * (rank + 1)).sum().backward()
    exp = torch.cat([torch.full((r + 1,), float(r + 1)) for r in range(size)]).to(DEV)
    assert torch.allclose(xg2.grad, exp), xg2.grad
    raises((HorovodInternalError, ValueError), lambda: hvd.alltoall(torch.ones(size * 2 + 1, device=DEV), name='a2a.err.len'))
    raises((HorovodInternalError, ValueError), lambda: hvd.alltoall(torch.ones(4, device=DEV), splits=[1] * (size + 1), name='a2a.err.nsplits'))
    raises((HorovodInternalError, ValueError), lambda: hvd.alltoall(torch.ones(4, device=DEV), splits=[-1] + [5] + [0] * (size - 2) if size >= 2 else [-1],
                                                                   name='a2a.err.neg'))
    raises((HorovodInternalError, ValueError, TypeError), lambda: hvd.alltoall(torch.ones(size, device=DEV),
                                                                              splits=torch.ones(size, dtype=torch.float32), name='a2a.err.type'))
    assert hvd.allreduce(torch.ones(1, device=DEV), op=hvd.Sum).item() == size


@check('reducescatter_variants')
def _():
    x = torch.ones(size * 2, 3, device=DEV) * (rank + 1)
    tot = size * (size + 1) / 2
    assert torch.allclose(hvd.reducescatter(x, op=hvd.Sum, name='rs.sum'), torch.full((2, 3), tot, device=DEV))
    assert torch.allclose(hvd.reducescatter(x, op=hvd.Average, name='rs.avg'), torch.full((2, 3), tot / size, device=DEV))
    assert torch.allclose(hvd.reducescatter(x, op=hvd.Sum, prescale_factor=0.5, name='rs.pre'), torch.full((2, 3), tot / 2, device=DEV))
    assert torch.allclose(hvd.reducescatter(x, op=hvd.Sum, postscale_factor=2.0, name='rs.post'), torch.full((2, 3), tot * 2, device=DEV))
    # uneven first dimension: earlier ranks get the extra rows
    u = torch.ones(size + 1, 2, device=DEV)
    out = hvd.reducescatter(u, op=hvd.Sum, name='rs.uneven')
    assert out.shape[0] == (2 if rank == 0 else 1) and torch.allclose(out, torch.full_like(out, float(size)))
    hs = [hvd.reducescatter_async(torch.ones(size, 4, device=DEV) * i, op=hvd.Sum, name=f'rs.async.{i}') for i in range(4)]
    for i, h in enumerate(hs):
        assert torch.allclose(hvd.synchronize(h), torch.full((1, 4), float(i * size), device=DEV))
    xg = torch.ones(size * 2, device=DEV, requires_grad=True)
    y = hvd.reducescatter(xg, op=hvd.Sum, name='rs.grad')
    y.backward(torch.ones(2, device=DEV) * (rank + 1))
    # reference semantics (torch/mpi_ops.py:1083-1091): Sum scales the upstream gradient by size, Average does not
    exp = torch.cat([torch.full((2,), float(size * (r + 1))) for r in range(size)]).to(DEV)
    assert torch.allclose(xg.grad, exp), xg.grad
    xa = torch.ones(size * 2, device=DEV, requires_grad=True)
    hvd.reducescatter(xa, op=hvd.Average, name='rs.grad.avg').sum().backward()
    assert torch.allclose(xa.grad, torch.ones(size * 2, device=DEV)), xa.grad
    raises((HorovodInternalError, ValueError), lambda: hvd.reducescatter(torch.tensor(1.0, device=DEV), name='rs.scalar'))
    raises((HorovodInternalError, ValueError, NotImplementedError), lambda: hvd.reducescatter(x, op=hvd.Adasum, name='rs.adasum'))
    if size >= 2:
        raises(HorovodInternalError, lambda: hvd.reducescatter(torch.ones(size, rank + 1, device=DEV), name='rs.shape'))
        raises(HorovodInternalError, lambda: hvd.reducescatter(torch.ones(size, 2, device=DEV, dtype=torch.float32 if rank % 2 else torch.float64),
                                                               name='rs.dtype'))
    outs = hvd.grouped_reducescatter([x, torch.ones(size, device=DEV)], op=hvd.Sum, name='grs')
    assert torch.allclose(outs[0], torch.full((2, 3), tot, device=DEV)) and torch.allclose(outs[1], torch.full((1,), float(size), device=DEV))
    outs = hvd.grouped_reducescatter([x, x], op=hvd.Average, prescale_factor=2.0, postscale_factor=0.5, name='grs.scaled')
    assert all(torch.allclose(o, torch.full((2, 3), tot / size, device=DEV)) for o in outs)
    xs = [torch.ones(size, device=DEV, requires_grad=True) for _ in range(2)]
    ys = hvd.grouped_reducescatter(xs, op=hvd.Sum, name='grs.grad')
    (ys[0].sum() + 2 * ys[1].sum()).backward()
    assert torch.allclose(xs[0].grad, torch.full((size,), float(size), device=DEV)) and torch.allclose(xs[1].grad, torch.full((size,), 2.0 * size, device=DEV))
    assert hvd.allreduce(torch.ones(1, device=DEV), op=hvd.Sum).item() == size


@check('duplicate_names_per_op')
def _():
    # rank 0 submits late: until then the first submission of the other ranks cannot start executing, so their second one
    # deterministically finds the name in flight (see ops_worker.py 'errors')
    for fn, kw in ((hvd.allgather_async, {}), (hvd.broadcast_async, {'root_rank': 0}), (hvd.reducescatter_async, {})):
        t = torch.ones(size * 64, 16, device=DEV)
        if rank == 0:
            time.sleep(0.5)
            hvd.synchronize(fn(t, name='dupname', **kw))
        else:
            h = fn(t, name='dupname', **kw)
            try:
                fn(t, name='dupname', **kw)
                raise AssertionError('duplicate in-flight name must be rejected')
            except (ValueError, HorovodInternalError) as e:
                assert 'dupname' in str(e) or 'duplicate' in str(e).lower() or 'same name' in str(e).lower(), e
            hvd.synchronize(h)
        hvd.barrier()


@check('sparse_gradients')
def _():
    torch.manual_seed(1234)
    emb = torch.nn.Embedding(10, 3, sparse=True).to(DEV)
    hvd.broadcast_parameters(emb.state_dict(), root_rank=0)
    w0 = emb.weight.detach().clone()
    opt = hvd.DistributedOptimizer(torch.optim.SGD(emb.parameters(), lr=1.0), named_parameters=emb.named_parameters())
    idx = torch.tensor([rank % 10, (rank + 1) % 10], device=DEV)
    opt.zero_grad()
    emb(idx).sum().backward()
    # CPU / zero_copy=False: the gradient stays sparse and is reduced through allgather (reference mpi_ops.py:567-588);
    # with zero-copy buckets the sparse gradient is accumulated into the dense registered bucket view
    assert emb.weight.grad.is_sparse or getattr(opt, '_zero_copy', False)
    opt.step()
    exp = w0.clone()
    for r in range(size):
        for i in (r % 10, (r + 1) % 10):
            exp[i] -= 1.0 / size
    assert torch.allclose(emb.weight.detach(), exp, atol=1e-6), (emb.weight.detach() - exp).abs().max()
    emb2 = torch.nn.Embedding(10, 3, sparse=True).to(DEV)
    hvd.broadcast_parameters(emb2.state_dict(), root_rank=0)
    opt2 = hvd.DistributedOptimizer(torch.optim.SGD(emb2.parameters(), lr=1.0), named_parameters=emb2.named_parameters(), sparse_as_dense=True)
    opt2.zero_grad()
    emb2(idx).sum().backward()
    opt2.step()
    w = hvd.allgather(emb2.weight.detach().reshape(1, -1))
    assert all(torch.allclose(w[r], w[0]) for r in range(size))
    ps = even_set()
    try:
        if ps.included():
            sp = torch.sparse_coo_tensor(torch.tensor([[rank % 4]]), torch.tensor([1.0]), (4,)).to(DEV)
            out = hvd.synchronize(hvd.sparse_allreduce_async(sp, name='sp.ps', op=hvd.Sum, process_set=ps)).to_dense()
            members = [r for r in range(size) if r % 2 == 0] if ps is not hvd.global_process_set else list(range(size))
            exp = torch.zeros(4)
            for r in members:
                exp[r % 4] += 1
            assert torch.allclose(out.cpu(), exp), out
    finally:
        if ps is not hvd.global_process_set:
            hvd.remove_process_set(ps)


@check('optimizer_corner_cases')
def _():
    model = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)).to(DEV)
    # no named_parameters
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1))
    opt.zero_grad()
    model(torch.ones(2, 4, device=DEV)).sum().backward()
    opt.step()
    # missing / duplicate names
    raises(ValueError, lambda: hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1),
                                                        named_parameters=list(model.named_parameters())[:1]), 'not named')
    raises(ValueError, lambda: hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1),
                                                        named_parameters=[('a', q) for q in model.parameters()]), 'unique')
    raises(ValueError, lambda: hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=list(model.parameters())),
           'tuples')
    # dynamic requires_grad: a parameter frozen after the optimizer was built must not hang the step
    m2 = torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.Linear(3, 1)).to(DEV)
    hvd.broadcast_parameters(m2.state_dict(), 0)
    o2 = hvd.DistributedOptimizer(torch.optim.SGD(m2.parameters(), lr=0.1), named_parameters=m2.named_parameters())
    for it in range(3):
        m2[0].weight.requires_grad_(it % 2 == 0)
        o2.zero_grad()
        m2(torch.ones(2, 3, device=DEV) * (rank + 1)).sum().backward()
        o2.step()
    w = hvd.allgather(torch.cat([q.detach().reshape(-1) for q in m2.parameters()]).reshape(1, -1))
    assert all(torch.allclose(w[r], w[0], atol=1e-6) for r in range(size))
    # synchronize() then step() without skip_synchronize() warns
    o2.zero_grad()
    m2[0].weight.requires_grad_(True)
    m2(torch.ones(2, 3, device=DEV)).sum().backward()
    o2.synchronize()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        o2.step()
    assert any('skip_synchronize' in str(x.message) for x in rec)
    # fp16 compression keeps dtype and value
    m3 = torch.nn.Linear(4, 2).to(DEV)
    hvd.broadcast_parameters(m3.state_dict(), 0)
    o3 = hvd.DistributedOptimizer(torch.optim.SGD(m3.parameters(), lr=0.5), named_parameters=m3.named_parameters(), compression=hvd.Compression.fp16)
    w_before = m3.weight.detach().clone()
    o3.zero_grad()
    m3(torch.ones(1, 4, device=DEV) * (rank + 1)).sum().backward()
    o3.step()
    assert m3.weight.grad.dtype == torch.float32
    assert torch.allclose(m3.weight.detach(), w_before - 0.5 * (size + 1) / 2, atol=1e-2)
    # optimizer restricted to a process set
    ps = even_set()
    try:
        m4 = torch.nn.Linear(2, 1).to(DEV)
        with torch.no_grad():
            m4.weight.fill_(1.0)
            m4.bias.zero_()
        o4 = hvd.DistributedOptimizer(torch.optim.SGD(m4.parameters(), lr=1.0), named_parameters=m4.named_parameters(), process_set=ps)
        o4.zero_grad()
        m4(torch.ones(1, 2, device=DEV) * (rank + 1)).sum().backward()
        o4.step()
        if ps.included():
            members = [r for r in range(size) if r % 2 == 0] if ps is not hvd.global_process_set else list(range(size))
            mean = sum(r + 1 for r in members) / len(members)
            assert torch.allclose(m4.weight.detach(), torch.full((1, 2), 1.0 - mean, device=DEV)), m4.weight
        else:
            assert torch.allclose(m4.weight.detach(), torch.full((1, 2), 1.0 - (rank + 1), device=DEV))
    finally:
        if ps is not hvd.global_process_set:
            hvd.remove_process_set(ps)


@check('join_with_other_ops')
def _():
    if size < 2:
        return
    # a joined rank cannot serve allgather / broadcast: the active ranks get an error, then everybody joins
    if rank == 0:
        raises(HorovodInternalError, lambda: hvd.allgather(torch.ones(2, device=DEV), name='join.ag'), 'join')
        raises(HorovodInternalError, lambda: hvd.broadcast(torch.ones(2, device=DEV), root_rank=0, name='join.bc'), 'join')
        if size == 2:
            pass
    if rank != 0:
        # give rank 0 time to submit; ranks != 0 go straight to join only when they are not needed for the error path
        pass
    if rank == 0 or True:
        if rank != 0 and size > 2:
            # every non-joined rank must submit the same ops for the coordinator to produce the error response
            raises(HorovodInternalError, lambda: hvd.allgather(torch.ones(2, device=DEV), name='join.ag'), 'join') if rank != size - 1 else None
            raises(HorovodInternalError, lambda: hvd.broadcast(torch.ones(2, device=DEV), root_rank=0, name='join.bc'), 'join') if rank != size - 1 else None
    last = hvd.join()
    assert 0 <= last < size
    hvd.barrier()


@check('dynamic_requires_grad')
def _():
    """GAN-style alternation (reference test_torch.py::test_dynamic_requires_grad): two DistributedOptimizers, parameters
    switched on and off between steps; only the trained half gets (reduced, identical) gradients, the other half none."""
    torch.manual_seed(1234)
    gen, disc = torch.nn.Conv2d(1, 4, 1).to(DEV), torch.nn.Conv2d(4, 1, 1).to(DEV)
    hvd.broadcast_parameters(gen.state_dict(), root_rank=0)
    hvd.broadcast_parameters(disc.state_dict(), root_rank=0)
    inp = torch.rand([1, 1, 8, 8], generator=torch.Generator().manual_seed(100 + rank)).to(DEV)   # rank-specific data
    gen_opt = hvd.DistributedOptimizer(torch.optim.SGD(gen.parameters(), lr=0.1), named_parameters=[('gen.' + n, q) for n, q in gen.named_parameters()])
    disc_opt = hvd.DistributedOptimizer(torch.optim.SGD(disc.parameters(), lr=0.1), named_parameters=[('disc.' + n, q) for n, q in disc.named_parameters()])

    def has_grad(q):
        return q.grad is not None and bool(q.grad.abs().max() > 0)

    def train_step(train_generator=False, train_discriminator=False):
        for q in gen.parameters():
            q.requires_grad_(train_generator)
        for q in disc.parameters():
            q.requires_grad_(train_discriminator)
        gen_opt.zero_grad()
        disc_opt.zero_grad()
        disc(gen(inp)).sum().backward()
        if train_generator:
            gen_opt.step()
        if train_discriminator:
            disc_opt.step()
        for q in gen.parameters():
            assert train_generator == has_grad(q), ('generator', train_generator, q.grad)
        for q in disc.parameters():
            assert train_discriminator == has_grad(q), ('discriminator', train_discriminator, q.grad)

    for _ in range(4):
        train_step(train_generator=True)
        train_step(train_discriminator=True)
    train_step(train_generator=True, train_discriminator=True)
    # different data on every rank, averaged gradients: the replicas must still be identical
    for name, q in list(gen.named_parameters()) + list(disc.named_parameters()):
        got = hvd.allgather(q.detach().reshape(1, -1), name='dyn.' + name + str(q.numel()))
        assert torch.allclose(got, got[0:1].expand_as(got), atol=1e-6), name


@check('missing_named_parameters')
def _():
    """named_parameters that covers only part of the optimizer's parameters is an error (reference
    test_missing_named_parameters); so are duplicate names (test_duplicate_names is in duplicate_names_per_op)."""
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 1), torch.nn.Conv2d(4, 1, 1)).to(DEV)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    raises(ValueError, lambda: hvd.DistributedOptimizer(opt, named_parameters=list(net.named_parameters())[0:1]))


@check('barriers_mixed')
def _():
    ps = even_set()
    try:
        for i in range(3):
            h = hvd.allreduce_async(torch.ones(1 << 12, device=DEV), op=hvd.Sum, name=f'bm.{i}')
            hvd.barrier()
            if ps.included():
                hvd.barrier(process_set=ps)
            assert torch.allclose(hvd.synchronize(h), torch.full((1 << 12,), float(size), device=DEV))
        if ps is not hvd.global_process_set and not ps.included():
            raises((ValueError, HorovodInternalError), lambda: hvd.barrier(process_set=ps))
    finally:
        if ps is not hvd.global_process_set:
            hvd.remove_process_set(ps)


only = [x for x in args.only.split(',') if x]
failed = []
for fn in CHECKS:
    if only and fn._check_name not in only:
        continue
    t0 = time.time()
    try:
        fn()
        hvd.barrier()
        if rank == 0:
            print(f'[ok] {fn._check_name} ({time.time() - t0:.2f}s)', flush=True)
    except Exception:
        traceback.print_exc()
        print(f'[FAIL] rank {rank}: {fn._check_name}', flush=True)
        failed.append(fn._check_name)
        break
if failed:
    sys.exit(1)
hvd.barrier()
if rank == 0:
    print('EXTRA ALL OK', flush=True)
hvd.shutdown()
