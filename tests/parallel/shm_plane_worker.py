"""Host-tensor collectives through the shared-memory data plane: many pieces (HVD_SHM_SLOT_BYTES is set tiny by the test),
odd sizes, every dtype / reduce op, uneven allgather, broadcast from every root, reducescatter with uneven splits, and a
long mixed sequence (the piece counter / double buffering must stay consistent across different collectives)."""
import os

import torch

import horovod_b200.torch as hvd

FAKE_HOSTS = int(os.environ.get('HVD_TEST_FAKE_HOSTS', '0'))
if FAKE_HOSTS > 1:          # present the ranks as FAKE_HOSTS machines: two-level control plane, TCP data plane
    _r, _n = int(os.environ['HOROVOD_RANK']), int(os.environ['HOROVOD_SIZE'])
    _L = _n // FAKE_HOSTS
    os.environ.update(HOROVOD_HOSTNAME='fakehost%d' % (_r // _L), HOROVOD_LOCAL_RANK=str(_r % _L), HOROVOD_LOCAL_SIZE=str(_L),
                      HOROVOD_CROSS_RANK=str(_r // _L), HOROVOD_CROSS_SIZE=str(FAKE_HOSTS))

HOST_MAP = os.environ.get('HVD_TEST_FAKE_HOST_MAP')          # e.g. "0,0,1": uneven hosts
if HOST_MAP:
    _hosts = [int(x) for x in HOST_MAP.split(',')]
    _r = int(os.environ['HOROVOD_RANK'])
    _mine = [i for i, h in enumerate(_hosts) if h == _hosts[_r]]
    _uniq = sorted(set(_hosts))
    os.environ.update(HOROVOD_HOSTNAME='fakehost%d' % _hosts[_r], HOROVOD_LOCAL_RANK=str(_mine.index(_r)), HOROVOD_LOCAL_SIZE=str(len(_mine)),
                      HOROVOD_CROSS_RANK=str(_uniq.index(_hosts[_r])), HOROVOD_CROSS_SIZE=str(len(_uniq)))
    FAKE_HOSTS = len(_uniq)

hvd.init()
rank, size = hvd.rank(), hvd.size()
gen = torch.Generator().manual_seed(7)
if FAKE_HOSTS > 1:
    assert 'two-level' in hvd.control_plane_info(), hvd.control_plane_info()


def data(n, dtype, r):
    g = torch.Generator().manual_seed(1000 * r + n % 97)
    if dtype.is_floating_point:
        return (torch.rand(n, generator=g, dtype=torch.float32) * 2 - 1).to(dtype)
    if dtype == torch.bool:
        return torch.rand(n, generator=g) > 0.5
    return torch.randint(-3 if dtype != torch.uint8 else 0, 4, (n,), generator=g).to(dtype)


checked = 0
for n in (1, 5, 1023, 1024, 1025, 4099, 70001):
    for dtype in (torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8, torch.int8, torch.int16):
        xs = [data(n, dtype, r) for r in range(size)]
        for op, fn in ((hvd.Sum, lambda a, b: a + b), (hvd.Min, torch.minimum), (hvd.Max, torch.maximum), (hvd.Product, lambda a, b: a * b)):
            if op == hvd.Product and not dtype.is_floating_point:
                continue
            out = hvd.allreduce(xs[rank], op=op, name='a.%d.%s.%d' % (n, dtype, op))
            # reference in the order the data plane reduces a chunk: owner first, then (owner+1), ...; for 16-bit floats every
            # step rounds, so compare against a float32 accumulation with a tolerance instead
            if dtype in (torch.float16, torch.bfloat16):
                ref = xs[0].float()
                for r in range(1, size):
                    ref = fn(ref, xs[r].float())
                assert torch.allclose(out.float(), ref, rtol=2e-2, atol=2e-2), (n, dtype, op)
            else:
                ref = xs[0].clone()
                for r in range(1, size):
                    ref = fn(ref, xs[r])
                if dtype.is_floating_point:
                    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6), (n, dtype, op)
                else:
                    assert torch.equal(out, ref), (n, dtype, op)
            checked += 1
        avg = hvd.allreduce(xs[rank].float(), op=hvd.Average, name='avg.%d.%s' % (n, dtype))
        assert torch.allclose(avg, sum(x.float() for x in xs) / size, rtol=1e-5, atol=1e-5)

# big messages (several MiB): the helper-thread team splits the copy / reduce phases when HVD_CPU_THREADS > 1
if int(os.environ.get('HVD_CPU_THREADS', '1')) != 1:
    for n, dtype in ((3000001, torch.float32), (2500003, torch.bfloat16), (1500001, torch.int64)):
        xs = [data(n, dtype, r) for r in range(size)]
        for op, fn in ((hvd.Sum, lambda a, b: a + b), (hvd.Max, torch.maximum)):
            out = hvd.allreduce(xs[rank], op=op, name='big.%s.%d' % (dtype, op))
            ref = xs[0].float() if dtype == torch.bfloat16 else xs[0].clone()
            for r in range(1, size):
                ref = fn(ref, xs[r].float() if dtype == torch.bfloat16 else xs[r])
            if dtype == torch.bfloat16:
                assert torch.allclose(out.float(), ref, rtol=2e-2, atol=2e-2), (n, dtype, op)
            elif dtype.is_floating_point:
                assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6), (n, dtype, op)
            else:
                assert torch.equal(out, ref), (n, dtype, op)
            checked += 1

# uneven allgather (rank r contributes (r+1)*k rows), including an empty contribution
for k in (0, 1, 333, 5000):
    mine = torch.arange((rank + 1) * k * 3, dtype=torch.float32).reshape(-1, 3) + 1000 * rank
    out = hvd.allgather(mine, name='ag.%d' % k)
    ref = torch.cat([torch.arange((r + 1) * k * 3, dtype=torch.float32).reshape(-1, 3) + 1000 * r for r in range(size)])
    assert torch.equal(out, ref), k

# broadcast from every root, sizes around the slot size
for root in range(size):
    for n in (1, 1024, 1025, 50000):
        t = torch.full((n,), float(rank + 1), dtype=torch.float64)
        hvd.broadcast_(t, root_rank=root, name='bc.%d.%d' % (root, n))
        assert torch.all(t == root + 1)

# reducescatter, even and uneven first dimension
for rows in (size, size * 7 + 1, 4000 + size - 1):
    x = torch.arange(rows * 5, dtype=torch.float32).reshape(rows, 5) * (rank + 1)
    out = hvd.reducescatter(x, op=hvd.Sum, name='rs.%d' % rows)
    full = torch.arange(rows * 5, dtype=torch.float32).reshape(rows, 5) * (size * (size + 1) // 2)
    base, extra = divmod(rows, size)
    starts = [r * base + min(r, extra) for r in range(size + 1)]
    assert torch.equal(out, full[starts[rank]:starts[rank + 1]]), rows

# alltoall: even splits, uneven splits (rank r sends (r + q + 1) * k rows to rank q), empty blocks, more than one piece
for k in (1, 7, 400):
    splits = [(rank + q + 1) * k if (rank + q) % 3 else 0 for q in range(size)]
    rows = sum(splits)
    x = torch.arange(rows * 2, dtype=torch.float32).reshape(rows, 2) + 10000 * rank
    out, rsplits = hvd.alltoall(x, splits=torch.tensor(splits), name='a2a.%d' % k)
    expect, exp_splits = [], []
    for q in range(size):                       # what rank q sends to me
        qs = [(q + d + 1) * k if (q + d) % 3 else 0 for d in range(size)]
        start = sum(qs[:rank])
        full = torch.arange(sum(qs) * 2, dtype=torch.float32).reshape(sum(qs), 2) + 10000 * q
        expect.append(full[start:start + qs[rank]])
        exp_splits.append(qs[rank])
    assert rsplits.tolist() == exp_splits, (rsplits, exp_splits)
    assert torch.equal(out, torch.cat(expect)), k
even = hvd.alltoall(torch.arange(size * 3, dtype=torch.int64) + 100 * rank, name='a2a.even')
assert even.tolist() == [100 * q + 3 * rank + j for q in range(size) for j in range(3)]

# long mixed async sequence with fusion
handles = []
for i in range(40):
    t = torch.full((257 * (i % 5 + 1),), float(rank + i))
    handles.append((i, t.numel(), hvd.allreduce_async(t, op=hvd.Sum, name='mix.%d' % i)))
    if i % 3 == 0:
        b = torch.full((100 + i,), float(rank))
        hvd.broadcast_(b, root_rank=i % size, name='mixb.%d' % i)
        assert torch.all(b == i % size)
for i, n, h in handles:
    out = hvd.synchronize(h)
    assert torch.all(out == sum(r + i for r in range(size))) and out.numel() == n

# process sets have their own control channel and their own data slots
if size >= 3 and FAKE_HOSTS <= 1:
    evens = hvd.add_process_set([q for q in range(size) if q % 2 == 0])
    odds = hvd.add_process_set([q for q in range(size) if q % 2 == 1])
    mine = evens if rank % 2 == 0 else odds
    members = mine.ranks
    desc = hvd.control_plane_info(mine)
    if os.environ.get('HVD_CONTROL_PLANE') != 'tcp':
        assert 'shared memory channel' in desc, desc
        if os.environ.get('HVD_SHM_DATA_PLANE', '1') != '0' and len(members) > 1:
            assert 'shared-memory slots' in desc, desc
    for n in (3, 5000, 70001):
        x = torch.full((n,), float(rank + 1), dtype=torch.float64)
        out = hvd.allreduce(x, op=hvd.Sum, process_set=mine, name='ps.ar.%d' % n)
        assert torch.all(out == sum(q + 1 for q in members)), (n, out[:3])
        g = hvd.allgather(torch.full((rank + 1, 2), float(rank)), process_set=mine, name='ps.ag.%d' % n)
        assert g.shape[0] == sum(q + 1 for q in members) and g[0, 0] == members[0] and g[-1, 0] == members[-1]
        b = torch.full((n,), float(rank))
        hvd.broadcast_(b, root_rank=members[-1], process_set=mine, name='ps.bc.%d' % n)
        assert torch.all(b == members[-1])
        # interleave with the global set: the two planes keep separate piece counters and slots
        tot = hvd.allreduce(torch.full((n,), 1.0), op=hvd.Sum, name='ps.global.%d' % n)
        assert torch.all(tot == size)
    hvd.remove_process_set(odds)
    hvd.remove_process_set(evens)

paths = hvd.metrics()['host_paths']
if 'two-level (shared-memory slots' in hvd.control_plane_info():
    assert paths['two_level'] > 0 and paths['shared_memory'] == 0, paths        # allreduce / allgather / broadcast took the two-level path
elif 'host data: shared-memory slots' in hvd.control_plane_info():
    assert paths['shared_memory'] > 0 and paths['two_level'] == 0, paths
info = os.environ.get('HVD_SHM_DATA_PLANE', '1')
hvd.barrier()
if rank == 0:
    print('SHM PLANE OK', checked, 'plane=' + info, '|', hvd.control_plane_info())
hvd.shutdown()
