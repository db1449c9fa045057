"""Property-based tests (hypothesis) of the pure-Python planning logic: host layouts, shard ranges, elastic sampler."""
from unittest import mock

from hypothesis import given, settings, strategies as st

from horovod_b200.parallel.sharded import shard_range
from horovod_b200.runner.cluster_job import assign_slots
from horovod_b200.runner.common.util import hosts


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(min_value=1, max_value=8), min_size=1, max_size=6), st.data())
def test_host_assignments_invariants(slot_counts, data):
    infos = [hosts.HostInfo('h%d' % i, s) for i, s in enumerate(slot_counts)]
    total = sum(slot_counts)
    max_np = data.draw(st.integers(min_value=1, max_value=total))
    slots = hosts.get_host_assignments(infos, 1, max_np)
    assert [s.rank for s in slots] == list(range(max_np)) and all(s.size == max_np for s in slots)
    by_host = {}
    for s in slots:
        by_host.setdefault(s.hostname, []).append(s)
    for name, ss in by_host.items():
        assert [s.local_rank for s in ss] == list(range(len(ss))) and all(s.local_size == len(ss) for s in ss)
        assert [s.rank for s in ss] == list(range(ss[0].rank, ss[0].rank + len(ss)))       # contiguous per host
    for lr in {s.local_rank for s in slots}:
        same = [s for s in slots if s.local_rank == lr]
        assert [s.cross_rank for s in same] == list(range(len(same))) and all(s.cross_size == len(same) for s in same)
    # the scheduler-agnostic layout (Ray / Spark integrations) agrees with the launcher's
    node_ids = [s.hostname for s in slots]
    again = assign_slots(node_ids)
    assert [(a.rank, a.local_rank, a.cross_rank, a.local_size, a.cross_size) for a in again] == \
           [(s.rank, s.local_rank, s.cross_rank, s.local_size, s.cross_size) for s in slots]


@settings(max_examples=300, deadline=None)
@given(st.integers(min_value=0, max_value=10_000), st.integers(min_value=1, max_value=64))
def test_shard_ranges_tile_the_vector(numel, size):
    pieces = [shard_range(numel, r, size) for r in range(size)]
    assert pieces[0][0] == 0 and pieces[-1][1] == numel
    assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    lens = [hi - lo for lo, hi in pieces]
    assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=1, max_value=200), st.integers(min_value=1, max_value=8), st.booleans(), st.integers(0, 3), st.data())
def test_elastic_sampler_partitions_remaining_samples(n, world, shuffle, epoch, data):
    from horovod_b200.torch.elastic.sampler import ElasticSampler
    done = set(data.draw(st.lists(st.integers(min_value=0, max_value=n - 1), max_size=n, unique=True)))
    seen = []
    lengths = set()
    for rank in range(world):
        with mock.patch('horovod_b200.torch.elastic.sampler.size', return_value=world), \
                mock.patch('horovod_b200.torch.elastic.sampler.rank', return_value=rank):
            s = ElasticSampler(list(range(n)), shuffle=shuffle, seed=7)
            s.load_state_dict({'epoch': epoch, 'processed_indices': set(done)})
            idx = list(iter(s))
            lengths.add(len(idx))
            assert len(idx) == len(s)
            seen += idx
    remaining = set(range(n)) - done
    assert len(lengths) == 1                                   # every rank iterates the same number of samples
    assert set(seen) == remaining or (not remaining and not seen)
    assert len(seen) - len(remaining) < world                  # only the padding repeats
