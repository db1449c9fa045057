"""hvd.init(process_sets=[...]): statically registered sets (reference test/parallel/test_process_sets_static.py and
test_process_sets_multi_comm.py)."""
import torch

import horovod_b200.torch as hvd

import os
r_env, n_env = int(os.environ['HOROVOD_RANK']), int(os.environ['HOROVOD_SIZE'])
even = hvd.ProcessSet([q for q in range(n_env) if q % 2 == 0])
odd = hvd.ProcessSet([q for q in range(n_env) if q % 2 == 1])
hvd.init(process_sets=[even, odd])
r, n = hvd.rank(), hvd.size()
assert even.process_set_id == 1 and odd.process_set_id == 2
mine, other = (even, odd) if r % 2 == 0 else (odd, even)
assert mine.included() and not other.included()
assert mine.size() == len(mine.ranks) and mine.rank() == mine.ranks.index(r)
assert hvd.global_process_set.size() == n
out = hvd.allreduce(torch.ones(3) * (r + 1), op=hvd.Sum, process_set=mine, name='static.sum')
assert out.tolist() == [float(sum(q + 1 for q in mine.ranks))] * 3
g = hvd.allgather(torch.tensor([float(r)]), process_set=mine, name='static.ag')
assert g.tolist() == [float(q) for q in mine.ranks]
b = hvd.broadcast(torch.tensor([float(r)]), root_rank=mine.ranks[-1], process_set=mine, name='static.bc')
assert b.item() == float(mine.ranks[-1])
# on one host the sets negotiate through their own shared-memory channel (unless the test forces TCP)
if os.environ.get('HVD_CONTROL_PLANE') != 'tcp' and n > 2:
    assert 'shared memory channel' in hvd.control_plane_info(mine), hvd.control_plane_info(mine)
    assert 'not a member' in hvd.control_plane_info(other)
try:
    hvd.allreduce(torch.ones(1), process_set=other, name='static.notmember')
    raise AssertionError('a non-member must not be able to use the set')
except (ValueError, hvd.HorovodInternalError):
    pass
# a set with the same ranks cannot be added twice; a new one can, and ids are handed out in order
try:
    hvd.add_process_set(list(even.ranks))
    raise AssertionError('duplicate set accepted')
except ValueError:
    pass
if n > 2:
    most = hvd.add_process_set(list(range(n - 1)))
    assert most.process_set_id == 3
    hvd.remove_process_set(most)
# re-init with the same static sets keeps working
hvd.shutdown()
even2, odd2 = hvd.ProcessSet(list(even.ranks)), hvd.ProcessSet(list(odd.ranks))
hvd.init(process_sets=[even2, odd2])
mine2 = even2 if r % 2 == 0 else odd2
assert hvd.allreduce(torch.ones(1), op=hvd.Sum, process_set=mine2, name='static.again').item() == len(mine2.ranks)
hvd.barrier()
if r == 0:
    print('STATIC SETS OK')
hvd.shutdown()
