"""Auxiliary subsystems end to end: stall inspector (warning + shutdown), autotuner log, timeline file."""
import json
import os
import sys
import time

import torch

import horovod_b200.torch as hvd
from horovod_b200.common.exceptions import HorovodInternalError

mode = sys.argv[1]
hvd.init()
r, n = hvd.rank(), hvd.size()
if mode == 'stall_warning':
    # rank 1 submits late: rank 0 (coordinator) must warn and name the missing rank, then the op completes
    if r != 0:
        time.sleep(3.5)
    out = hvd.allreduce(torch.ones(2), op=hvd.Sum, name='late.tensor')
    assert out.tolist() == [float(n)] * 2
    print('STALL WARNING DONE', r, flush=True)
elif mode == 'stall_cached':
    # the tensor is in the response cache (fast path, no coordinator involved); a late rank must still be reported
    for _ in range(4):
        hvd.allreduce(torch.ones(2), op=hvd.Sum, name='cached.tensor')
    if r != 0:
        time.sleep(3.5)
    out = hvd.allreduce(torch.ones(2), op=hvd.Sum, name='cached.tensor')
    assert out.tolist() == [float(n)] * 2
    print('STALL CACHED DONE', r, flush=True)
elif mode == 'stall_shutdown':
    if r == 0:
        try:
            hvd.allreduce(torch.ones(2), op=hvd.Sum, name='never.matched')
            raise AssertionError('must fail')
        except HorovodInternalError as e:
            print('STALL SHUTDOWN RAISED', str(e)[:80].replace('\n', ' '), flush=True)
    else:
        time.sleep(6)
        try:
            hvd.allreduce(torch.ones(2), op=hvd.Sum, name='other.name')
        except HorovodInternalError:
            pass
        print('PEER SAW SHUTDOWN', flush=True)
elif mode == 'autotune':
    x = torch.ones(1 << 14)
    for step in range(400):
        hs = [hvd.allreduce_async(x, op=hvd.Sum, name=f'at.{i}') for i in range(4)]
        for h in hs:
            hvd.synchronize(h)
    sys.stdout.write('AUTOTUNE PARAMS ' + json.dumps(hvd.tunable_params()) + '\n')  # one write: lines of two ranks must not interleave
    sys.stdout.flush()
elif mode == 'timeline':
    path = sys.argv[2]
    hvd.start_timeline(path, mark_cycles=True)
    for i in range(5):
        hvd.allreduce(torch.ones(64), name='tl.ar')
        hvd.allgather(torch.ones(2, 2), name='tl.ag')
    hvd.barrier()
    if hvd.rank() == 0:
        # the file must be a complete JSON document WHILE the timeline is still running (closing bracket rewritten in place
        # after every drain of the writer's ring), not only after stop_timeline()
        # (the writer keeps appending cycle markers, and a read that straddles one of its updates sees a torn file — old
        # tail plus new tail — so the check is what a person reloading the trace viewer does: read again)
        live, last_error = None, None
        for _ in range(40):
            time.sleep(0.05)
            fd = os.open(path, os.O_RDONLY)
            try:
                data = os.read(fd, os.fstat(fd).st_size)         # one read call of the size the file has right now
            finally:
                os.close(fd)
            try:
                live = json.loads(data)
                break
            except ValueError as e:
                last_error = e
        assert live is not None, last_error
        assert isinstance(live, list) and len(live) > 10, len(live)
        print('TIMELINE LIVE JSON OK', len(live), flush=True)
    hvd.barrier()
    hvd.stop_timeline()
    hvd.barrier()
    print('TIMELINE DONE', flush=True)
hvd.shutdown()
