"""A peer that is alive but stopped (SIGSTOP) makes the others fail with an error after HVD_SHM_TIMEOUT_SECONDS instead of
spinning forever."""
import os
import signal
import sys
import time

import torch

import horovod_b200.torch as hvd

hvd.init()
r = hvd.rank()
pids = hvd.allgather_object(os.getpid())
hvd.allreduce(torch.ones(4), name='warm')
if r == 1:
    os.kill(os.getpid(), signal.SIGSTOP)         # every thread of this rank freezes, including the cycle thread
    time.sleep(60)
    sys.exit(0)
t0 = time.time()
try:
    for i in range(10000):
        hvd.allreduce(torch.ones(4), name='after.%d' % i)
        time.sleep(0.01)
    print('NO ERROR', flush=True)
except hvd.HorovodInternalError as e:
    dt = time.time() - t0
    print('TIMEOUT RAISED after %.1f s: %s' % (dt, str(e)[:120]), flush=True)
finally:
    os.kill(pids[1], signal.SIGKILL)             # a stopped process ignores SIGTERM until it is continued
sys.exit(0)
