import time, torch, horovod_b200.torch as hvd
hvd.init()
r = hvd.rank()
hvd.allreduce(torch.ones(2))
if r == 1:
    hvd.shutdown()          # job-wide: rank 0's loop ends too
    print('rank1 done')
else:
    time.sleep(1.0)
    assert hvd.rank() == 0 and hvd.size() == 2 and hvd.is_initialized()   # queries stay valid until the local shutdown
    try:
        hvd.allreduce(torch.ones(2))
        raise AssertionError('collective after peer shutdown must fail')
    except Exception as e:
        assert 'shut' in str(e).lower(), e
    hvd.shutdown()
    assert not hvd.is_initialized()
    print('rank0 done')
