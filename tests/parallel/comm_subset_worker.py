import os, torch
import horovod_b200.torch as hvd
g = int(os.environ['HOROVOD_RANK'])
if g in (0, 2):
    hvd.init(comm=[2, 0])          # list order defines the new ranks: global 2 -> rank 0, global 0 -> rank 1
    assert hvd.size() == 2 and hvd.rank() == (0 if g == 2 else 1), (g, hvd.rank())
    assert hvd.local_size() == 2 and hvd.cross_size() == 1
    out = hvd.allreduce(torch.ones(3) * (g + 1), op=hvd.Sum)
    assert out.tolist() == [4.0] * 3
    b = hvd.broadcast(torch.tensor([float(g)]), root_rank=0)
    assert b.item() == 2.0
else:
    try:
        hvd.init(comm=[0, 2])
        raise AssertionError('non-member accepted')
    except ValueError:
        pass
    hvd.init(comm=[1])
    assert hvd.size() == 1 and hvd.rank() == 0
    assert hvd.allreduce(torch.ones(2), op=hvd.Sum).tolist() == [1.0, 1.0]
hvd.shutdown()
print('COMM SUBSET OK', g)
