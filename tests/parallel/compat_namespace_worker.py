"""The canonical program of the reference docs, verbatim imports: `import horovod.torch as hvd`."""
import torch, horovod.torch as hvd
import torch.nn.functional as F
hvd.init()
model = torch.nn.Linear(8, 4)
optimizer = torch.optim.SGD(model.parameters(), lr=0.01 * hvd.size())
optimizer = hvd.DistributedOptimizer(optimizer, named_parameters=model.named_parameters(), compression=hvd.Compression.none, op=hvd.Average)
hvd.broadcast_parameters(model.state_dict(), root_rank=0)
hvd.broadcast_optimizer_state(optimizer, root_rank=0)
for step in range(3):
    data, target = torch.randn(16, 8), torch.randint(0, 4, (16,))
    optimizer.zero_grad(); loss = F.cross_entropy(model(data), target); loss.backward(); optimizer.step()
import horovod, horovod_b200, horovod_b200.torch
assert horovod.torch is horovod_b200.torch
from horovod.runner.common.util import hosts
import horovod.torch.elastic as el
assert hosts.parse_hosts('a:2')[0].slots == 2 and el is horovod_b200.torch.elastic
from horovod.common.exceptions import HorovodInternalError
w = hvd.allgather(model.weight.detach().reshape(1, -1))
assert torch.allclose(w[0], w[-1])
if hvd.rank() == 0: print('COMPAT OK', horovod.__version__)
hvd.shutdown()
