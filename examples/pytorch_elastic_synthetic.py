"""Elastic training skeleton (reference docs/elastic.rst):

    hvdrun -np 2 --min-np 1 --max-np 4 --host-discovery-script ./discover_hosts.sh python examples/pytorch_elastic_synthetic.py
"""
import torch
import torch.nn.functional as F

import horovod_b200.torch as hvd

hvd.init()
if torch.cuda.is_available():
    torch.cuda.set_device(hvd.local_rank())
dev = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
model = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 4)).to(dev)
base_lr = 0.01
optimizer = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=base_lr * hvd.size()),
                                     named_parameters=model.named_parameters())


@hvd.elastic.run
def train(state):
    for state.epoch in range(state.epoch, 5):
        for state.batch in range(state.batch, 50):
            x = torch.randn(16, 32, device=dev)
            y = torch.randint(0, 4, (16,), device=dev)
            optimizer.zero_grad()
            F.cross_entropy(model(x), y).backward()
            optimizer.step()
            if state.batch % 10 == 0:
                state.commit()       # snapshot + check for host changes
        state.batch = 0
        if hvd.rank() == 0:
            print('epoch', state.epoch, 'done on', hvd.size(), 'ranks')


def on_state_reset():
    for g in optimizer.param_groups:   # rescale the LR to the new world size
        g['lr'] = base_lr * hvd.size()


state = hvd.elastic.TorchState(model, optimizer, batch=0, epoch=0)
state.register_reset_callbacks([on_state_reset])
train(state)
hvd.shutdown()
