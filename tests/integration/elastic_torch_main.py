"""Elastic training script used by the fault-injection tests (role of the reference's
test/integration/data/elastic_torch_main.py): trains a tiny model for a few epochs, logs (epoch, rank, size) per epoch to
a JSON-lines file and kills/raises on a schedule {"epoch,batch": [ranks]}."""
import argparse
import json
import os
import sys

import psutil
import torch

import horovod_b200.torch as hvd

p = argparse.ArgumentParser()
p.add_argument('--batches-per-epoch', type=int, default=6)
p.add_argument('--epochs', type=int, default=3)
p.add_argument('--logfile', required=True)
p.add_argument('--exit-schedule', default='{}')
p.add_argument('--exit-mode', default='exception', choices=['exception', 'kill'])
p.add_argument('--discovery-schedule-epoch-file', default=None)
p.add_argument('--batch-sleep', type=float, default=0.0)
p.add_argument('--device', default='cpu', choices=['cpu', 'cuda'])
args = p.parse_args()
schedule = {tuple(int(x) for x in k.split(',')): v for k, v in json.loads(args.exit_schedule).items()}

hvd.init()
torch.manual_seed(1234)
start_rank = int(os.environ.get('HOROVOD_RANK', 0))
DEV = torch.device('cpu')
if args.device == 'cuda':
    # every worker keeps the GPU of its ORIGINAL rank for its whole life (ranks are renumbered after a reset, and the two
    # launcher "hosts" of the fault-injection tests both start their local ranks at 0)
    DEV = torch.device('cuda', start_rank % torch.cuda.device_count())
    torch.cuda.set_device(DEV)
model = torch.nn.Linear(4, 1).to(DEV)
optimizer = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.01), named_parameters=model.named_parameters())
hostname = os.environ.get('HOROVOD_HOSTNAME')


def check_exit(epoch, batch):
    ranks = schedule.get((epoch, batch))
    # only the ORIGINAL incarnation of a rank dies (a respawned worker with the same rank must not loop forever)
    if ranks and start_rank in ranks and not os.path.exists(args.logfile + f'.died.{start_rank}.{epoch}.{batch}'):
        open(args.logfile + f'.died.{start_rank}.{epoch}.{batch}', 'w').close()
        if args.exit_mode == 'exception':
            raise RuntimeError(f'scheduled failure of rank {start_rank} at epoch {epoch} batch {batch}')
        psutil.Process(os.getpid()).kill()


def log_state(state):
    rec = {'epoch': state.epoch, 'commits': state.commits, 'rank': hvd.rank(), 'size': hvd.size(), 'start_rank': start_rank,
           'hostname': hostname}
    with open(args.logfile, 'a') as f:
        f.write(json.dumps(rec) + os.linesep)
    if args.discovery_schedule_epoch_file and hvd.rank() == 0:
        with open(args.discovery_schedule_epoch_file, 'w') as f:
            f.write(str(state.epoch))


@hvd.elastic.run
def train(state):
    state.rendezvous += 1
    while state.epoch < args.epochs:
        while state.batch < args.batches_per_epoch:
            check_exit(state.epoch, state.batch)
            if args.batch_sleep:
                import time
                time.sleep(args.batch_sleep)
            optimizer.zero_grad()
            loss = model(torch.ones(2, 4, device=DEV) * (hvd.rank() + 1)).pow(2).mean()
            loss.backward()
            optimizer.step()
            state.batch += 1
            if state.batch % 2 == 0:
                state.commits += 1
                state.commit()
        log_state(state)
        state.epoch += 1
        state.batch = 0
        state.commits += 1
        state.commit()


def on_reset():
    for g in optimizer.param_groups:
        g['lr'] = 0.01 * hvd.size()


state = hvd.elastic.TorchState(model, optimizer, batch=0, epoch=0, commits=0, rendezvous=0)
state.register_reset_callbacks([on_reset])
train(state)
# all ranks must agree on the final model
w = hvd.allgather(model.weight.detach().reshape(1, -1).contiguous())
assert all(torch.allclose(w[r], w[0]) for r in range(hvd.size())), w
if hvd.rank() == 0:
    with open(args.logfile, 'a') as f:
        f.write(json.dumps({'done': True, 'size': hvd.size(), 'rendezvous': state.rendezvous}) + os.linesep)
hvd.shutdown()
