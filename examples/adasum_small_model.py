"""Adasum vs averaging on a small regression model.

    hvdrun -np 4 python examples/adasum_small_model.py [--no-cuda]

With op=hvd.Average the learning rate has to grow with the number of ranks to keep up; with op=hvd.Adasum the combined update
adapts by itself: orthogonal gradients add, parallel gradients average, so the SAME base learning rate works at any
scale.  The script trains the same model both ways from the same initial weights and prints both loss curves.
"""
import argparse

import torch

import horovod_b200.torch as hvd


def make_problem(seed, device):
    g = torch.Generator().manual_seed(seed)
    w_true = torch.randn(16, 1, generator=g)
    x = torch.randn(2048, 16, generator=g)
    y = x @ w_true + 0.01 * torch.randn(2048, 1, generator=g)
    return x.to(device), y.to(device)


def train(op, lr, steps, device, batch):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(device)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=lr), named_parameters=model.named_parameters(), op=op)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    x, y = make_problem(1, device)
    n = x.shape[0] // hvd.size()
    x, y = x[hvd.rank() * n:(hvd.rank() + 1) * n], y[hvd.rank() * n:(hvd.rank() + 1) * n]     # each rank sees its own shard
    curve = []
    for step in range(steps):
        idx = torch.randint(0, n, (batch,), generator=torch.Generator().manual_seed(100 * step + hvd.rank()))
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(x[idx]), y[idx])
        loss.backward()
        opt.step()
        if step % max(1, steps // 5) == 0 or step == steps - 1:
            curve.append(hvd.allreduce(loss.detach(), name='loss.%s.%d' % (op, step)).item())
    return curve


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    p.add_argument('--steps', type=int, default=200)
    p.add_argument('--batch-size', type=int, default=32)
    p.add_argument('--lr', type=float, default=0.05)
    p.add_argument('--no-cuda', action='store_true')
    a = p.parse_args()
    hvd.init()
    use_cuda = torch.cuda.is_available() and not a.no_cuda
    if use_cuda:
        torch.cuda.set_device(hvd.local_rank())
    device = torch.device('cuda', hvd.local_rank()) if use_cuda else torch.device('cpu')
    avg = train(hvd.Average, a.lr, a.steps, device, a.batch_size)
    ada = train(hvd.Adasum, a.lr, a.steps, device, a.batch_size)
    if hvd.rank() == 0:
        print('ranks: %d, base lr %.3g' % (hvd.size(), a.lr))
        print('Average:', ' '.join('%.4f' % v for v in avg))
        print('Adasum :', ' '.join('%.4f' % v for v in ada))
        print('ADASUM EXAMPLE OK' if ada[-1] < ada[0] and avg[-1] < avg[0] else 'did not converge')
    hvd.shutdown()
