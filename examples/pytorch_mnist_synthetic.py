"""The canonical user program of docs/pytorch.rst (init, pin GPU, scale LR, DistributedOptimizer, broadcast state,
train, metric averaging) on synthetic MNIST-shaped data:

    hvdrun -np 2 python examples/pytorch_mnist_synthetic.py --epochs 2
"""
import argparse

import torch
import torch.nn as nn
import torch.nn.functional as F

import horovod_b200.torch as hvd
from horovod_b200.torch.callbacks import MetricAverage

p = argparse.ArgumentParser()
p.add_argument('--epochs', type=int, default=2)
p.add_argument('--batch-size', type=int, default=64)
p.add_argument('--lr', type=float, default=0.01)
p.add_argument('--no-cuda', action='store_true')
args = p.parse_args()

hvd.init()
cuda = not args.no_cuda and torch.cuda.is_available()
if cuda:
    torch.cuda.set_device(hvd.local_rank())
dev = torch.device('cuda' if cuda else 'cpu')
torch.manual_seed(42)


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 10, kernel_size=5)
        self.conv2 = nn.Conv2d(10, 20, kernel_size=5)
        self.fc1 = nn.Linear(320, 50)
        self.fc2 = nn.Linear(50, 10)

    def forward(self, x):
        x = F.relu(F.max_pool2d(self.conv1(x), 2))
        x = F.relu(F.max_pool2d(self.conv2(x), 2))
        x = F.relu(self.fc1(x.view(-1, 320)))
        return self.fc2(x)


model = Net().to(dev)
# scale the learning rate by the number of workers, then wrap the optimizer
optimizer = torch.optim.SGD(model.parameters(), lr=args.lr * hvd.size(), momentum=0.5)
optimizer = hvd.DistributedOptimizer(optimizer, named_parameters=model.named_parameters(), op=hvd.Average)
hvd.broadcast_parameters(model.state_dict(), root_rank=0)
hvd.broadcast_optimizer_state(optimizer, root_rank=0)

# every rank draws its own shard of a synthetic data set whose label is a function of the image
g = torch.Generator().manual_seed(1000 + hvd.rank())
images = torch.randn(512, 1, 28, 28, generator=g)
labels = (images.mean(dim=(1, 2, 3)) > 0).long() + 2 * (images[:, 0, 0, 0] > 0).long()
average = MetricAverage()
for epoch in range(args.epochs):
    model.train()
    total, correct, loss_sum = 0, 0, 0.0
    for i in range(0, len(images), args.batch_size):
        x, y = images[i:i + args.batch_size].to(dev), labels[i:i + args.batch_size].to(dev)
        optimizer.zero_grad()
        out = model(x)
        loss = F.cross_entropy(out, y)
        loss.backward()
        optimizer.step()
        loss_sum += loss.item() * len(x)
        correct += (out.argmax(1) == y).sum().item()
        total += len(x)
    m = average({'loss': loss_sum / total, 'accuracy': correct / total})
    if hvd.rank() == 0:
        print('epoch %d: loss %.4f accuracy %.3f (averaged over %d ranks)' % (epoch, m['loss'], m['accuracy'], hvd.size()))
hvd.shutdown()
