"""Synthetic benchmark with the command line of the reference's examples/pytorch/pytorch_synthetic_benchmark.py:

    hvdrun -np 8 python examples/pytorch_synthetic_benchmark.py --model resnet50 --batch-size 64
"""
import argparse
import timeit

import numpy as np
import torch
import torch.nn.functional as F
import torch.optim as optim

import horovod_b200.torch as hvd
from horovod_b200 import models

parser = argparse.ArgumentParser(description='PyTorch Synthetic Benchmark', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
parser.add_argument('--fp16-allreduce', action='store_true', default=False, help='use fp16 compression during allreduce')
parser.add_argument('--bf16-wire', action='store_true', default=False, help='bf16 on the wire with the cast fused into the kernel')
parser.add_argument('--model', type=str, default='resnet50', help='model to benchmark (resnet50 | resnet101)')
parser.add_argument('--batch-size', type=int, default=32, help='input batch size')
parser.add_argument('--num-warmup-batches', type=int, default=10)
parser.add_argument('--num-batches-per-iter', type=int, default=10)
parser.add_argument('--num-iters', type=int, default=10)
parser.add_argument('--no-cuda', action='store_true', default=False)
parser.add_argument('--use-adasum', action='store_true', default=False, help='use adasum algorithm to do reduction')
args = parser.parse_args()
args.cuda = not args.no_cuda and torch.cuda.is_available()

hvd.init()
if args.cuda:
    torch.cuda.set_device(hvd.local_rank())
torch.backends.cudnn.benchmark = True

model = getattr(models, args.model)()
# By default, Adasum doesn't need scaling up learning rate.
lr_scaler = hvd.size() if not args.use_adasum else 1
if args.cuda:
    model.cuda()
    # If using GPU Adasum allreduce, scale learning rate by local_size.
    if args.use_adasum and hvd.nccl_built():
        lr_scaler = hvd.local_size()
optimizer = optim.SGD(model.parameters(), lr=0.01 * lr_scaler)
compression = hvd.Compression.fp16 if args.fp16_allreduce else hvd.Compression.none
optimizer = hvd.DistributedOptimizer(optimizer, named_parameters=model.named_parameters(), compression=compression,
                                     op=hvd.Adasum if args.use_adasum else hvd.Average)
hvd.broadcast_parameters(model.state_dict(), root_rank=0)
hvd.broadcast_optimizer_state(optimizer, root_rank=0)

data = torch.randn(args.batch_size, 3, 224, 224)
target = torch.LongTensor(args.batch_size).random_() % 1000
if args.cuda:
    data, target = data.cuda(), target.cuda()


def benchmark_step():
    optimizer.zero_grad()
    output = model(data)
    loss = F.cross_entropy(output, target)
    loss.backward()
    optimizer.step()


def log(s, nl=True):
    if hvd.rank() != 0:
        return
    print(s, end='\n' if nl else '')


log('Model: %s' % args.model)
log('Batch size: %d' % args.batch_size)
device = 'GPU' if args.cuda else 'CPU'
log('Number of %ss: %d' % (device, hvd.size()))
log('Running warmup...')
timeit.timeit(benchmark_step, number=args.num_warmup_batches)
log('Running benchmark...')
img_secs = []
for x in range(args.num_iters):
    if args.cuda:
        torch.cuda.synchronize()
    time = timeit.timeit(lambda: (benchmark_step(), torch.cuda.synchronize() if args.cuda else None), number=args.num_batches_per_iter)
    img_sec = args.batch_size * args.num_batches_per_iter / time
    log('Iter #%d: %.1f img/sec per %s' % (x, img_sec, device))
    img_secs.append(img_sec)
img_sec_mean = np.mean(img_secs)
img_sec_conf = 1.96 * np.std(img_secs)
log('Img/sec per %s: %.1f +-%.1f' % (device, img_sec_mean, img_sec_conf))
log('Total img/sec on %d %s(s): %.1f +-%.1f' % (hvd.size(), device, hvd.size() * img_sec_mean, hvd.size() * img_sec_conf))
