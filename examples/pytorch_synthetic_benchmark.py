"""Synthetic data-parallel throughput benchmark.

    hvdrun -np 8 python examples/pytorch_synthetic_benchmark.py --model resnet50 --batch-size 64

Accepts the flags of the reference example of the same name (--model, --batch-size, --num-warmup-batches,
--num-batches-per-iter, --num-iters, --fp16-allreduce, --use-adasum, --no-cuda) and prints the same kind of report, but is
organised as a small harness: rounds are timed with CUDA events on the device (the host clock only when running on CPU),
the per-round rate is the slowest rank's, and `--graphed` captures forward + backward in a CUDA graph (hvd.GraphedStep).
`bench.py` at the repository root is the full benchmark driver; this file is the minimal user-level program.
"""
import argparse
import statistics
import time

import torch

import horovod_b200.torch as hvd
from horovod_b200 import models


def cli():
    p = argparse.ArgumentParser(description=__doc__.splitlines()[0], formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--model', default='resnet50', choices=['resnet50', 'resnet101'])
    p.add_argument('--batch-size', type=int, default=32, help='images per step and process')
    p.add_argument('--num-warmup-batches', type=int, default=10)
    p.add_argument('--num-batches-per-iter', type=int, default=10)
    p.add_argument('--num-iters', type=int, default=10)
    p.add_argument('--fp16-allreduce', action='store_true', help='cast gradients to fp16 for the allreduce')
    p.add_argument('--bf16-wire', action='store_true', help='bf16 on the wire, cast fused into the allreduce kernel')
    p.add_argument('--use-adasum', action='store_true', help='combine gradients with Adasum instead of averaging')
    p.add_argument('--graphed', action='store_true', help='capture forward + backward in a CUDA graph')
    p.add_argument('--no-cuda', action='store_true')
    return p.parse_args()


class Harness:
    def __init__(self, args):
        self.args = args
        self.on_gpu = torch.cuda.is_available() and not args.no_cuda
        self.device = torch.device('cuda', hvd.local_rank()) if self.on_gpu else torch.device('cpu')
        if self.on_gpu:
            torch.cuda.set_device(self.device)
            torch.backends.cudnn.benchmark = True
        self.model = getattr(models, args.model)().to(self.device)
        self.optimizer = self._optimizer()
        hvd.broadcast_parameters(self.model.state_dict(), root_rank=0)
        hvd.broadcast_optimizer_state(self.optimizer, root_rank=0)
        gen = torch.Generator().manual_seed(1234 + hvd.rank())
        self.images = torch.randn(args.batch_size, 3, 224, 224, generator=gen).to(self.device)
        self.labels = torch.randint(0, 1000, (args.batch_size,), generator=gen).to(self.device)
        self.step = self._graphed_step() if args.graphed and self.on_gpu else self._eager_step

    def _optimizer(self):
        a = self.args
        # averaging N gradients wants the learning rate scaled by N; Adasum adapts by itself (one node: by local size)
        scale = (hvd.local_size() if self.on_gpu and hvd.nccl_built() else 1) if a.use_adasum else hvd.size()
        base = torch.optim.SGD(self.model.parameters(), lr=0.01 * scale)
        if a.bf16_wire:
            import os
            os.environ.setdefault('HVD_WIRE_DTYPE', 'bf16')
        return hvd.DistributedOptimizer(base, named_parameters=self.model.named_parameters(),
                                        compression=hvd.Compression.fp16 if a.fp16_allreduce else hvd.Compression.none,
                                        op=hvd.Adasum if a.use_adasum else hvd.Average)

    def _loss(self):
        return torch.nn.functional.cross_entropy(self.model(self.images), self.labels)

    def _eager_step(self):
        self.optimizer.zero_grad()
        self._loss().backward()
        self.optimizer.step()

    def _graphed_step(self):
        graphed = hvd.GraphedStep(lambda x, y: torch.nn.functional.cross_entropy(self.model(x), y), self.optimizer,
                                  (self.images, self.labels))
        return lambda: graphed(self.images, self.labels)

    def timed_round(self, steps):
        """Seconds for `steps` steps on this rank."""
        if not self.on_gpu:
            t0 = time.perf_counter()
            for _ in range(steps):
                self.step()
            return time.perf_counter() - t0
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            self.step()
        stop.record()
        stop.synchronize()
        return start.elapsed_time(stop) / 1e3

    def run(self):
        a, say = self.args, (print if hvd.rank() == 0 else (lambda *x, **k: None))
        unit = 'GPU' if self.on_gpu else 'CPU'
        say('Model: %s\nBatch size: %d\nNumber of %ss: %d' % (a.model, a.batch_size, unit, hvd.size()))
        say('Running warmup...')
        self.timed_round(a.num_warmup_batches)
        say('Running benchmark...')
        rates = []
        for i in range(a.num_iters):
            seconds = self.timed_round(a.num_batches_per_iter)
            slowest = hvd.allreduce(torch.tensor([seconds], dtype=torch.float64), op=hvd.Max, name='round_seconds').item()
            rates.append(a.batch_size * a.num_batches_per_iter / slowest)
            say('Iter #%d: %.1f img/sec per %s' % (i, rates[-1], unit))
        mean = statistics.fmean(rates)
        spread = 1.96 * (statistics.pstdev(rates) if len(rates) > 1 else 0.0)
        say('Img/sec per %s: %.1f +-%.1f' % (unit, mean, spread))
        say('Total img/sec on %d %s(s): %.1f +-%.1f' % (hvd.size(), unit, hvd.size() * mean, hvd.size() * spread))


if __name__ == '__main__':
    hvd.init()
    Harness(cli()).run()
    hvd.shutdown()
