#!/usr/bin/env python
"""One launch of EVERY kernel family with ONE simulated rank, as a target for ncu (a multi-rank rendezvous would deadlock
under ncu's kernel serialisation).  Used by bench/run_ncu_captures.sh:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/prof_<name> \\
        python bench/ncu_targets.py --only <family>

Families: allreduce (fused pack+reduce+unpack), inplace (zero-copy), exchange (allgather data mover; HVD_EXCHANGE_TMA=1
for the TMA variant), pipelined (HVD role-specialised kernel, P2P reduce stage), adasum (pack/dots/combine/gather are
skipped at one rank: only pack + gather run), optim (fused SGD / Adam).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from horovod_b200.ops import sim  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--only', default='allreduce,inplace,exchange,pipelined,adasum,optim')
p.add_argument('--mb', type=int, default=64)
args = p.parse_args()
want = set(args.only.split(','))
n_el = (args.mb << 20) // 4

if 'allreduce' in want:
    ins = [[torch.ones(n_el // 161, device='cuda') for _ in range(161)]]
    outs = [[torch.empty_like(t) for t in ins[0]]]
    sim.allreduce(ins, outs, variant=sim.ONESHOT, ctas=128)
    sim.allreduce(ins, outs, variant=sim.TWOSHOT, ctas=256)
if 'inplace' in want:
    sim.inplace_allreduce([torch.ones(n_el, device='cuda')], ctas=128)
if 'exchange' in want:
    src = [torch.ones(args.mb << 20, device='cuda', dtype=torch.uint8)]
    dst = [torch.empty(args.mb << 20, device='cuda', dtype=torch.uint8)]
    sim.allgather(src, dst, ctas=128)
if 'pipelined' in want:
    os.environ.setdefault('HVD_PIPE_CHUNK_BYTES', str(8 << 20))
    ins = [[torch.ones(n_el, device='cuda')]]
    sim.allreduce(ins, ins, variant=3, ctas=128)
if 'adasum' in want:
    ins = [[torch.ones(n_el // 4, device='cuda') for _ in range(4)]]
    outs = [[torch.empty_like(t) for t in ins[0]]]
    sim.adasum(ins, outs, ctas=64)
if 'optim' in want:
    import horovod_b200.ops as ops
    import horovod_b200.torch as hvd
    hvd.init()
    ps = [torch.randn(n_el // 64, device='cuda') for _ in range(64)]
    gs = [torch.randn_like(q) for q in ps]
    ms = [torch.zeros_like(q) for q in ps]
    vs = [torch.zeros_like(q) for q in ps]
    ops.fused_sgd_step(ps, gs, ms, lr=0.1, momentum=0.9, first_step=True)
    ops.fused_adam_step(ps, gs, ms, vs, lr=1e-3, step=1)
    hvd.shutdown()
torch.cuda.synchronize()
print('NCU TARGETS DONE', sorted(want))
