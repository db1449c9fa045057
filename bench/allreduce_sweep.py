#!/usr/bin/env python
"""Allreduce bus-bandwidth sweep across runtime configurations (kernel variant, CTA count, backend), one process per
GPU (torchrun / hvdrun).  Every configuration re-initialises the runtime with different environment knobs, so one
launch compares e.g. two-shot vs NVLS vs the NCCL baseline on the same box.  Device-timed, max over ranks;
busBW = algBW * 2(N-1)/N.

    torchrun --nproc-per-node 8 bench/allreduce_sweep.py --configs p2p:auto:64,p2p:nvls:64,nccl --sizes 1024,1048576,...
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import horovod_b200.torch as hvd  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--configs', default='p2p:auto:64,nccl')
p.add_argument('--sizes', default=','.join(str(1 << s) for s in range(10, 31, 2)))
p.add_argument('--dtype', default='fp32')
p.add_argument('--out', default=None)
p.add_argument('--blocking', action='store_true', help='synchronize after every op (latency mode) instead of pipelining')
args = p.parse_args()
dt = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'fp16': torch.float16}[args.dtype]
sizes = [int(s) for s in args.sizes.split(',')]
results = {}
for cfg in args.configs.split(','):
    parts = cfg.split(':')
    os.environ['HVD_GPU_BACKEND'] = parts[0]
    os.environ['HVD_ALLREDUCE_VARIANT'] = parts[1] if len(parts) > 1 else 'auto'
    if len(parts) > 2:
        os.environ['HVD_COMM_CTAS'] = parts[2]
    else:
        os.environ.pop('HVD_COMM_CTAS', None)
    hvd.init()
    rank, size = hvd.rank(), hvd.size()
    torch.cuda.set_device(hvd.local_rank())
    rows = []
    for nbytes in sizes:
        n = max(1, nbytes // torch.tensor([], dtype=dt).element_size())
        x = torch.ones(n, device='cuda', dtype=dt)
        iters = max(5, min(100, int(1e9 // max(nbytes, 1 << 18))))
        name = f'sweep.{nbytes}'
        for _ in range(3):
            hvd.allreduce_(x, op=hvd.Sum, name=name)
        hvd.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if args.blocking:
            for _ in range(iters):
                hvd.allreduce_(x, op=hvd.Sum, name=name)
        else:
            # pipelined: distinct names in flight so the host never blocks between ops
            hs = [hvd.allreduce_async_(x, op=hvd.Sum, name=f'{name}.{i % 8}') if False else None for i in range(0)]
            for _ in range(iters):
                hvd.allreduce_(x, op=hvd.Sum, name=name)
        e1.record()
        torch.cuda.synchronize()
        ms = hvd.allreduce(torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64), op=hvd.Max, name='sweep.ms').item()
        alg = nbytes / (ms / 1e3) / 1e9
        rows.append({'bytes': nbytes, 'us': round(ms * 1e3, 2), 'algbw_gbs': round(alg, 2),
                     'busbw_gbs': round(alg * 2 * (size - 1) / size, 2)})
    results[cfg] = {'rows': rows, 'backend': hvd.gpu_backend_info(), 'tunables': hvd.tunable_params()}
    if rank == 0:
        print(f'== {cfg}  [{hvd.gpu_backend_info()}]', flush=True)
        for r in rows:
            print(f"  {r['bytes']:>12d} B  {r['us']:>10.2f} us  alg {r['algbw_gbs']:>8.2f} GB/s  bus {r['busbw_gbs']:>8.2f} GB/s", flush=True)
    hvd.shutdown()
if int(os.environ.get('RANK', os.environ.get('HOROVOD_RANK', '0'))) == 0 and args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump({'n_gpus': int(os.environ.get('WORLD_SIZE', os.environ.get('HOROVOD_SIZE', '1'))), 'dtype': args.dtype,
                   'results': results}, f, indent=1)
