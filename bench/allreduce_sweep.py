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
p.add_argument('--symm', action='store_true', help='allocate the tensors in registered symmetric memory (zero-copy path)')
p.add_argument('--inflight', type=int, default=1, help='number of distinct tensors kept in flight (1 = blocking loop)')
args = p.parse_args()
dt = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'fp16': torch.float16}[args.dtype]
sizes = [int(s) for s in args.sizes.split(',')]
results = {}
_extra_env = set()
for cfg in args.configs.split(','):
    # "<backend>:<variant>:<ctas>[:...]+KEY=VALUE+KEY=VALUE": extra environment knobs for this configuration only
    for k in _extra_env:
        os.environ.pop(k, None)
    _extra_env = set()
    base, *kvs = cfg.split('+')
    for kv in kvs:
        k, v = kv.split('=', 1)
        os.environ[k] = v
        _extra_env.add(k)
    parts = base.split(':')
    os.environ['HVD_GPU_BACKEND'] = parts[0]
    os.environ['HVD_ALLREDUCE_VARIANT'] = parts[1] if len(parts) > 1 else 'auto'
    if len(parts) > 2:
        os.environ['HVD_COMM_CTAS'] = parts[2]
    else:
        os.environ.pop('HVD_COMM_CTAS', None)
    # optional 4th / 5th fields: multimem.ld_reduce per thread in flight (4|8), max chunk bytes of the zero-copy kernel
    os.environ['HVD_NVLS_UNROLL'] = parts[3] if len(parts) > 3 else '4'
    if len(parts) > 4:
        os.environ['HVD_INPLACE_CHUNK_BYTES'] = parts[4]
    else:
        os.environ.pop('HVD_INPLACE_CHUNK_BYTES', None)
    hvd.init()
    rank, size = hvd.rank(), hvd.size()
    torch.cuda.set_device(hvd.local_rank())
    rows = []
    for nbytes in sizes:
        n = max(1, nbytes // torch.tensor([], dtype=dt).element_size())
        k = max(1, args.inflight)
        use_symm = args.symm and parts[0] == 'p2p' and hvd.size() > 1
        xs = [(hvd.symm_empty(n, dtype=dt) if use_symm else torch.empty(n, device='cuda', dtype=dt)).fill_(1) for _ in range(k)]
        x = xs[0]
        iters = max(5, min(100, int(1e9 // max(nbytes, 1 << 18))))
        name = f'sweep.{nbytes}'
        for _ in range(3):
            hvd.allreduce_(x, op=hvd.Sum, name=name)
        hvd.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if k == 1:
            for _ in range(iters):
                hvd.allreduce_(x, op=hvd.Sum, name=name)
        else:
            # k tensors in flight: the host enqueues ahead, kernels run back to back on the hvd stream
            for _ in range(max(1, iters // k)):
                hs = [hvd.allreduce_async_(xs[i], op=hvd.Sum, name=f'{name}.{i}') for i in range(k)]
                for h in hs:
                    hvd.synchronize(h)
            iters = max(1, iters // k) * k
        e1.record()
        torch.cuda.synchronize()
        # value check (a wrong-but-fast kernel must not produce a number)
        x.fill_(1)
        hvd.allreduce_(x, op=hvd.Sum, name=name)
        torch.cuda.synchronize()
        probe = torch.cat([x[:64].float(), x[-64:].float(), x[n // 2: n // 2 + 64].float()])
        assert torch.all(probe == size), f'{cfg}: wrong allreduce result at {nbytes} B: {probe[:4].tolist()}'
        ms = hvd.allreduce(torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64), op=hvd.Max, name='sweep.ms').item()
        alg = nbytes / (ms / 1e3) / 1e9
        rows.append({'bytes': nbytes, 'us': round(ms * 1e3, 2), 'algbw_gbs': round(alg, 2),
                     'busbw_gbs': round(alg * 2 * (size - 1) / size, 2)})
    results[cfg + (':symm' if args.symm else '')] = {'rows': rows, 'backend': hvd.gpu_backend_info(), 'tunables': hvd.tunable_params()}
    if rank == 0:
        print(f'== {cfg}{" symm" if args.symm else ""} inflight={args.inflight}  [{hvd.gpu_backend_info()}]', flush=True)
        for r in rows:
            print(f"  {r['bytes']:>12d} B  {r['us']:>10.2f} us  alg {r['algbw_gbs']:>8.2f} GB/s  bus {r['busbw_gbs']:>8.2f} GB/s", flush=True)
    hvd.shutdown()
if int(os.environ.get('RANK', os.environ.get('HOROVOD_RANK', '0'))) == 0 and args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump({'n_gpus': int(os.environ.get('WORLD_SIZE', os.environ.get('HOROVOD_SIZE', '1'))), 'dtype': args.dtype,
                   'results': results}, f, indent=1)
