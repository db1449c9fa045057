#!/usr/bin/env python
"""Single-GPU micro-benchmark / ncu target for the communication kernels.

With ONE simulated rank the kernels run their full pack -> (barrier) -> reduce -> (barrier) -> unpack pipeline against
local HBM instead of NVLink, which is what `ncu` can replay (a multi-kernel rendezvous would deadlock under ncu's
serialisation).  With N simulated ranks (N concurrently running kernels on one GPU) the flag protocol is exercised and
timed with all "peers" in local HBM.

    python bench/kernel_micro.py                 # table of achieved HBM bandwidth vs the measured copy peak
    ncu --set full ... python bench/kernel_micro.py --ncu   # one launch of each kernel for profiling
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from horovod_b200.common.basics import load_library  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--ncu', action='store_true')
p.add_argument('--mb', type=int, default=256)
p.add_argument('--out', default=None)
args = p.parse_args()
lib = load_library()
lib.hvd_sim_allreduce.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                  ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_float)]
lib.hvd_sim_inplace.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
except Exception:
    pass
hbm_peak = peaks.get('hbm_gbs', 6650.0)
FP32, BF16 = 7, 10


def run_packed(n, nbytes, variant, ctas, dtype=torch.float32, wire=FP32, repeats=5, ntensors=1):
    per = nbytes // ntensors // dtype.itemsize
    ins = [[torch.ones(per, device='cuda', dtype=dtype) for _ in range(ntensors)] for _ in range(n)]
    outs = [[torch.empty_like(t) for t in row] for row in ins]
    counts = (ctypes.c_int64 * ntensors)(*[per] * ntensors)
    ip = (ctypes.c_uint64 * (n * ntensors))(*[t.data_ptr() for r in ins for t in r])
    op = (ctypes.c_uint64 * (n * ntensors))(*[t.data_ptr() for r in outs for t in r])
    ms = ctypes.c_float(0)
    dt = FP32 if dtype == torch.float32 else BF16
    rc = lib.hvd_sim_allreduce(n, 0, ntensors, counts, ip, op, dt, wire, 1, variant, ctas, 1.0, 1.0 / n, 1, ctypes.byref(ms))
    assert rc == 0, rc
    rc = lib.hvd_sim_allreduce(n, 0, ntensors, counts, ip, op, dt, wire, 1, variant, ctas, 1.0, 1.0 / n, repeats, ctypes.byref(ms))
    assert rc == 0, rc
    assert torch.allclose(outs[0][0][:8].float(), torch.ones(8, device='cuda'))
    return ms.value / repeats


def run_inplace(n, nbytes, ctas, repeats=5):
    ts = [torch.ones(nbytes // 4, device='cuda') for _ in range(n)]
    ptrs = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in ts])
    ms = ctypes.c_float(0)
    rc = lib.hvd_sim_inplace(n, 0, nbytes, ptrs, FP32, 1, ctas, 1.0, 1, ctypes.byref(ms))
    assert rc == 0, rc
    for t in ts:
        t.fill_(1)
    rc = lib.hvd_sim_inplace(n, 0, nbytes, ptrs, FP32, 1, ctas, 1.0 / n, repeats, ctypes.byref(ms))
    assert rc == 0, rc
    return ms.value / repeats


nbytes = args.mb << 20
rows = []
if args.ncu:
    run_packed(1, 64 << 20, 0, 128, repeats=1, ntensors=161)
    run_packed(1, 64 << 20, 1, 128, repeats=1, ntensors=161)
    run_inplace(1, 64 << 20, 128, repeats=1)
    sys.exit(0)
for ctas in (32, 64, 128, 256):
    # one simulated rank: traffic = pack (r+w) + reduce (r+w) [+ unpack (r+w)] of the message
    ms = run_packed(1, nbytes, 0, ctas)
    rows.append({'kernel': 'allreduce one-shot (pack+reduce->out)', 'ranks': 1, 'ctas': ctas, 'ms': ms, 'hbm_bytes': 4 * nbytes})
    ms = run_packed(1, nbytes, 1, ctas)
    rows.append({'kernel': 'allreduce two-shot (pack+reduce+unpack)', 'ranks': 1, 'ctas': ctas, 'ms': ms, 'hbm_bytes': 6 * nbytes})
    ms = run_inplace(1, nbytes, ctas)
    rows.append({'kernel': 'zero-copy in-place', 'ranks': 1, 'ctas': ctas, 'ms': ms, 'hbm_bytes': 2 * nbytes})
for n, ctas in ((2, 64), (4, 32), (8, 16)):
    ms = run_packed(n, nbytes // 4, 1, ctas)
    rows.append({'kernel': 'allreduce two-shot, N kernels on one GPU', 'ranks': n, 'ctas': ctas, 'ms': ms,
                 'hbm_bytes': n * (nbytes // 4) * (4 + 2)})
    ms = run_inplace(n, nbytes // 4, ctas)
    rows.append({'kernel': 'zero-copy in-place, N kernels on one GPU', 'ranks': n, 'ctas': ctas, 'ms': ms,
                 'hbm_bytes': n * (nbytes // 4) * 2})
for r in rows:
    r['gbs'] = r['hbm_bytes'] / (r['ms'] / 1e3) / 1e9
    r['frac_of_measured_hbm_peak'] = r['gbs'] / hbm_peak
    print(f"{r['kernel']:<48s} ranks={r['ranks']} ctas={r['ctas']:<4d} {r['ms']:8.3f} ms  {r['gbs']:8.1f} GB/s  "
          f"{100 * r['frac_of_measured_hbm_peak']:5.1f}% of measured copy peak ({hbm_peak:.0f} GB/s)")
if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump({'message_mb': args.mb, 'hbm_peak_gbs': hbm_peak, 'rows': rows}, open(args.out, 'w'), indent=1)
