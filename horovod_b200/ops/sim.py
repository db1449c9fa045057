"""ctypes front end of `csrc/common/sim_api.cc`: run the P2P kernels with N simulated ranks on ONE GPU.

Every "rank" is a set of plain device tensors plus a kernel of its own on its own stream; the cross-"GPU" flag
protocol, chunk ownership and descriptor walking are exactly what runs across NVLink — only the address mapping differs.
All functions take lists indexed [rank][tensor] of CUDA tensors and return the average milliseconds per repeat where that
makes sense.  A non-zero native return code raises RuntimeError (-3 = a barrier of the simulation timed out).
"""
import ctypes

import torch

DTYPE_CODE = {torch.uint8: 0, torch.int8: 1, torch.int16: 3, torch.int32: 4, torch.int64: 5, torch.float16: 6, torch.float32: 7,
              torch.float64: 8, torch.bool: 9, torch.bfloat16: 10}
ONESHOT, TWOSHOT, NVLS = 0, 1, 2
AVERAGE, SUM, ADASUM, MIN, MAX, PRODUCT = 0, 1, 2, 3, 4, 5
_lib = None


def lib():
    global _lib
    if _lib is None:
        from horovod_b200.common.basics import load_library
        L = load_library()
        u64p, i64p, f32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float)
        L.hvd_sim_allreduce.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i64p, u64p, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, f32p]
        L.hvd_sim_allgather.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, u64p, u64p, ctypes.c_int]
        L.hvd_sim_adasum.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i64p, u64p, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_double]
        L.hvd_sim_inplace.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_int, f32p]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed with code %d%s' % (what, rc, ' (simulated barrier timed out)' if rc == -3 else ''))


def _ptrs(rows):
    flat = [t.data_ptr() for row in rows for t in row]
    return (ctypes.c_uint64 * len(flat))(*flat)


def allreduce(ins, outs, op=SUM, variant=TWOSHOT, ctas=8, wire_dtype=None, prescale=1.0, postscale=1.0, repeats=1, device=0):
    """ins/outs: [nranks][ntensors] (outs may alias ins).  wire_dtype: on-the-wire dtype for fp32 tensors (fused cast)."""
    n, t = len(ins), len(ins[0])
    dtype = ins[0][0].dtype
    counts = (ctypes.c_int64 * t)(*[x.numel() for x in ins[0]])
    ms = ctypes.c_float(0)
    _check(lib().hvd_sim_allreduce(n, device, t, counts, _ptrs(ins), _ptrs(outs), DTYPE_CODE[dtype], DTYPE_CODE[wire_dtype or dtype], op,
                                   variant, ctas, prescale, postscale, repeats, ctypes.byref(ms)), 'sim allreduce')
    return ms.value / max(repeats, 1)


def inplace_allreduce(tensors, op=SUM, ctas=8, scale=1.0, repeats=1, device=0):
    """tensors: one tensor per simulated rank, reduced in place by the zero-copy kernel (P2P two-shot variant)."""
    n = len(tensors)
    ptrs = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in tensors])
    ms = ctypes.c_float(0)
    nbytes = tensors[0].numel() * tensors[0].element_size()
    _check(lib().hvd_sim_inplace(n, device, nbytes, ptrs, DTYPE_CODE[tensors[0].dtype], op, ctas, scale, repeats, ctypes.byref(ms)),
           'sim in-place allreduce')
    return ms.value / max(repeats, 1)


def allgather(ins, outs, ctas=8, device=0):
    """ins[r]: rank r's contribution (all the same byte size); outs[r]: nranks * that size."""
    n = len(ins)
    nbytes = ins[0].numel() * ins[0].element_size()
    ip = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in ins])
    op = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in outs])
    _check(lib().hvd_sim_allgather(n, device, nbytes, ip, op, ctas), 'sim allgather')


def adasum(ins, outs, ctas=8, prescale=1.0, postscale=1.0, device=0):
    n, t = len(ins), len(ins[0])
    counts = (ctypes.c_int64 * t)(*[x.numel() for x in ins[0]])
    _check(lib().hvd_sim_adasum(n, device, t, counts, _ptrs(ins), _ptrs(outs), DTYPE_CODE[ins[0][0].dtype], ctas, prescale, postscale),
           'sim adasum')
