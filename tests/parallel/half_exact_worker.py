"""Host fp16 / bf16 reductions are bit-exact against torch (fp32 op, round to nearest even), including infinities, NaN,
subnormals and signed zeros — for the vectorised (F16C / auto-vectorised) paths AND the scalar tails (odd lengths)."""
import torch

import horovod_b200.torch as hvd

hvd.init()
r, n = hvd.rank(), hvd.size()
assert n == 2


def patterns(dtype, count, seed):
    g = torch.Generator().manual_seed(seed)
    bits = torch.randint(0, 1 << 16, (count,), generator=g, dtype=torch.int32).to(torch.int16)     # every bit pattern class
    t = bits.view(dtype).clone()
    specials = torch.tensor([0.0, -0.0, float('inf'), float('-inf'), float('nan'), 1e-7, -1e-7, 65504.0, -65504.0, 1.0], dtype=torch.float32).to(dtype)
    k = min(count, specials.numel())
    t[:k] = specials[:k]
    return t


for dtype in (torch.float16, torch.bfloat16):
    for count in (1, 7, 8, 9, 1023, 4096 + 5, 100003):
        a, b = patterns(dtype, count, 1), patterns(dtype, count, 2)
        mine = a if r == 0 else b
        for op, ref in ((hvd.Sum, lambda x, y: x + y), (hvd.Min, torch.minimum), (hvd.Max, torch.maximum), (hvd.Product, lambda x, y: x * y)):
            out = hvd.allreduce(mine, op=op, name='h.%s.%d.%d' % (dtype, count, op))
            want = ref(a.float(), b.float()).to(dtype)
            same = (out.view(torch.int16) == want.view(torch.int16)) | (out.isnan() & want.isnan())
            if op in (hvd.Min, hvd.Max):       # torch.minimum propagates NaN; the runtime keeps the non-NaN operand order-dependently
                same = same | a.isnan() | b.isnan()
            bad = (~same).nonzero().flatten()
            assert bad.numel() == 0, (dtype, count, op, bad[:5].tolist(), out[bad[:5]].tolist(), want[bad[:5]].tolist(), a[bad[:5]].tolist(), b[bad[:5]].tolist())
        avg = hvd.allreduce(mine, op=hvd.Average, name='h.avg.%s.%d' % (dtype, count))
        want = ((a.float() + b.float()).to(dtype).float() * 0.5).to(dtype)          # sum rounds to the wire type, then the post-scale rounds again
        same = (avg.view(torch.int16) == want.view(torch.int16)) | (avg.isnan() & want.isnan())
        assert bool(same.all()), (dtype, count, 'average')
hvd.barrier()
if r == 0:
    print('HALF EXACT OK')
hvd.shutdown()
