"""Data-loader helpers (reference: test/single/test_torch.py::test_async_data_loader-style coverage of
horovod/data/data_loader_base.py)."""
import threading
import time

import pytest
import torch

from horovod_b200.data import AsyncDataLoaderMixin, BaseDataLoader, DevicePrefetcher


class _Range(BaseDataLoader):
    def __init__(self, n, fail_at=None, delay=0.0):
        self.n, self.fail_at, self.delay = n, fail_at, delay
        self.epochs = 0

    def __len__(self):
        return self.n

    def _iterate(self):
        self.epochs += 1
        for i in range(self.n):
            if self.fail_at is not None and i == self.fail_at:
                raise ValueError('boom at %d' % i)
            if self.delay:
                time.sleep(self.delay)
            yield None if i == 1 else i  # None is a legal batch


class _AsyncRange(AsyncDataLoaderMixin, _Range):
    def _process_batch(self, b):
        return ('p', b)


def test_sync_base_loader():
    assert list(_Range(4)) == [0, None, 2, 3]


@pytest.mark.parametrize('qsize', [0, 1, 8])
def test_async_loader_epochs(qsize):
    ld = _AsyncRange(5, async_loader_queue_size=qsize)
    for _ in range(3):
        assert list(ld) == [('p', 0), ('p', None), ('p', 2), ('p', 3), ('p', 4)]
    ld.close_async_loader()
    ld.close_async_loader()
    assert len(ld) == 5


def test_async_loader_error_forwarding_and_recovery():
    ld = _AsyncRange(5, fail_at=3, async_loader_queue_size=2)
    got = []
    with pytest.raises(ValueError, match='boom at 3'):
        for b in ld:
            got.append(b)
    assert got == [('p', 0), ('p', None), ('p', 2)]
    ld.fail_at = None
    assert len(list(ld)) == 5
    ld.close_async_loader()


def test_async_loader_close_while_blocked():
    ld = _AsyncRange(1000, async_loader_queue_size=1)
    it = iter(ld)
    next(it)
    t0 = time.time()
    ld.close_async_loader()  # the producer is blocked on a full queue
    assert time.time() - t0 < 5
    assert not any(t.name == 'hvd-data-prefetch' and t.is_alive() for t in threading.enumerate())


def test_async_loader_prefetches_ahead():
    ld = _AsyncRange(6, delay=0.02, async_loader_queue_size=6)
    it = iter(ld)
    next(it)
    time.sleep(0.3)  # the producer keeps going while the consumer is idle
    t0 = time.time()
    rest = list(it)
    assert len(rest) == 5 and time.time() - t0 < 0.1
    ld.close_async_loader()


def test_device_prefetcher_cpu_passthrough():
    data = [(torch.full((2, 3), float(i)), {'y': torch.tensor([i])}) for i in range(5)]
    out = list(DevicePrefetcher(data, device='cpu', depth=2))
    assert len(out) == 5
    for i, (x, d) in enumerate(out):
        assert torch.equal(x, data[i][0]) and torch.equal(d['y'], data[i][1]['y'])


@pytest.mark.gpu
def test_device_prefetcher_cuda():
    data = [(torch.randn(4, 3, 8, 8), torch.tensor([i, i + 1])) for i in range(7)]
    pf = DevicePrefetcher(data, device='cuda:0', depth=2, channels_last=True)
    out = list(pf)
    assert len(out) == 7 and len(pf) == 7
    for i, (x, y) in enumerate(out):
        assert x.is_cuda and x.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(x.cpu(), data[i][0]) and torch.equal(y.cpu(), data[i][1])
    assert pf.h2d_bytes == sum(x.numel() * 4 + y.numel() * 8 for x, y in data)
