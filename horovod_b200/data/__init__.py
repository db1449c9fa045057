from .data_loader_base import BaseDataLoader, AsyncDataLoaderMixin  # noqa: F401
from .device_prefetcher import DevicePrefetcher  # noqa: F401
