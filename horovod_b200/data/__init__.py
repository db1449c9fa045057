from horovod_b200.data.data_loader_base import AsyncDataLoaderMixin, BaseDataLoader  # noqa: F401
