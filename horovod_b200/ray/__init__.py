"""Ray integration (parity: horovod/ray/__init__.py: RayExecutor, BaseHorovodWorker, ElasticRayExecutor)."""
from horovod_b200.ray.worker import BaseHorovodWorker  # noqa: F401
from horovod_b200.ray.runner import RayExecutor, RayBackend  # noqa: F401
from horovod_b200.ray.elastic import RayHostDiscovery, ElasticRayExecutor  # noqa: F401

__all__ = ['RayExecutor', 'BaseHorovodWorker', 'ElasticRayExecutor', 'RayHostDiscovery', 'RayBackend']
