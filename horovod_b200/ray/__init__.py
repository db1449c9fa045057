"""Ray integration (parity: horovod/ray/__init__.py: RayExecutor, ElasticRayExecutor, RayHostDiscovery)."""
from horovod_b200.ray.runner import RayExecutor, RayBackend  # noqa: F401
from horovod_b200.ray.elastic import RayHostDiscovery, ElasticRayExecutor  # noqa: F401
from horovod_b200.runner.cluster_job import WorkerActor as BaseHorovodWorker  # noqa: F401  (the reference's worker actor name)
