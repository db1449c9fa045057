"""Ray integration (role parity: horovod/ray): RayExecutor runs one worker actor per slot and wires the rendezvous.
`ray` is not part of this image; the executor imports it lazily. The placement / rank logic (horovod_b200.ray.strategy,
horovod_b200.ray.elastic.RayHostDiscovery) is plain Python and unit-tested without Ray."""
from horovod_b200.ray.elastic import RayHostDiscovery  # noqa: F401
from horovod_b200.ray.runner import RayExecutor  # noqa: F401
