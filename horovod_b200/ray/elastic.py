"""Host discovery backed by the Ray cluster state (role parity: horovod/ray/elastic_v2.py RayHostDiscovery)."""
from horovod_b200.runner.elastic.discovery import HostDiscovery


class RayHostDiscovery(HostDiscovery):
    """Uses Ray global state to obtain host mapping. Assumes that the whole global state is available for usage."""

    def __init__(self, use_gpu=False, cpus_per_worker=1, gpus_per_worker=1, nodes_fn=None):
        self.use_gpu = use_gpu
        self.cpus_per_worker = cpus_per_worker
        self.gpus_per_worker = gpus_per_worker
        self._nodes_fn = nodes_fn  # injectable for tests; defaults to ray.nodes

    def find_available_hosts_and_slots(self):
        """Returns a dict mapping <hostname> -> <number of slots>."""
        if self._nodes_fn is None:
            import ray
            nodes = ray.nodes()
        else:
            nodes = self._nodes_fn()
        host_mapping = {}
        for node in nodes:
            if not node.get('alive', node.get('Alive', False)):
                continue
            hostname = node.get('NodeManagerAddress') or node.get('NodeManagerHostname')
            resources = node.get('Resources', {})
            slots = resources.get('CPU', 0) // self.cpus_per_worker
            if self.use_gpu:
                slots = min(slots, resources.get('GPU', 0) // self.gpus_per_worker)
            slots = int(slots)
            if slots:
                host_mapping[hostname] = slots
        return host_mapping
