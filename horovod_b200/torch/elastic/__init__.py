"""`hvd.elastic`: fault-tolerant training for PyTorch (API parity: horovod/torch/elastic/__init__.py)."""
from horovod_b200.common.elastic import ObjectState, run_fn
from horovod_b200.torch.elastic.sampler import ElasticSampler  # noqa: F401
from horovod_b200.torch.elastic.state import TorchState  # noqa: F401
from horovod_b200.torch.mpi_ops import init, shutdown


def run(func):
    """Decorator: runs `func(state, ...)` inside the elastic retry loop.

    On a failed collective the last committed state is restored; on host changes training continues from the
    current state; either way the runtime re-initialises (new rendezvous, new ranks) and `state.sync()` re-broadcasts
    from the new rank 0."""
    return run_fn(func, _reset)


def _reset():
    shutdown()
    init()
