"""What every RayExecutor flavour implements (reference horovod/ray/adapter.py: `BaseParams` :6-19, `Adapter` :22-126).

`RayExecutor` turns its keyword arguments into a params object (`StaticParams` in runner.py, `ElasticParams` in
elastic_v2.py); the params object builds its adapter; the executor's public methods are one-line delegations."""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional


@dataclass
class BaseParams:
    cpus_per_worker: int = 1
    use_gpu: bool = False
    gpus_per_worker: Optional[int] = None

    def __post_init__(self):
        if self.gpus_per_worker and not self.use_gpu:
            raise ValueError('gpus_per_worker is set, but use_gpu is False. use_gpu must be True if gpus_per_worker is set.')
        if self.use_gpu and isinstance(self.gpus_per_worker, int) and self.gpus_per_worker < 1:
            raise ValueError(f'gpus_per_worker must be >= 1: Got {self.gpus_per_worker}.')
        self.gpus_per_worker = (self.gpus_per_worker or 1) if self.use_gpu else 0

    @property
    def elastic(self) -> bool:
        raise NotImplementedError

    @property
    def adapter(self):
        """The Adapter class that runs a job described by these params."""
        raise NotImplementedError


class Adapter(ABC):
    """Creates the workers of one job and runs functions on them."""

    @abstractmethod
    def start(self, executable_cls: type = None, executable_args: Optional[List] = None,
              executable_kwargs: Optional[Dict] = None, extra_env_vars: Optional[Dict] = None):
        """Creates the workers, wires the rendezvous and (static jobs) instantiates `executable_cls` on every worker."""

    @abstractmethod
    def execute(self, fn: Callable[[Any], Any], callbacks: Optional[List[Callable]] = None) -> List[Any]:
        """fn(executable) on every worker; results in rank order."""

    @abstractmethod
    def run(self, fn: Callable[..., Any], args: Optional[List] = None, kwargs: Optional[Dict] = None,
            callbacks: Optional[List[Callable]] = None) -> List[Any]:
        """fn(*args, **kwargs) on every worker; results in rank order."""

    @abstractmethod
    def run_remote(self, fn: Callable[..., Any], args: Optional[List] = None, kwargs: Optional[Dict] = None,
                   callbacks: Optional[List[Callable]] = None) -> List[Any]:
        """Like `run`, but returns the backend's futures (Ray ObjectRefs) without waiting."""

    @abstractmethod
    def execute_single(self, fn: Callable[[Any], Any]) -> Any:
        """fn(executable) on the rank-0 worker only."""

    @abstractmethod
    def shutdown(self):
        """Destroys the workers and what was created for them (placement group, rendezvous server)."""
