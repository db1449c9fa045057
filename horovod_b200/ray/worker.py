"""The actor class of the Ray integration (reference horovod/ray/worker.py:8-65 `BaseHorovodWorker`).

The scheduler-independent worker is `horovod_b200.runner.cluster_job.WorkerActor`; this subclass adds the method names
user code written against the reference calls on the actor handles (`get_gpu_ids`, `update_env_vars`, `env_vars`,
`start_executable`, `set_queue`)."""
import os

from horovod_b200.runner.cluster_job import WorkerActor


class BaseHorovodWorker(WorkerActor):
    executable = None

    def __init__(self, world_rank=0, world_size=1):
        super().__init__(world_rank)
        # an actor is its own process: its environment describes this worker until the executor hands out the rank table
        os.environ['HOROVOD_HOSTNAME'] = self.node_id()
        os.environ['HOROVOD_RANK'] = str(world_rank)
        os.environ['HOROVOD_SIZE'] = str(world_size)

    def get_gpu_ids(self):
        """GPU ids Ray assigned to this actor (falls back to CUDA_VISIBLE_DEVICES outside Ray)."""
        try:
            import ray
            return list(ray.get_gpu_ids())
        except Exception:  # noqa: BLE001 - not inside a Ray worker
            return self.gpu_ids()

    def update_env_vars(self, env_vars):
        """Updates the environment of the actor process; values are stringified."""
        return self.update_env(env_vars)

    def env_vars(self):
        return self.env()

    def start_executable(self, executable_cls=None, executable_args=None, executable_kwargs=None):
        """Instantiates `executable_cls` inside the actor; `execute(fn)` then calls fn(that object)."""
        args, kwargs = tuple(executable_args or ()), dict(executable_kwargs or {})
        if executable_cls is not None:
            self.executable = executable_cls(*args, **kwargs)
        return True

    def execute(self, func, args=(), kwargs=None):
        """func(executable) when an executable was started, else func(*args, **kwargs)."""
        if self.executable is not None and not args and not kwargs:
            return func(self.executable)
        return super().execute(func, args, kwargs)

    def set_queue(self, queue):
        """Worker -> driver log channel (see ray_logger): whatever `ray_logger.log` receives is put on `queue`."""
        from horovod_b200.ray import ray_logger
        ray_logger.configure(queue)
        return True
