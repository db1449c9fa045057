"""Helpers for code that runs inside a training task (role parity: horovod/spark/task/__init__.py exports
`get_available_devices`).  The reference's task-side services (`task_service.py`, `*_exec_fn.py`: a per-task RPC server through
which mpirun's rsh agent starts the worker command) have no counterpart: the function is called in the task directly
(`runner/cluster_job.py`)."""
from horovod_b200.spark.task import task_info  # noqa: F401
from horovod_b200.spark.task.task_info import get_available_devices, set_resources  # noqa: F401
