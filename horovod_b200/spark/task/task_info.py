"""What the scheduler gave THIS training process (role parity: horovod/spark/task/task_info.py).

In the reference the Spark task records its resources here before it executes the worker command in a child process.  Here the
training function runs inside the Spark task's own Python worker, so the live `TaskContext` is asked first; `set_resources`
remains for schedulers (or tests) that hand resources over explicitly."""


class TaskInfo(object):
    def __init__(self):
        self.resources = {}


_info = TaskInfo()


def set_resources(resources):
    """`resources`: name -> object with `.addresses` (pyspark's ResourceInformation) or a plain list of addresses."""
    _info.resources = dict(resources or {})


def get_available_devices():
    """GPU addresses assigned to this task: explicitly recorded resources win, then the running Spark task's
    `TaskContext.resources()['gpu']`, else []."""
    gpu = _info.resources.get('gpu')
    if gpu is not None:
        return list(getattr(gpu, 'addresses', gpu))
    from horovod_b200.spark.common.util import get_available_devices as from_task_context
    return from_task_context()
