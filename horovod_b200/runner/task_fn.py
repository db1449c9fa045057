"""`python -m horovod_b200.runner.task_fn <index> <num_hosts> <driver_addresses> <settings>` — the per-host probe agent of
the NIC discovery (role parity: horovod/runner/task_fn.py).  The implementation lives with the agent's service in
`horovod_b200.runner.task.task_service`; this module keeps the reference's entry-point name."""
from horovod_b200.runner.task.task_service import main

if __name__ == '__main__':
    main()
