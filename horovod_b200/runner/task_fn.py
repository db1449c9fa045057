"""Entry point of the per-host probe task: register with the driver, ping the next task over every interface, report
the interfaces that worked (reference horovod/runner/task_fn.py)."""
import sys
import time

from horovod_b200.runner.common.util import codec, host_hash, timeout as timeout_util
from horovod_b200.runner.driver import driver_service
from horovod_b200.runner.task import task_service


def _task_fn(index, num_hosts, driver_addresses, settings):
    task = task_service.HorovodRunTaskService(index, settings.key, settings.nics)
    try:
        driver = driver_service.HorovodRunDriverClient(driver_addresses, settings.key, settings.verbose)
        driver.register_task(index, task.addresses(), host_hash.host_hash())
        tmout = settings.start_timeout if isinstance(settings.start_timeout, timeout_util.Timeout) else \
            timeout_util.Timeout(settings.start_timeout or 30, message='Timed out waiting for {activity}.')
        task.wait_for_initial_registration(tmout)
        # Tasks ping each other in a circular fashion to determine interfaces reachable within the cluster.
        next_task_index = (index + 1) % num_hosts
        next_task_addresses = driver.all_task_addresses(next_task_index)
        # We request interface matching to weed out all the NAT'ed interfaces.
        next_task = task_service.HorovodRunTaskClient(next_task_index, next_task_addresses, settings.key, settings.verbose,
                                                      match_intf=True, attempts=10)
        driver.register_task_to_task_addresses(next_task_index, next_task.addresses())
        # Notify the next task that the address checks are completed.
        next_task.notify_initial_registration_complete()
        time.sleep(2)  # let the driver collect everything before the ssh session is torn down
    finally:
        task.shutdown()


if __name__ == '__main__':
    index = codec.loads_base64(sys.argv[1])
    num_hosts = codec.loads_base64(sys.argv[2])
    driver_addresses = codec.loads_base64(sys.argv[3])
    settings = codec.loads_base64(sys.argv[4])
    _task_fn(index, num_hosts, driver_addresses, settings)
