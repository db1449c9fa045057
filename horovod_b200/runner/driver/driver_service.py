"""Find the network interfaces that are routable between ALL hosts of the job before launching it: a driver service
runs here, one task service is started on every host (ssh), each task pings the next one over every interface and the
intersection of interfaces that worked is returned (reference horovod/runner/driver/driver_service.py)."""
import os
import sys

from horovod_b200.runner.common.service import driver_service
from horovod_b200.runner.common.util import codec, hosts, safe_shell_exec, timeout
from horovod_b200.runner.util import network, threads


class HorovodRunDriverService(driver_service.BasicDriverService):
    NAME = 'horovod driver service'

    def __init__(self, num_hosts, key, nics):
        super(HorovodRunDriverService, self).__init__(num_hosts, HorovodRunDriverService.NAME, key, nics)


class HorovodRunDriverClient(driver_service.BasicDriverClient):
    def __init__(self, driver_addresses, key, verbose, match_intf=False):
        super(HorovodRunDriverClient, self).__init__(HorovodRunDriverService.NAME, driver_addresses, key, verbose, match_intf=match_intf)


def _launch_task_servers(all_host_names, local_host_names, driver_addresses, settings):
    """Starts `python -m horovod_b200.runner.task_fn` on every host (locally or through ssh)."""
    from horovod_b200.runner.mesh_run import get_ssh_command

    def _exec_command(command):
        host_output = safe_shell_exec.execute(command)
        if host_output != 0:
            print('Launching task function was not successful: exit code {}'.format(host_output))
            os._exit(host_output)
        return host_output

    args_list = []
    num_hosts = len(all_host_names)
    for index in range(num_hosts):
        host_name = all_host_names[index]
        command = ('{python} -m horovod_b200.runner.task_fn {index} {num_hosts} {driver_addresses} {settings}'
                   .format(python=sys.executable, index=codec.dumps_base64(index), num_hosts=codec.dumps_base64(num_hosts),
                           driver_addresses=codec.dumps_base64(driver_addresses), settings=codec.dumps_base64(settings)))
        if host_name not in local_host_names:
            command = get_ssh_command(command, host=host_name, port=settings.ssh_port, identity_file=settings.ssh_identity_file)
        args_list.append([command])
    # Each thread will use ssh command to launch the server on one task. If an error occurs in one thread, entire
    # process will be terminated. Otherwise, threads will keep running and ssh session -- and the task server -- will be
    # bound to the thread. In case, the horovodrun process dies, all the ssh sessions and all the task servers will die too.
    threads.execute_function_multithreaded(_exec_command, args_list, block_until_all_done=False)


def _run_probe(driver, settings, num_hosts):
    # wait for all the hosts to register with the service service.
    if settings.verbose >= 2:
        print('Waiting for the hosts to acknowledge.')
    driver.wait_for_initial_registration(settings.start_timeout)
    tasks = [
        __import__('horovod_b200.runner.task.task_service', fromlist=['HorovodRunTaskClient']).HorovodRunTaskClient(
            index, driver.task_addresses_for_driver(index), settings.key, settings.verbose) for index in range(num_hosts)]
    # Notify all the drivers that the initial registration is complete.
    for task in tasks:
        task.notify_initial_registration_complete()
    if settings.verbose >= 2:
        print('Notified all the hosts that the registration is complete.')
    # Each worker should probe the interfaces of the next worker in a ring manner and filter only the routed ones --
    # it should filter out interfaces that are not really connected to any external networks such as lo0 with address 127.0.0.1.
    if settings.verbose >= 2:
        print('Waiting for hosts to perform host-to-host interface checking.')
    driver.wait_for_task_to_task_address_updates(settings.start_timeout)
    if settings.verbose >= 2:
        print('Host-to-host interface checking successful.')
    # Determine a set of common interfaces for task-to-task communication.
    nics = set(driver.task_addresses_for_tasks(0).keys())
    for index in range(1, num_hosts):
        nics.intersection_update(driver.task_addresses_for_tasks(index).keys())
    if not nics:
        raise Exception('Unable to find a set of common task-to-task communication interfaces: %s' %
                        [(index, driver.task_addresses_for_tasks(index)) for index in range(num_hosts)])
    return nics


def get_common_interfaces(settings, all_host_names, remote_host_names=None, fn_cache=None):
    """Interfaces common to all hosts; a single-host job returns the local loopback-capable set without probing."""
    if remote_host_names is None:
        remote_host_names = network.filter_local_addresses(all_host_names)
    if len(remote_host_names) > 0:
        if settings.nics:
            # If args.nics is provided, we will use those interfaces. All the workers must have at least one of those.
            return settings.nics
        # Find the set of common, routed interfaces on all the hosts (remote and local) and specify it in the args.
        local_host_names = set(all_host_names) - set(remote_host_names)
        if not isinstance(settings.start_timeout, timeout.Timeout):
            settings.start_timeout = timeout.Timeout(settings.start_timeout or 30,
                                                     message='Timed out waiting for {activity}. Please check connectivity between servers.')
        driver = HorovodRunDriverService(len(all_host_names), settings.key, settings.nics)
        try:
            _launch_task_servers(all_host_names, local_host_names, driver.addresses(), settings)
            return _run_probe(driver, settings, len(all_host_names))
        finally:
            driver.shutdown()
    if settings.verbose >= 2:
        print('All hosts are local, finding the interfaces with address 127.0.0.1')
    # If all the given hosts are local, find the interfaces with address 127.0.0.1
    nics = set()
    import psutil, socket
    for iface, addrs in psutil.net_if_addrs().items():
        if settings.nics and iface not in settings.nics:
            continue
        for addr in addrs:
            if addr.family == socket.AF_INET and addr.address == '127.0.0.1':
                nics.add(iface)
                break
    if len(nics) == 0:
        raise ValueError('No interface is found for address 127.0.0.1.')
    return nics
