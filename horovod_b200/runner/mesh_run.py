"""Static launcher: starts the HTTP rendezvous server, computes the rank layout, builds each slot's environment and
command, runs them locally or through ssh, streams their output, and tears everything down when one rank fails.

Role parity: horovod/runner/gloo_run.py (create_slot_env_vars :66-77, _slot_info_to_command_fn, _exec_command_fn,
launch_gloo :242, gloo_run :295; the elastic variant lives in elastic_run below).  The worker transport is this repo's
native TCP/shm mesh (csrc/transport), hence the file name; the environment contract (HOROVOD_RANK, HOROVOD_SIZE,
HOROVOD_LOCAL_RANK, ..., HOROVOD_GLOO_RENDEZVOUS_ADDR/PORT, HOROVOD_CONTROLLER) is the reference's.
"""
import os
import shlex
import sys
import threading
import time

from horovod_b200.runner.common.util import env as env_util
from horovod_b200.runner.common.util import hosts, safe_shell_exec
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.util import network, threads


class MultiFile(object):
    def __init__(self, files):
        self._files = files

    def write(self, text):
        for f in self._files:
            f.write(text)

    def flush(self):
        for f in self._files:
            f.flush()


def _pad_rank(rank, size):
    width = len(str(size - 1))
    return str(rank).zfill(width)


def create_slot_env_vars(slot_info):
    # for the host name use the given one (NOT socket.gethostname()): that is what the rendezvous knows
    return {
        'HOROVOD_HOSTNAME': str(slot_info.hostname),
        'HOROVOD_RANK': str(slot_info.rank),
        'HOROVOD_SIZE': str(slot_info.size),
        'HOROVOD_LOCAL_RANK': str(slot_info.local_rank),
        'HOROVOD_LOCAL_SIZE': str(slot_info.local_size),
        'HOROVOD_CROSS_RANK': str(slot_info.cross_rank),
        'HOROVOD_CROSS_SIZE': str(slot_info.cross_size),
    }


def create_run_env_vars(server_ip, server_port, nics=None, elastic=False):
    run_envs = {
        'HOROVOD_GLOO_RENDEZVOUS_ADDR': server_ip,
        'HOROVOD_GLOO_RENDEZVOUS_PORT': str(server_port),
        'HOROVOD_CONTROLLER': 'gloo',
        'HOROVOD_CPU_OPERATIONS': 'gloo',
    }
    if nics:
        iface = list(nics)[0]
        run_envs['HOROVOD_GLOO_IFACE'] = iface
        run_envs['NCCL_SOCKET_IFNAME'] = ','.join(nics)
    if elastic:
        run_envs['HOROVOD_ELASTIC'] = '1'
    # workers must be able to import this package even when it is used from a source tree (not pip-installed)
    pkg_parent = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    existing = os.environ.get('PYTHONPATH', '')
    if pkg_parent not in existing.split(os.pathsep):
        run_envs['PYTHONPATH'] = pkg_parent + (os.pathsep + existing if existing else '')
    return run_envs


from horovod_b200.runner.util.remote import get_ssh_command  # noqa: E402,F401  (kept importable from here)


def _slot_info_to_command_fn(run_command, env, settings=None):
    def slot_info_to_command(slot_info):
        """Given a slot_info, creates a command used to start the worker on that slot."""
        env_vars = create_slot_env_vars(slot_info)
        horovod_rendez_env = ' '.join(f'{k}={shlex.quote(v)}' for k, v in env_vars.items())
        return f'{horovod_rendez_env} {run_command}'
    return slot_info_to_command


def _exec_command_fn(settings):
    """Returns exec_command(command, slot_info, events) -> (exit_code, timestamp)."""
    def _exec_command(command, slot_info, events):
        index = slot_info.rank
        host_name = slot_info.hostname
        host_address = network.resolve_host_address(host_name)
        local_addresses = network.get_local_host_addresses()
        if host_address not in local_addresses and host_name not in ('localhost', '127.0.0.1'):
            exports = ' '.join(f'{k}={shlex.quote(v)}' for k, v in os.environ.items()
                               if env_util.is_exportable(k) and k.startswith(('HOROVOD_', 'HVD_', 'NCCL_', 'CUDA_', 'PATH', 'PYTHONPATH', 'LD_LIBRARY_PATH')))
            command = get_ssh_command(f'cd {shlex.quote(os.getcwd())} > /dev/null 2>&1 ; {exports} {command}', host=host_name,
                                      port=settings.ssh_port, identity_file=settings.ssh_identity_file)
        if settings.verbose >= 2:
            print(command)
        stdout = stderr = None
        stdout_file = stderr_file = None
        if settings.output_filename:
            padded_rank = _pad_rank(index, settings.num_proc)
            output_dir_rank = os.path.join(settings.output_filename, 'rank.{rank}'.format(rank=padded_rank))
            os.makedirs(output_dir_rank, exist_ok=True)
            stdout_file = open(os.path.join(output_dir_rank, 'stdout'), 'w')
            stderr_file = open(os.path.join(output_dir_rank, 'stderr'), 'w')
            stdout = MultiFile([sys.stdout, stdout_file])
            stderr = MultiFile([sys.stderr, stderr_file])
        try:
            exit_code = safe_shell_exec.execute(command, index=index, stdout=stdout, stderr=stderr, events=events,
                                                prefix_output_with_timestamp=settings.prefix_output_with_timestamp)
            if exit_code != 0:
                print('Process {idx} exit with status code {ec}.'.format(idx=index, ec=exit_code))
        except Exception as e:
            print('Exception happened during safe_shell_exec, exception message: {message}'.format(message=e))
            exit_code = 1
        finally:
            if stdout_file:
                stdout_file.close()
            if stderr_file:
                stderr_file.close()
        return exit_code, time.time()
    return _exec_command


def launch_static(command, exec_command, settings, nics, env, server_ip):
    """Launches the job: one process per slot, first non-zero exit terminates everybody."""
    host_alloc_plan = hosts.get_host_assignments(hosts.parse_hosts(settings.hosts), settings.num_proc)
    rendezvous = RendezvousServer(settings.verbose)
    global_rendezv_port = rendezvous.start_server()
    rendezvous.init(host_alloc_plan)
    run_env = create_run_env_vars(server_ip, global_rendezv_port, nics)
    run_command = ' '.join(f'{k}={shlex.quote(v)}' for k, v in run_env.items()) + ' ' + command
    slot_info_to_command = _slot_info_to_command_fn(run_command, env)
    event = threading.Event()  # set as soon as one rank fails: the others are killed
    args_list = [[slot_info_to_command(slot_info), slot_info, [event]] for slot_info in host_alloc_plan]

    def exec_and_flag(cmd, slot_info, events):
        rc, ts = exec_command(cmd, slot_info, events)
        if rc != 0:
            event.set()
        return rc, ts

    try:
        res = threads.execute_function_multithreaded(exec_and_flag, args_list, block_until_all_done=True)
    finally:
        rendezvous.stop()
    failures = []
    for name, value in sorted(res.items(), key=lambda item: item[1][1]):
        exit_code, timestamp = value
        if exit_code != 0:
            failures.append((name, exit_code))
    if failures:
        raise RuntimeError('Horovod detected that one or more processes exited with non-zero status, thus causing the job '
                           'to be terminated. The first process to do so was:\nProcess name: {name}\nExit code: {code}\n'
                           .format(name=failures[0][0], code=failures[0][1]))


def mesh_run(settings, nics, env, server_ip, command):
    """`command` is a list (argv) or a string."""
    exec_command = _exec_command_fn(settings)
    if isinstance(command, (list, tuple)):
        command = ' '.join(shlex.quote(c) for c in command)
    launch_static(command, exec_command, settings, nics, env, server_ip)


def elastic_run(settings, env, command, discovery, min_np, max_np, elastic_timeout, reset_limit, cooldown_range=None):
    """Elastic launch: workers come and go with the discovered host set (reference gloo_run.py:303-380)."""
    from horovod_b200.runner.elastic.driver import ElasticDriver
    from horovod_b200.runner.elastic.rendezvous import create_rendezvous_handler
    if isinstance(command, (list, tuple)):
        command = ' '.join(shlex.quote(c) for c in command)
    rendezvous = RendezvousServer(settings.verbose)
    driver = ElasticDriver(rendezvous, discovery, min_np, max_np, timeout=elastic_timeout, reset_limit=reset_limit,
                           cooldown_range=cooldown_range, verbose=settings.verbose)
    handler = create_rendezvous_handler(driver)
    global_rendezv_port = rendezvous.start_server()
    handler.install(rendezvous)
    driver.wait_for_available_slots(min_np)
    nics = settings.nics
    server_ip = network.get_driver_ip(nics)
    run_env = create_run_env_vars(server_ip, global_rendezv_port, nics, elastic=True)
    run_command = ' '.join(f'{k}={shlex.quote(v)}' for k, v in run_env.items()) + ' ' + command
    exec_command = _exec_command_fn(settings)
    slot_info_to_command = _slot_info_to_command_fn(run_command, env)

    def create_worker(slot_info, events):
        return exec_command(slot_info_to_command(slot_info), slot_info, events)

    try:
        driver.start(settings.num_proc, create_worker)
        res = driver.get_results()
        driver.stop()
    finally:
        rendezvous.stop()
    if res.error_message is not None:
        raise RuntimeError(res.error_message)
    for name, value in sorted(res.worker_results.items(), key=lambda item: item[1][1]):
        exit_code, timestamp = value
        if exit_code != 0:
            raise RuntimeError('Horovod detected that one or more processes exited with non-zero status, thus causing the '
                               'job to be terminated. The first process to do so was:\nProcess name: {name}\nExit code: '
                               '{code}\n'.format(name=name, code=exit_code))
