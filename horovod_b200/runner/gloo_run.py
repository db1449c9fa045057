"""The reference's module name for the non-MPI launch path.

`horovod_b200.runner.mesh_run` is the implementation (native TCP / shared-memory mesh instead of Gloo); this module keeps the
names programs and tools import from `horovod.runner.gloo_run`.
"""
from horovod_b200.runner.mesh_run import (  # noqa: F401
    create_run_env_vars, create_slot_env_vars, elastic_run, launch_static, mesh_run)


def gloo_run(settings, nics, env, server_ip, command):
    """Static launch of `command` on the hosts in `settings` (one process per slot)."""
    return mesh_run(settings, nics, env, server_ip, command)


def launch_gloo(command, exec_command, settings, nics, env, server_ip):
    return launch_static(command, exec_command, settings, nics, env, server_ip)


def gloo_run_elastic(settings, env, command_or_func, executable=None):
    """Elastic launch; the discovery / size limits come from the elastic settings object."""
    if callable(command_or_func):
        raise ValueError('pass functions through horovod_b200.run(func, ..., min_num_proc=..., max_num_proc=...): it ships the function to the '
                         'workers and then calls the elastic launcher with the task command')
    return elastic_run(settings, env, command_or_func, settings.discovery, settings.min_num_proc, settings.max_num_proc,
                       settings.elastic_timeout, settings.reset_limit, getattr(settings, 'cooldown_range', None))
