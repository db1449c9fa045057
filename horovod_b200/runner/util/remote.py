"""ssh command lines for running something on another host (role parity: horovod/runner/util/remote.py
`get_ssh_command` / `get_remote_command`)."""
import shlex

from horovod_b200.runner.common.util import env as env_util
from horovod_b200.runner.util.network import is_local_host

SSH_COMMAND_PREFIX = 'ssh -o PasswordAuthentication=no -o StrictHostKeyChecking=no'

SSH_BASE_OPTIONS = ('-o', 'PasswordAuthentication=no', '-o', 'StrictHostKeyChecking=no')


def ssh_argv(host, port=None, identity_file=None, timeout_s=None, extra_options=()):
    """['ssh', options..., host] — the remote command is appended by the caller (as ONE quoted argument)."""
    argv = ['ssh'] + list(SSH_BASE_OPTIONS)
    if timeout_s is not None:
        argv += ['-o', 'ConnectTimeout=%d' % int(timeout_s)]
    argv += list(extra_options)
    if port is not None:
        argv += ['-p', str(port)]
    if identity_file is not None:
        argv += ['-i', str(identity_file)]
    return argv + [host]


def get_ssh_command(local_command, host, port=None, identity_file=None, timeout_s=None):
    """Shell string that runs `local_command` on `host` through ssh (no password prompts, no host-key questions); inside a
    Kubeflow MPI-operator launcher pod (no sshd in the worker pods) through the operator's `kubexec.sh` instead."""
    if env_util.is_kubeflow_mpi():
        return '%s %s %s' % (env_util.KUBEFLOW_MPI_EXEC, host, shlex.quote(local_command))
    return ' '.join(ssh_argv(host, port, identity_file, timeout_s)) + ' ' + shlex.quote(local_command)


def get_remote_command(local_command, host, port=None, identity_file=None, timeout_s=None):
    """`local_command` itself when `host` is this machine, the ssh form otherwise."""
    if is_local_host(host):
        return local_command
    return get_ssh_command(local_command, host, port, identity_file, timeout_s)
