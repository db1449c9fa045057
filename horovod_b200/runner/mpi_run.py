"""Launch through an MPI implementation's `mpirun` (Open MPI, Spectrum MPI, MPICH, Intel MPI).

Role parity: horovod/runner/mpi_run.py (implementation detection :60-126, command construction :129-259).
Workers started this way read OMPI_COMM_WORLD_RANK/SIZE (or PMI_RANK/SIZE) in hvd.init() and still rendezvous
through the launcher's HTTP store, so the native TCP/shm transport is used for negotiation and CPU tensors: MPI is
only the process launcher here.
"""
import copy
import os
import shlex
import sys

from horovod_b200.runner.common.util import env as env_util
from horovod_b200.runner.common.util import hosts, safe_shell_exec, tiny_shell_exec
from horovod_b200.runner.http.http_server import RendezvousServer
from horovod_b200.runner.util import network

# MPI implementations
_OMPI_IMPL = 'OpenMPI'
_SMPI_IMPL = 'SpectrumMPI'
_MPICH_IMPL = 'MPICH'
_IMPI_IMPL = 'IntelMPI'
_UNKNOWN_IMPL = 'Unknown'
_MISSING_IMPL = 'Missing'

# Open MPI Flags
_OMPI_FLAGS = ['-mca pml ob1', '-mca btl ^openib']
# Spectrum MPI Flags
_SMPI_FLAGS = []
_SMPI_FLAGS_TCP = ['-tcp']
# MPICH Flags
_MPICH_FLAGS = []
# Intel MPI Flags
_IMPI_FLAGS = []

# Threshold for large cluster MPI issues:
_LARGE_CLUSTER_THRESHOLD = 64
# No process binding args
_NO_BINDING_ARGS = ['-bind-to none', '-map-by slot']
# Process socket binding args
_SOCKET_BINDING_ARGS = ['-bind-to socket', '-map-by socket', '-rank-by core']

# MPI not found error message
_MPI_NOT_FOUND_ERROR_MSG = ('hvdrun does not find an installed MPI.\n\n'
                            'Choose one of:\n'
                            '1. Install Open MPI 4.0.0+ or IBM Spectrum MPI or MPICH and re-run.\n'
                            '2. Use the built-in native mesh controller (default; no MPI required).')


def mpi_available(env=None):
    return _get_mpi_implementation(env) not in {_UNKNOWN_IMPL, _MISSING_IMPL}


def is_open_mpi(env=None):
    return _get_mpi_implementation(env) == _OMPI_IMPL


def is_spectrum_mpi(env=None):
    return _get_mpi_implementation(env) == _SMPI_IMPL


def is_mpich(env=None):
    return _get_mpi_implementation(env) == _MPICH_IMPL


def is_intel_mpi(env=None):
    return _get_mpi_implementation(env) == _IMPI_IMPL


def _get_mpi_implementation(env=None):
    """Detects the available MPI implementation by invoking `mpirun --version`."""
    command = 'mpirun --version'
    res = tiny_shell_exec.execute(command)
    if res is None:
        return _MISSING_IMPL
    (output, exit_code) = res
    if exit_code == 0:
        if 'Open MPI' in output or 'OpenRTE' in output:
            return _OMPI_IMPL
        if 'IBM Spectrum MPI' in output:
            return _SMPI_IMPL
        if 'MPICH' in output or 'HYDRA' in output:
            return _MPICH_IMPL
        if 'Intel(R) MPI' in output:
            return _IMPI_IMPL
        print('Unknown MPI implementation given in output of mpirun --version:', file=sys.stderr)
        print(output, file=sys.stderr)
        return _UNKNOWN_IMPL
    return _MISSING_IMPL


def _get_mpi_implementation_flags(tcp_flag, env=None):
    if is_open_mpi(env):
        return list(_OMPI_FLAGS), list(_NO_BINDING_ARGS), _OMPI_IMPL
    if is_spectrum_mpi(env):
        return (list(_SMPI_FLAGS_TCP) if tcp_flag else list(_SMPI_FLAGS)), list(_SOCKET_BINDING_ARGS), _SMPI_IMPL
    if is_mpich(env):
        return list(_MPICH_FLAGS), [], _MPICH_IMPL
    if is_intel_mpi(env):
        return list(_IMPI_FLAGS), [], _IMPI_IMPL
    return None, None, None


def build_mpi_command(settings, nics, env, command, impl_flags, binding_args, mpi_impl, rendezvous_addr=None,
                      rendezvous_port=None):
    """Pure function (unit-testable): the full mpirun command line as a string."""
    impl_flags = list(impl_flags)
    binding_args = settings.binding_args if settings.binding_args and mpi_impl != _IMPI_IMPL else ' '.join(binding_args)
    basic_args = '-l' if mpi_impl in (_MPICH_IMPL, _IMPI_IMPL) else '--allow-run-as-root --tag-output'
    output = []
    if settings.output_filename:
        output.append('-outfile-pattern' if mpi_impl in (_MPICH_IMPL, _IMPI_IMPL) else '--output-filename')
        output.append(settings.output_filename)
    env_list = '' if mpi_impl in (_MPICH_IMPL, _IMPI_IMPL) else ' '.join(
        '-x %s' % key for key in sorted(env.keys()) if env_util.is_exportable(key))
    host_names, _ = hosts.parse_hosts_and_slots(settings.hosts)
    if mpi_impl == _IMPI_IMPL:
        hosts_arg = ''
    elif mpi_impl == _MPICH_IMPL:
        hosts_arg = '-hosts {hosts}'.format(hosts=','.join(host_names))
    else:
        hosts_arg = '-{opt} {hosts}'.format(opt='H', hosts=settings.hosts)
    if len(host_names) >= _LARGE_CLUSTER_THRESHOLD and mpi_impl == _OMPI_IMPL:
        impl_flags.append('-mca plm_rsh_no_tree_spawn true')
        impl_flags.append('-mca plm_rsh_num_concurrent {}'.format(len(host_names)))
    # if user does not specify any hosts, mpirun by default uses local host: no need to specify NIC
    nic_args = ''
    if nics and mpi_impl == _OMPI_IMPL:
        nic_args = '-mca btl_tcp_if_include {nics} -x NCCL_SOCKET_IFNAME={nics}'.format(nics=','.join(nics))
    ssh_args = []
    if settings.ssh_port:
        ssh_args += [f'-p {settings.ssh_port}']
    if settings.ssh_identity_file:
        ssh_args += [f'-i {settings.ssh_identity_file}']
    ssh_arg = ''
    if ssh_args:
        joined = ' '.join(ssh_args)
        ssh_arg = f'-bootstrap=ssh -bootstrap-exec-args "{joined}"' if mpi_impl in (_MPICH_IMPL, _IMPI_IMPL) \
            else f'-mca plm_rsh_args "{joined}"'
    rdzv = ''
    if rendezvous_addr:
        if mpi_impl in (_MPICH_IMPL, _IMPI_IMPL):
            rdzv = f'-genv HOROVOD_GLOO_RENDEZVOUS_ADDR {rendezvous_addr} -genv HOROVOD_GLOO_RENDEZVOUS_PORT {rendezvous_port}'
        else:
            rdzv = f'-x HOROVOD_GLOO_RENDEZVOUS_ADDR={rendezvous_addr} -x HOROVOD_GLOO_RENDEZVOUS_PORT={rendezvous_port}'
    if isinstance(command, (list, tuple)):
        command = ' '.join(shlex.quote(par) for par in command)
    mpirun_command = ('mpirun {basic_args} -np {num_proc} {ppn_arg}{hosts_arg} {binding_args} {mpi_args} {ssh_args} '
                      '{nic_args} {output_filename_arg} {env} {rdzv} {extra_mpi_args} {command}'
                      .format(basic_args=basic_args, num_proc=settings.num_proc,
                              ppn_arg='', hosts_arg=hosts_arg, binding_args=binding_args, mpi_args=' '.join(impl_flags),
                              ssh_args=ssh_arg, nic_args=nic_args, output_filename_arg=' '.join(output), env=env_list,
                              rdzv=rdzv, extra_mpi_args=settings.extra_mpi_args if settings.extra_mpi_args else '',
                              command=command))
    return ' '.join(mpirun_command.split())


def mpi_run(settings, nics, env, command, stdout=None, stderr=None):
    """Runs mpirun; the rendezvous KV server of this launcher process serves the workers' bootstrap."""
    mpi_impl_flags, impl_binding_args, mpi = _get_mpi_implementation_flags(settings.tcp_flag, env=env)
    if mpi_impl_flags is None:
        raise Exception(_MPI_NOT_FOUND_ERROR_MSG)
    rendezvous = RendezvousServer(settings.verbose)
    port = rendezvous.start_server()
    rendezvous.init(hosts.get_host_assignments(hosts.parse_hosts(settings.hosts), settings.num_proc))
    addr = network.get_driver_ip(nics)
    env = copy.copy(env)
    env['HOROVOD_GLOO_RENDEZVOUS_ADDR'] = addr
    env['HOROVOD_GLOO_RENDEZVOUS_PORT'] = str(port)
    mpirun_command = build_mpi_command(settings, nics, env, command, mpi_impl_flags, impl_binding_args, mpi, addr, port)
    if settings.verbose >= 2:
        print(mpirun_command)
    # we need the driver's PATH and PYTHONPATH in env to run mpirun,
    # env for mpirun is different to env encoded in mpirun_command
    for var in ['PATH', 'PYTHONPATH']:
        if var not in env and var in os.environ:
            env[var] = os.environ[var]
    try:
        exit_code = safe_shell_exec.execute(mpirun_command, env=env, stdout=stdout, stderr=stderr)
    finally:
        rendezvous.stop()
    if exit_code != 0:
        raise RuntimeError('mpirun failed with exit code {exit_code}'.format(exit_code=exit_code))
