"""Launch through `jsrun` inside an IBM LSF allocation, with an explicit resource file that gives every rank one GPU
and its share of cores.  Role parity: horovod/runner/js_run.py."""
import os
import shlex
import tempfile

from horovod_b200.runner.common.util import safe_shell_exec
from horovod_b200.runner.util import lsf


def is_jsrun_installed():
    """Returns True if jsrun is installed."""
    for p in os.environ.get('PATH', '').split(os.pathsep):
        if os.path.isfile(os.path.join(p, 'jsrun')):
            return True
    return False


def generate_jsrun_rankfile(settings, path=None):
    """Writes the explicit resource file (ERF): one `rank: N: { host: H; cpu: {a-b} ; gpu: * ; mem: * }` per rank."""
    cpu_per_gpu = (lsf.LSFUtils.get_num_cores() * lsf.LSFUtils.get_num_threads()) // max(1, lsf.LSFUtils.get_num_gpus())
    host_list = (x.split(':') for x in settings.hosts.split(','))
    # Verify and truncate host list if necessary
    validated_list = []
    remaining_slots = settings.num_proc
    for host, slots in host_list:
        slots = int(slots)
        if slots > lsf.LSFUtils.get_num_gpus():
            raise ValueError('Invalid host input, slot count for host \'{host}:{slots}\' is greater than number of GPUs per '
                             'host \'{gpus}\'.'.format(host=host, slots=slots, gpus=lsf.LSFUtils.get_num_gpus()))
        needed_slots = min(slots, remaining_slots)
        validated_list.append((host, needed_slots))
        remaining_slots -= needed_slots
        if remaining_slots == 0:
            break
    if remaining_slots != 0:
        raise ValueError('Not enough slots on the hosts to fulfill the {slots} requested.'.format(slots=settings.num_proc))
    # Generate rankfile
    path = tempfile.mktemp() if path is None else path
    with open(path, 'w') as tmp:
        tmp.write('overlapping_rs: allow\n')
        tmp.write('cpu_index_using: logical\n')
        rank = 0
        for host, slots in validated_list:
            cpu_val = 0
            tmp.write('\n')
            for s in range(slots):
                tmp.write('rank: {rank}: {{ hostname: {host}; cpu: {{{scpu}-{ecpu}}} ; gpu: * ; mem: * }}\n'.format(
                    rank=rank, host=host, scpu=cpu_val, ecpu=cpu_val + cpu_per_gpu - 1))
                rank += 1
                cpu_val += cpu_per_gpu
    return path


def build_jsrun_command(settings, env, command, rankfile):
    if isinstance(command, (list, tuple)):
        command = ' '.join(shlex.quote(par) for par in command)
    smpiargs = '-gpu' if not settings.extra_mpi_args else '-gpu ' + settings.extra_mpi_args
    binding = settings.binding_args if settings.binding_args else ''
    out = f'--stdio_stdout {settings.output_filename} --stdio_stderr {settings.output_filename}' if settings.output_filename else ''
    return ' '.join(f'jsrun --erf_input {rankfile} {out} --smpiargs {shlex.quote(smpiargs)} {binding} {command}'.split())


def js_run(settings, nics, env, command, stdout=None, stderr=None):
    if not is_jsrun_installed():
        raise Exception('hvdrun did not find the jsrun command.')
    rankfile = generate_jsrun_rankfile(settings)
    jsrun_command = build_jsrun_command(settings, env, command, rankfile)
    if settings.verbose >= 2:
        print(jsrun_command)
    exit_code = safe_shell_exec.execute(jsrun_command, env=env, stdout=stdout, stderr=stderr)
    if exit_code != 0:
        raise RuntimeError('jsrun failed with exit code {exit_code}'.format(exit_code=exit_code))
