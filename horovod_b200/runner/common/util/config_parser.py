"""Launcher configuration: YAML config file -> argparse namespace -> worker environment variables.

Role parity: horovod/runner/common/util/config_parser.py (`set_args_from_config`, `validate_config_args`,
`set_env_from_args`).  Here everything is driven by ONE table (`OPTIONS`): an option is declared once with its argparse
attribute, its place in the YAML file (section, key), the environment variable it becomes and how the value is
converted — adding a knob is one line, and file / CLI / env cannot drift apart.
"""
from collections import namedtuple

LOG_LEVELS = ['TRACE', 'DEBUG', 'INFO', 'WARNING', 'ERROR', 'FATAL']


def _flag(v):
    return 1 if v else 0


def _only_when_set(v):
    return 1 if v else None


def _mb_to_bytes(v):
    return int(v * 1024 * 1024)


# attr: argparse dest; section/key: position in the YAML file; env: variable exported to the workers; conv: value -> env
# text; needs: attr that must be truthy for the variable to be exported; nonneg: validated as >= 0
Option = namedtuple('Option', 'attr section key env conv needs nonneg')


def _opt(attr, section, key, env, conv=None, needs=None, nonneg=False):
    return Option(attr, section, key, env, conv, needs, nonneg)


OPTIONS = [
    _opt('fusion_threshold_mb', 'params', 'fusion_threshold_mb', 'HOROVOD_FUSION_THRESHOLD', _mb_to_bytes, nonneg=True),
    _opt('cycle_time_ms', 'params', 'cycle_time_ms', 'HOROVOD_CYCLE_TIME', nonneg=True),
    _opt('cache_capacity', 'params', 'cache_capacity', 'HOROVOD_CACHE_CAPACITY', nonneg=True),
    _opt('hierarchical_allreduce', 'params', 'hierarchical_allreduce', 'HOROVOD_HIERARCHICAL_ALLREDUCE', _flag),
    _opt('hierarchical_allgather', 'params', 'hierarchical_allgather', 'HOROVOD_HIERARCHICAL_ALLGATHER', _flag),
    _opt('torus_allreduce', 'params', 'torus_allreduce', 'HOROVOD_TORUS_ALLREDUCE', _flag),
    _opt('thread_affinity', 'params', 'thread_affinity', 'HOROVOD_THREAD_AFFINITY'),
    _opt('num_nccl_streams', 'params', 'num_nccl_streams', 'HOROVOD_NUM_NCCL_STREAMS', nonneg=True),
    _opt('autotune', 'autotune', 'enabled', 'HOROVOD_AUTOTUNE', _flag, needs='autotune'),
    _opt('autotune_log_file', 'autotune', 'log_file', 'HOROVOD_AUTOTUNE_LOG', needs='autotune'),
    _opt('autotune_warmup_samples', 'autotune', 'warmup_samples', 'HOROVOD_AUTOTUNE_WARMUP_SAMPLES', needs='autotune', nonneg=True),
    _opt('autotune_steps_per_sample', 'autotune', 'steps_per_sample', 'HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE', needs='autotune', nonneg=True),
    _opt('autotune_bayes_opt_max_samples', 'autotune', 'bayes_opt_max_samples', 'HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES',
         needs='autotune', nonneg=True),
    _opt('autotune_gaussian_process_noise', 'autotune', 'gaussian_process_noise', 'HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE',
         needs='autotune', nonneg=True),
    _opt('timeline_filename', 'timeline', 'filename', 'HOROVOD_TIMELINE', needs='timeline_filename'),
    _opt('timeline_mark_cycles', 'timeline', 'mark_cycles', 'HOROVOD_TIMELINE_MARK_CYCLES', _flag, needs='timeline_filename'),
    _opt('no_stall_check', 'stall_check', None, 'HOROVOD_STALL_CHECK_DISABLE', _flag),   # YAML: stall_check.enabled (inverted)
    _opt('stall_check_warning_time_seconds', 'stall_check', 'warning_time_seconds', 'HOROVOD_STALL_CHECK_TIME_SECONDS', nonneg=True),
    _opt('stall_check_shutdown_time_seconds', 'stall_check', 'shutdown_time_seconds', 'HOROVOD_STALL_SHUTDOWN_TIME_SECONDS', nonneg=True),
    _opt('mpi_threads_disable', 'library_options', 'mpi_threads_disable', 'HOROVOD_MPI_THREADS_DISABLE', _flag),
    _opt('tcp_flag', 'library_options', 'tcp', 'NCCL_IB_DISABLE', _only_when_set),      # --tcp: no InfiniBand for NCCL either
    _opt('gloo_timeout_seconds', 'library_options', 'gloo_timeout_seconds', 'HOROVOD_GLOO_TIMEOUT_SECONDS', nonneg=True),
    _opt('gpu_backend', 'library_options', 'gpu_backend', 'HVD_GPU_BACKEND'),
    _opt('allreduce_variant', 'library_options', 'allreduce_variant', 'HVD_ALLREDUCE_VARIANT'),
    _opt('wire_dtype', 'library_options', 'wire_dtype', 'HVD_WIRE_DTYPE'),
    _opt('comm_ctas', 'library_options', 'comm_ctas', 'HVD_COMM_CTAS', nonneg=True),
    _opt('log_level', 'logging', 'level', 'HOROVOD_LOG_LEVEL'),
    _opt('log_hide_timestamp', 'logging', 'hide_timestamp', 'HOROVOD_LOG_HIDE_TIME', _flag),
]

# names other modules import
for _o in OPTIONS:
    globals()[_o.env] = _o.env
HOROVOD_NUM_NCCL_STREAMS = 'HOROVOD_NUM_NCCL_STREAMS'


def set_args_from_config(args, config, override_args):
    """Copies the values of a parsed YAML config onto the argparse namespace, except where the command line already set
    the option (`override_args` holds the dests given on the command line)."""
    controller = (config.get('controller') or '').lower()
    for name in ('gloo', 'mpi', 'js'):
        dest = 'use_' + name
        if controller == name and dest not in override_args:
            setattr(args, dest, True)
    for o in OPTIONS:
        section = config.get(o.section) or {}
        if o.attr in override_args or not section:
            continue
        if o.attr == 'no_stall_check':
            if 'enabled' in section:
                args.no_stall_check = not section['enabled']
        elif o.attr == 'autotune':
            args.autotune = bool(section.get('enabled', False))
        elif section.get(o.key) is not None:
            setattr(args, o.attr, section[o.key])


def validate_config_args(args):
    for o in OPTIONS:
        value = getattr(args, o.attr, None)
        if o.nonneg and value is not None and value < 0:
            raise ValueError('{}={} must be >= 0'.format(o.attr, value))
    noise = getattr(args, 'autotune_gaussian_process_noise', None)
    if noise is not None and noise > 1:
        raise ValueError('autotune_gaussian_process_noise={} must be in [0, 1]'.format(noise))


def set_env_from_args(env, args):
    """Adds the runtime's environment variables for every option that was given (file or command line)."""
    for o in OPTIONS:
        value = getattr(args, o.attr, None)
        if value is None or (o.needs and not getattr(args, o.needs, None)):
            continue
        value = o.conv(value) if o.conv else value
        if value is not None:
            env[o.env] = str(value)
    return env
