"""Launcher configuration: YAML config file -> argparse namespace -> worker environment variables.

Role parity: horovod/runner/common/util/config_parser.py (set_args_from_config :66-123, set_env_from_args :160-205).
"""
# environment variable names understood by the native runtime
HOROVOD_FUSION_THRESHOLD = 'HOROVOD_FUSION_THRESHOLD'
HOROVOD_CYCLE_TIME = 'HOROVOD_CYCLE_TIME'
HOROVOD_CACHE_CAPACITY = 'HOROVOD_CACHE_CAPACITY'
HOROVOD_HIERARCHICAL_ALLREDUCE = 'HOROVOD_HIERARCHICAL_ALLREDUCE'
HOROVOD_HIERARCHICAL_ALLGATHER = 'HOROVOD_HIERARCHICAL_ALLGATHER'
HOROVOD_AUTOTUNE = 'HOROVOD_AUTOTUNE'
HOROVOD_AUTOTUNE_LOG = 'HOROVOD_AUTOTUNE_LOG'
HOROVOD_AUTOTUNE_WARMUP_SAMPLES = 'HOROVOD_AUTOTUNE_WARMUP_SAMPLES'
HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE = 'HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE'
HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES = 'HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES'
HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE = 'HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE'
HOROVOD_TIMELINE = 'HOROVOD_TIMELINE'
HOROVOD_TIMELINE_MARK_CYCLES = 'HOROVOD_TIMELINE_MARK_CYCLES'
HOROVOD_STALL_CHECK_DISABLE = 'HOROVOD_STALL_CHECK_DISABLE'
HOROVOD_STALL_CHECK_TIME_SECONDS = 'HOROVOD_STALL_CHECK_TIME_SECONDS'
HOROVOD_STALL_SHUTDOWN_TIME_SECONDS = 'HOROVOD_STALL_SHUTDOWN_TIME_SECONDS'
HOROVOD_MPI_THREADS_DISABLE = 'HOROVOD_MPI_THREADS_DISABLE'
HOROVOD_NUM_NCCL_STREAMS = 'HOROVOD_NUM_NCCL_STREAMS'
HOROVOD_THREAD_AFFINITY = 'HOROVOD_THREAD_AFFINITY'
HOROVOD_LOG_LEVEL = 'HOROVOD_LOG_LEVEL'
HOROVOD_LOG_HIDE_TIME = 'HOROVOD_LOG_HIDE_TIME'
HVD_GPU_BACKEND = 'HVD_GPU_BACKEND'
HVD_ALLREDUCE_VARIANT = 'HVD_ALLREDUCE_VARIANT'
HVD_WIRE_DTYPE = 'HVD_WIRE_DTYPE'
HVD_COMM_CTAS = 'HVD_COMM_CTAS'
LOG_LEVELS = ['TRACE', 'DEBUG', 'INFO', 'WARNING', 'ERROR', 'FATAL']


def _set_arg_from_config(args, arg_base_name, override_args, config, arg_prefix=''):
    arg_name = arg_prefix + arg_base_name
    if arg_name in override_args:
        return  # the command line wins over the config file
    value = config.get(arg_base_name)
    if value is not None:
        setattr(args, arg_name, value)


def set_args_from_config(args, config, override_args):
    """Applies a parsed YAML config (see docs/launcher.md) onto the argparse namespace."""
    # Controller
    controller = config.get('controller')
    if controller:
        for c in ('gloo', 'mpi', 'js'):
            if controller.lower() == c and f'use_{c}' not in override_args:
                setattr(args, f'use_{c}', True)
    params = config.get('params')
    if params:
        for name in ('fusion_threshold_mb', 'cycle_time_ms', 'cache_capacity', 'hierarchical_allreduce',
                     'hierarchical_allgather', 'thread_affinity', 'num_nccl_streams'):
            _set_arg_from_config(args, name, override_args, params)
    autotune = config.get('autotune')
    if autotune:
        if 'autotune' not in override_args:
            args.autotune = autotune.get('enabled', False)
        for name in ('log_file', 'warmup_samples', 'steps_per_sample', 'bayes_opt_max_samples', 'gaussian_process_noise'):
            _set_arg_from_config(args, name, override_args, autotune, arg_prefix='autotune_')
    timeline = config.get('timeline')
    if timeline:
        _set_arg_from_config(args, 'filename', override_args, timeline, arg_prefix='timeline_')
        _set_arg_from_config(args, 'mark_cycles', override_args, timeline, arg_prefix='timeline_')
    stall_check = config.get('stall_check')
    if stall_check:
        if 'no_stall_check' not in override_args:
            args.no_stall_check = not stall_check.get('enabled', True)
        _set_arg_from_config(args, 'warning_time_seconds', override_args, stall_check, arg_prefix='stall_check_')
        _set_arg_from_config(args, 'shutdown_time_seconds', override_args, stall_check, arg_prefix='stall_check_')
    library_options = config.get('library_options')
    if library_options:
        _set_arg_from_config(args, 'mpi_threads_disable', override_args, library_options)
        _set_arg_from_config(args, 'gpu_backend', override_args, library_options)
        _set_arg_from_config(args, 'allreduce_variant', override_args, library_options)
        _set_arg_from_config(args, 'wire_dtype', override_args, library_options)
    logging = config.get('logging')
    if logging:
        _set_arg_from_config(args, 'level', override_args, logging, arg_prefix='log_')
        _set_arg_from_config(args, 'hide_timestamp', override_args, logging, arg_prefix='log_')


def _validate_arg_nonnegative(args, arg_name):
    value = getattr(args, arg_name, None)
    if value is not None and value < 0:
        raise ValueError('{}={} must be >= 0'.format(arg_name, value))


def validate_config_args(args):
    for name in ('fusion_threshold_mb', 'cycle_time_ms', 'cache_capacity', 'autotune_warmup_samples',
                 'autotune_steps_per_sample', 'autotune_bayes_opt_max_samples', 'autotune_gaussian_process_noise',
                 'stall_check_warning_time_seconds', 'stall_check_shutdown_time_seconds', 'num_nccl_streams'):
        _validate_arg_nonnegative(args, name)
    noise = getattr(args, 'autotune_gaussian_process_noise', None)
    if noise is not None and noise > 1:
        raise ValueError('autotune_gaussian_process_noise={} must be in [0, 1]'.format(noise))


def _add_arg_to_env(env, env_key, arg_value, transform_fn=None):
    if arg_value is not None:
        value = arg_value
        if transform_fn:
            value = transform_fn(value)
        env[env_key] = str(value)


def set_env_from_args(env, args):
    def identity(value):
        return 1 if value else 0

    # Params
    _add_arg_to_env(env, HOROVOD_FUSION_THRESHOLD, getattr(args, 'fusion_threshold_mb', None), lambda v: int(v * 1024 * 1024))
    _add_arg_to_env(env, HOROVOD_CYCLE_TIME, getattr(args, 'cycle_time_ms', None))
    _add_arg_to_env(env, HOROVOD_CACHE_CAPACITY, getattr(args, 'cache_capacity', None))
    _add_arg_to_env(env, HOROVOD_HIERARCHICAL_ALLREDUCE, getattr(args, 'hierarchical_allreduce', None), identity)
    _add_arg_to_env(env, HOROVOD_HIERARCHICAL_ALLGATHER, getattr(args, 'hierarchical_allgather', None), identity)
    _add_arg_to_env(env, HOROVOD_THREAD_AFFINITY, getattr(args, 'thread_affinity', None))
    _add_arg_to_env(env, HOROVOD_NUM_NCCL_STREAMS, getattr(args, 'num_nccl_streams', None))
    # Autotune
    if getattr(args, 'autotune', None):
        _add_arg_to_env(env, HOROVOD_AUTOTUNE, args.autotune, identity)
        _add_arg_to_env(env, HOROVOD_AUTOTUNE_LOG, getattr(args, 'autotune_log_file', None))
        _add_arg_to_env(env, HOROVOD_AUTOTUNE_WARMUP_SAMPLES, getattr(args, 'autotune_warmup_samples', None))
        _add_arg_to_env(env, HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE, getattr(args, 'autotune_steps_per_sample', None))
        _add_arg_to_env(env, HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES, getattr(args, 'autotune_bayes_opt_max_samples', None))
        _add_arg_to_env(env, HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE, getattr(args, 'autotune_gaussian_process_noise', None))
    # Timeline
    if getattr(args, 'timeline_filename', None):
        _add_arg_to_env(env, HOROVOD_TIMELINE, args.timeline_filename)
        _add_arg_to_env(env, HOROVOD_TIMELINE_MARK_CYCLES, getattr(args, 'timeline_mark_cycles', None), identity)
    # Stall check
    _add_arg_to_env(env, HOROVOD_STALL_CHECK_DISABLE, getattr(args, 'no_stall_check', None), identity)
    _add_arg_to_env(env, HOROVOD_STALL_CHECK_TIME_SECONDS, getattr(args, 'stall_check_warning_time_seconds', None))
    _add_arg_to_env(env, HOROVOD_STALL_SHUTDOWN_TIME_SECONDS, getattr(args, 'stall_check_shutdown_time_seconds', None))
    # Library options
    _add_arg_to_env(env, HOROVOD_MPI_THREADS_DISABLE, getattr(args, 'mpi_threads_disable', None), identity)
    _add_arg_to_env(env, HVD_GPU_BACKEND, getattr(args, 'gpu_backend', None))
    _add_arg_to_env(env, HVD_ALLREDUCE_VARIANT, getattr(args, 'allreduce_variant', None))
    _add_arg_to_env(env, HVD_WIRE_DTYPE, getattr(args, 'wire_dtype', None))
    _add_arg_to_env(env, HVD_COMM_CTAS, getattr(args, 'comm_ctas', None))
    # Logging
    _add_arg_to_env(env, HOROVOD_LOG_LEVEL, getattr(args, 'log_level', None))
    _add_arg_to_env(env, HOROVOD_LOG_HIDE_TIME, getattr(args, 'log_hide_timestamp', None), identity)
    return env
