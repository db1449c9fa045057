"""Programmatic launcher API: `horovod_b200.run(fn, args=(...), num_proc=4)` runs a Python function on every rank and returns
the list of results (reference horovod/runner/__init__.py:95-247)."""


class _HorovodArgs(object):
    def __init__(self):
        self.np = None
        self.check_build = None
        self.ssh_port = None
        self.ssh_identity_file = None
        self.disable_cache = None
        self.start_timeout = None
        self.nics = None
        self.output_filename = None
        self.verbose = None
        self.command = None
        self.run_func = None
        self.config_file = None
        self.prefix_output_with_timestamp = False
        # tuneable parameter arguments
        self.fusion_threshold_mb = None
        self.cycle_time_ms = None
        self.cache_capacity = None
        self.hierarchical_allreduce = None
        self.hierarchical_allgather = None
        self.thread_affinity = None
        self.num_nccl_streams = None
        # autotune arguments
        self.autotune = None
        self.autotune_log_file = None
        self.autotune_warmup_samples = None
        self.autotune_steps_per_sample = None
        self.autotune_bayes_opt_max_samples = None
        self.autotune_gaussian_process_noise = None
        # elastic arguments
        self.min_num_proc = None
        self.max_num_proc = None
        self.slots = None
        self.elastic_timeout = None
        self.reset_limit = None
        self.cooldown_range = None
        # timeline arguments
        self.timeline_filename = None
        self.timeline_mark_cycles = None
        # stall check arguments
        self.no_stall_check = None
        self.stall_check_warning_time_seconds = None
        self.stall_check_shutdown_time_seconds = None
        # library arguments
        self.mpi_threads_disable = None
        self.mpi_args = None
        self.tcp_flag = None
        self.binding_args = None
        self.gpu_backend = None
        self.allreduce_variant = None
        self.wire_dtype = None
        self.comm_ctas = None
        # logging arguments
        self.log_level = None
        self.log_hide_timestamp = None
        # host arguments
        self.hosts = None
        self.hostfile = None
        self.host_discovery_script = None
        # controller arguments
        self.use_gloo = None
        self.use_mpi = None
        self.use_jsrun = None

    @property
    def num_proc(self):
        return self.np

    @num_proc.setter
    def num_proc(self, v):
        self.np = v


def _pickle_by_value_if_not_importable(func):
    """Workers unpickle the function in a fresh interpreter.  A function defined in a script or a test module that is not
    on the workers' import path would be pickled by reference and fail there with ModuleNotFoundError; such modules are
    shipped by value instead (what cloudpickle already does for `__main__`)."""
    import sys
    import cloudpickle
    name = getattr(func, '__module__', None)
    mod = sys.modules.get(name)
    if not name or name == '__main__' or mod is None or name.split('.')[0] == 'horovod_b200':
        return
    path = getattr(mod, '__file__', '') or ''
    in_site = any(part in path for part in ('site-packages', 'dist-packages'))
    if not in_site and hasattr(cloudpickle, 'register_pickle_by_value'):
        try:
            cloudpickle.register_pickle_by_value(mod)
        except Exception:  # noqa: BLE001 - best effort; the by-reference pickle still works when the module is importable
            pass


def run(func, args=(), kwargs=None, num_proc=None, min_num_proc=None, max_num_proc=None, slots=None, reset_limit=None,
        cooldown_range=None, hosts=None, hostfile=None, host_discovery_script=None, start_timeout=None, ssh_port=None,
        ssh_identity_file=None, disable_cache=None, output_filename=None, verbose=None, use_gloo=None, use_mpi=None, mpi_args=None,
        network_interface=None, network_interfaces=None, executable=None, np=None, min_np=None, max_np=None):
    """Launches `func(*args, **kwargs)` on `num_proc` processes and returns the per-rank results, rank-ordered (reference
    horovod/runner/__init__.py `run` :95-260).  `host_discovery_script` (or `min_num_proc`) makes the job elastic.  `np`,
    `min_np`, `max_np` and `network_interface` are the deprecated spellings of `num_proc`, `min_num_proc`, `max_num_proc` and
    `network_interfaces`."""
    import warnings
    from horovod_b200.runner.launch import _run

    def pick(new, old, new_name, old_name, default=None):
        if old is not None:
            if new is not None and new != old:
                raise ValueError('%s and %s were both given with different values; %s is deprecated, use %s' % (new_name, old_name, old_name, new_name))
            warnings.warn('%s is deprecated, use %s instead' % (old_name, new_name), DeprecationWarning, stacklevel=3)
            return old
        return default if new is None else new
    np = pick(num_proc, np, 'num_proc', 'np', default=1)
    min_np = pick(min_num_proc, min_np, 'min_num_proc', 'min_np')
    max_np = pick(max_num_proc, max_np, 'max_num_proc', 'max_np')
    if network_interface is not None:
        if network_interfaces is not None:
            raise ValueError('network_interface and network_interfaces were both given; network_interface is deprecated')
        warnings.warn('network_interface is deprecated, use network_interfaces instead', DeprecationWarning, stacklevel=2)
        network_interfaces = network_interface
    if isinstance(network_interfaces, (list, tuple, set)):
        network_interfaces = ','.join(sorted(network_interfaces))
    if kwargs is None:
        kwargs = {}

    def wrapped_func():
        return func(*args, **kwargs)

    _pickle_by_value_if_not_importable(func)

    hargs = _HorovodArgs()
    hargs.np = np
    hargs.min_num_proc = min_np
    hargs.max_num_proc = max_np
    hargs.slots = slots
    hargs.reset_limit = reset_limit
    hargs.cooldown_range = cooldown_range
    hargs.hosts = hosts
    hargs.hostfile = hostfile
    hargs.host_discovery_script = host_discovery_script
    hargs.start_timeout = start_timeout
    hargs.ssh_port = ssh_port
    hargs.ssh_identity_file = ssh_identity_file
    hargs.mpi_args = mpi_args
    hargs.disable_cache = disable_cache
    hargs.output_filename = output_filename
    hargs.verbose = verbose
    hargs.use_gloo = use_gloo
    hargs.use_mpi = use_mpi
    hargs.nics = network_interfaces
    hargs.run_func = wrapped_func
    hargs.executable = executable
    return _run(hargs)
