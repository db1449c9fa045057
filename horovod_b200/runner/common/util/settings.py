"""What the launcher knows about one job, passed to the run functions (role parity:
horovod/runner/common/util/settings.py: `BaseSettings`, `Settings`)."""

_FIELDS = dict(num_proc=None, verbose=0, ssh_port=None, ssh_identity_file=None, extra_mpi_args=None, tcp_flag=None,
               binding_args=None, key=None, start_timeout=None, output_filename=None, run_func_mode=None, nics=None,
               elastic=False, prefix_output_with_timestamp=False)


class BaseSettings:
    """Keyword-only bag of launcher options; unknown keywords are rejected so that typos do not pass silently.

    num_proc: processes to start; verbose: 0..2; ssh_port / ssh_identity_file: remote shell options; extra_mpi_args,
    tcp_flag, binding_args: forwarded to mpirun; key: the job's HMAC secret; start_timeout: a `Timeout`;
    output_filename: directory for per-rank logs; run_func_mode: launched through `horovod_b200.run(fn)`; nics: set of
    interface names; elastic: elastic job; prefix_output_with_timestamp: timestamp forwarded worker output."""

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(_FIELDS)
        if unknown:
            raise TypeError('unknown setting(s): ' + ', '.join(sorted(unknown)))
        for name, default in _FIELDS.items():
            setattr(self, name, kwargs.get(name, default))

    def __repr__(self):
        return '%s(%s)' % (type(self).__name__, ', '.join('%s=%r' % (k, getattr(self, k)) for k in _FIELDS if k != 'key'))


class Settings(BaseSettings):
    """Static job: additionally the `host:slots,...` string the job runs on."""

    def __init__(self, hosts=None, **kwargs):
        super().__init__(**kwargs)
        self.hosts = hosts
