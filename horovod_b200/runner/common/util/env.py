"""Environment plumbing between launcher and workers (role parity: horovod/runner/common/util/env.py)."""
import os
import re

LOG_LEVEL_STR = ['FATAL', 'ERROR', 'WARNING', 'INFO', 'DEBUG', 'TRACE']

# never forwarded to workers: shell function exports (break `env VAR=...` quoting), the previous directory, the ssh
# session of the person who launched the job, and any secret key (the job's own key is passed explicitly)
_NOT_FORWARDED = re.compile(r'^(BASH_FUNC_.*|OLDPWD|SSH_.*|.*_SECRET_KEY)$')
IGNORE_REGEXES = {'BASH_FUNC_.*', 'OLDPWD', 'SSH_.*', '.*_SECRET_KEY'}

# (rank variable, size variable) in lookup order: hvdrun, Open MPI, PMI (MPICH / Intel MPI / Slurm), torchrun
_RANK_SIZE_VARS = (('HOROVOD_RANK', 'HOROVOD_SIZE'), ('OMPI_COMM_WORLD_RANK', 'OMPI_COMM_WORLD_SIZE'),
                   ('PMI_RANK', 'PMI_SIZE'), ('RANK', 'WORLD_SIZE'))


def is_exportable(name):
    return _NOT_FORWARDED.match(name) is None


def get_env_rank_and_size(environ=None):
    """(rank, size) from whichever launcher started this process; (0, 1) when none did."""
    environ = os.environ if environ is None else environ
    for rank_var, size_var in _RANK_SIZE_VARS:
        if rank_var in environ and size_var in environ:
            return int(environ[rank_var]), int(environ[size_var])
    return 0, 1


# Kubeflow's MPI operator gives launcher pods no sshd: remote commands go through this script (`kubectl exec` underneath), and
# the operator announces it as Open MPI's rsh agent
KUBEFLOW_MPI_EXEC = '/etc/mpi/kubexec.sh'


def is_kubeflow_mpi(environ=None):
    environ = os.environ if environ is None else environ
    return environ.get('OMPI_MCA_plm_rsh_agent') == KUBEFLOW_MPI_EXEC
