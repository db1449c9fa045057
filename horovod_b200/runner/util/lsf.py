"""LSF (IBM Spectrum) helpers: detect an allocation and read its hosts / GPUs (reference runner/util/lsf.py)."""
import os
import subprocess


class LSFUtils:
    _CSM_ALLOCATION_QUERY = '/opt/ibm/csm/bin/csm_allocation_query'
    _CSM_NODE_QUERY = '/opt/ibm/csm/bin/csm_node_attributes_query'
    _LSCPU_CMD = 'LANG=en_US.utf8 lscpu'
    _THREAD_KEY = 'Thread(s) per core'
    _csm_allocation_info = {}

    @staticmethod
    def using_lsf():
        """Returns True if LSF was used to start the current process."""
        return 'LSB_JOBID' in os.environ

    @staticmethod
    def get_compute_hosts():
        """Compute hosts of the allocation from LSB_MCPU_HOSTS ('launch 1 host1 N host2 N'), launch node dropped."""
        spec = os.environ.get('LSB_MCPU_HOSTS', '').split()
        hosts = [spec[i] for i in range(0, len(spec) - 1, 2)]
        slots = [int(spec[i]) for i in range(1, len(spec), 2)]
        if len(hosts) > 1 and slots and slots[0] == 1:
            hosts = hosts[1:]  # first entry is the launch node
        return sorted(set(hosts), key=hosts.index)

    @staticmethod
    def get_num_cores():
        try:
            return int(os.environ.get('LSB_MAX_NUM_PROCESSORS', os.cpu_count() or 1))
        except ValueError:
            return os.cpu_count() or 1

    @staticmethod
    def get_num_gpus():
        """GPUs per node: CUDA_VISIBLE_DEVICES if set, else nvidia-smi."""
        cvd = os.environ.get('CUDA_VISIBLE_DEVICES')
        if cvd:
            return len([d for d in cvd.split(',') if d.strip()])
        try:
            out = subprocess.run(['nvidia-smi', '-L'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=20).stdout.decode()
            return max(1, len([l for l in out.splitlines() if l.startswith('GPU ')]))
        except Exception:
            return 1

    @staticmethod
    def get_num_processes():
        return len(LSFUtils.get_compute_hosts()) * LSFUtils.get_num_gpus()

    @staticmethod
    def get_num_threads():
        try:
            out = subprocess.run(LSFUtils._LSCPU_CMD, shell=True, stdout=subprocess.PIPE, timeout=20).stdout.decode()
            for line in out.splitlines():
                if line.startswith(LSFUtils._THREAD_KEY):
                    return int(line.split(':')[1].strip())
        except Exception:
            pass
        return 1
