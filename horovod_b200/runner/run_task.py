"""Worker side of run-func mode: fetch the pickled function from the launcher's KV store, run it, put the result back
(reference horovod/runner/run_task.py + task_fn.py)."""
import sys

import cloudpickle

from horovod_b200.runner.common.util.env import get_env_rank_and_size
from horovod_b200.runner.http.http_client import put_data_into_kvstore, read_data_from_kvstore


def main(driver_addr, run_func_server_port):
    func = cloudpickle.loads(read_data_from_kvstore(driver_addr, run_func_server_port, 'runfunc', 'func'))
    try:
        ret_val = func()
    except BaseException as e:
        sys.stderr.write("User function raise error: {error}".format(error=str(e)))
        raise e
    rank, size = get_env_rank_and_size()
    put_data_into_kvstore(driver_addr, run_func_server_port, 'runfunc_result', str(rank), cloudpickle.dumps(ret_val))


if __name__ == '__main__':
    _, driver_addr, run_func_server_port_str = sys.argv
    main(driver_addr, int(run_func_server_port_str))
