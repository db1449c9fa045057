"""`python -m horovod_b200.runner.run_task <driver addr> <port>` — what every worker executes in run-func mode
(`horovod_b200.run(fn, ...)`): download the cloudpickled function from the launcher's KV server, call it, upload the
cloudpickled return value under this worker's rank.  Role parity: horovod/runner/run_task.py + task_fn.py."""
import sys
import traceback

import cloudpickle

from horovod_b200.runner.common.util.env import get_env_rank_and_size
from horovod_b200.runner.http.http_client import put_data_into_kvstore, read_data_from_kvstore

FUNC_SCOPE, RESULT_SCOPE = 'runfunc', 'runfunc_result'


def run_remote_function(addr, port):
    payload = read_data_from_kvstore(addr, port, FUNC_SCOPE, 'func')
    fn = cloudpickle.loads(payload)
    try:
        result = fn()
    except BaseException:
        sys.stderr.write('User function raised an error:\n' + traceback.format_exc())
        raise
    my_rank = get_env_rank_and_size()[0]
    put_data_into_kvstore(addr, port, RESULT_SCOPE, str(my_rank), cloudpickle.dumps(result))


main = run_remote_function  # name used by the reference's callers

if __name__ == '__main__':
    if len(sys.argv) != 3:
        sys.exit('usage: python -m horovod_b200.runner.run_task <driver addr> <run-func server port>')
    run_remote_function(sys.argv[1], int(sys.argv[2]))
