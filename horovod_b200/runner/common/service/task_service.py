"""A service that runs ONE shell command on the machine it lives on, on behalf of a remote client (reference
horovod/runner/common/service/task_service.py: `BasicTaskService` :108-312, `BasicTaskClient` :314-380).

The reference's Spark integration starts one of these inside every Spark task and lets the driver (or mpirun's rsh agent) run
the worker command through it; here the Spark and Ray integrations call Python functions in their tasks directly
(`runner/cluster_job.py`), and this service is the general-purpose building block for everything that still has to start a
PROCESS somewhere it cannot ssh to: run a command with an environment, follow its output while it runs, abort it, wait for its
exit code, hand a result object back.

Every request is idempotent (the RPC layer retries): running the same command twice is refused, output is fetched by offset.
"""
import os
import threading
import time

from horovod_b200.runner.common.util import network, safe_shell_exec
from horovod_b200.runner.util.streams import Pipe


class RunCommandRequest(object):
    def __init__(self, command, env, capture_stdout=False, capture_stderr=False, prefix_output_with_timestamp=False):
        self.command, self.env = command, env
        self.capture_stdout, self.capture_stderr = capture_stdout, capture_stderr
        self.prefix_output_with_timestamp = prefix_output_with_timestamp


class StreamCommandOutputRequest(object):
    """Next piece of a captured stream, starting at byte `offset` (so a retried request returns the same bytes)."""
    stream = None

    def __init__(self, offset=0):
        self.offset = offset


class StreamCommandStdOutRequest(StreamCommandOutputRequest):
    stream = 'stdout'


class StreamCommandStdErrRequest(StreamCommandOutputRequest):
    stream = 'stderr'


class StreamCommandOutputResponse(object):
    def __init__(self, data, next_offset, finished):
        self.data, self.next_offset, self.finished = data, next_offset, finished


class CommandOutputNotCaptured(Exception):
    """The command was started without capturing the stream that is being asked for."""


class AbortCommandRequest(object):
    pass


class CommandExitCodeRequest(object):
    pass


class CommandExitCodeResponse(object):
    def __init__(self, terminated, exit_code):
        self.terminated, self.exit_code = terminated, exit_code


class WaitForCommandExitCodeRequest(object):
    def __init__(self, delay):
        """`delay`: longest time the service may hold the request before it answers "not yet"."""
        self.delay = delay


class WaitForCommandExitCodeResponse(object):
    def __init__(self, exit_code):
        self.exit_code = exit_code           # None = still running


class NotifyInitialRegistrationCompleteRequest(object):
    pass


class RegisterCodeResultRequest(object):
    def __init__(self, result):
        self.result = result


class _Capture(object):
    """Everything a stream has produced so far, readable by offset while the producer is still writing."""

    def __init__(self):
        self._chunks, self._size, self._closed = [], 0, False
        self._cond = threading.Condition()

    def write(self, data):
        if isinstance(data, str):
            data = data.encode('utf-8', 'replace')
        with self._cond:
            self._chunks.append(data)
            self._size += len(data)
            self._cond.notify_all()

    def flush(self):
        pass

    def close(self):
        with self._cond:
            self._closed = True
            self._cond.notify_all()

    def read_from(self, offset, wait_s=1.0, limit=1 << 20):
        with self._cond:
            if self._size <= offset and not self._closed:
                self._cond.wait(wait_s)
            data = b''.join(self._chunks)[offset:offset + limit]
            return data, offset + len(data), self._closed and offset + len(data) >= self._size


class BasicTaskService(network.BasicService):
    def __init__(self, name, index, key, nic=None, command_env=None, verbose=0):
        super(BasicTaskService, self).__init__(name, key, nic)
        self._index, self._verbose = index, verbose
        self._command_env = dict(command_env or {})
        self._lock = threading.Lock()
        self._initial_registration_complete = threading.Event()
        self._command_started = threading.Event()
        self._command_done = threading.Event()
        self._abort = threading.Event()
        self._exit_code = None
        self._capture = {'stdout': None, 'stderr': None}
        self._fn_result = None
        self._thread = None

    # -- the command ----------------------------------------------------------------------------------------------------------
    def _add_envs(self, env, extra_env):
        """`extra_env` wins; a value of None removes the variable."""
        for k, v in extra_env.items():
            if v is None:
                env.pop(k, None)
            else:
                env[k] = str(v)
        return env

    def _run_command(self, command, env, capture_stdout, capture_stderr, prefix_output_with_timestamp):
        out, err = self._capture['stdout'], self._capture['stderr']
        try:
            code = safe_shell_exec.execute(command, env=env, stdout=out, stderr=err, index=self._index, events=[self._abort],
                                           prefix_output_with_timestamp=prefix_output_with_timestamp)
        except Exception as e:  # noqa: BLE001 - the command could not even be started
            code = 127
            if err is not None:
                err.write('could not run %r: %s\n' % (command, e))
        for c in (out, err):
            if c is not None:
                c.close()
        with self._lock:
            self._exit_code = code
        self._command_done.set()

    def _handle(self, req, client_address):
        if isinstance(req, RunCommandRequest):
            with self._lock:
                if self._thread is not None:          # a retried request, or a second command: one command per service
                    return network.AckResponse()
                env = self._add_envs(dict(os.environ), self._command_env)
                env = self._add_envs(env, req.env or {})
                self._capture['stdout'] = _Capture() if req.capture_stdout else None
                self._capture['stderr'] = _Capture() if req.capture_stderr else None
                self._thread = threading.Thread(target=self._run_command, name='hvd-task-command', daemon=True,
                                                args=(req.command, env, req.capture_stdout, req.capture_stderr,
                                                      req.prefix_output_with_timestamp))
                self._thread.start()
                self._command_started.set()
            return network.AckResponse()
        if isinstance(req, StreamCommandOutputRequest):
            self.wait_for_command_start(timeout=60)
            capture = self._capture[req.stream]
            if capture is None:
                raise CommandOutputNotCaptured('the command was run without capturing its %s' % req.stream)
            data, nxt, finished = capture.read_from(req.offset)
            return StreamCommandOutputResponse(data, nxt, finished)
        if isinstance(req, AbortCommandRequest):
            self._abort.set()
            return network.AckResponse()
        if isinstance(req, CommandExitCodeRequest):
            return CommandExitCodeResponse(self._command_done.is_set(), self._exit_code)
        if isinstance(req, WaitForCommandExitCodeRequest):
            self._command_done.wait(max(0.0, min(float(req.delay), 30.0)))
            return WaitForCommandExitCodeResponse(self._exit_code if self._command_done.is_set() else None)
        if isinstance(req, NotifyInitialRegistrationCompleteRequest):
            self._initial_registration_complete.set()
            return network.AckResponse()
        if isinstance(req, RegisterCodeResultRequest):
            self._fn_result = req.result
            return network.AckResponse()
        return super(BasicTaskService, self)._handle(req, client_address)

    # -- what the process that hosts the service asks ------------------------------------------------------------------------------
    def fn_result(self):
        return self._fn_result

    def wait_for_initial_registration(self, timeout):
        """`timeout`: a runner.common.util.timeout.Timeout (raises its own exception) or seconds."""
        while not self._initial_registration_complete.wait(0.1):
            if hasattr(timeout, 'check_time_out_for'):
                timeout.check_time_out_for('tasks to start')
            elif timeout is not None:
                timeout -= 0.1
                if timeout <= 0:
                    raise TimeoutError('initial registration did not complete in time')

    def wait_for_command_start(self, timeout=None):
        if not self._command_started.wait(timeout):
            raise TimeoutError('no command was started within %s s' % timeout)

    def check_for_command_start(self, seconds):
        return self._command_started.wait(seconds)

    def wait_for_command_termination(self):
        self._command_done.wait()

    def command_exit_code(self):
        return self._exit_code


class BasicTaskClient(network.BasicClient):
    def __init__(self, service_name, task_addresses, key, verbose=0, match_intf=False, attempts=3):
        super(BasicTaskClient, self).__init__(service_name, task_addresses, key, verbose, match_intf=match_intf, attempts=attempts)

    def run_command(self, command, env, capture_stdout=False, capture_stderr=False, prefix_output_with_timestamp=False):
        self._send(RunCommandRequest(command, env, capture_stdout, capture_stderr, prefix_output_with_timestamp))

    def stream_command_output(self, stdout=None, stderr=None):
        """Copies the command's captured output into the given file-like objects while it runs; returns the threads (join them
        to wait for end of output)."""
        def pump(request_cls, sink):
            offset = 0
            while True:
                resp = self._send(request_cls(offset))
                if resp.data:
                    sink.write(resp.data.decode('utf-8', 'replace') if not isinstance(sink, Pipe) and hasattr(sink, 'encoding') else resp.data)
                    if hasattr(sink, 'flush'):
                        sink.flush()
                offset = resp.next_offset
                if resp.finished:
                    return
        threads = []
        for cls, sink in ((StreamCommandStdOutRequest, stdout), (StreamCommandStdErrRequest, stderr)):
            if sink is not None:
                t = threading.Thread(target=pump, args=(cls, sink), daemon=True)
                t.start()
                threads.append(t)
        return threads

    def abort_command(self):
        self._send(AbortCommandRequest())

    def notify_initial_registration_complete(self):
        self._send(NotifyInitialRegistrationCompleteRequest())

    def command_terminated(self):
        return self._send(CommandExitCodeRequest()).terminated

    def command_result(self):
        """(terminated, exit_code)"""
        resp = self._send(CommandExitCodeRequest())
        return resp.terminated, resp.exit_code

    def register_code_result(self, result):
        self._send(RegisterCodeResultRequest(result))

    def wait_for_command_termination(self, delay=1.0):
        self.wait_for_command_exit_code(delay)

    def wait_for_command_exit_code(self, delay=1.0):
        """Blocks until the command has ended; `delay` is how long one request may be held by the service."""
        while True:
            code = self._send(WaitForCommandExitCodeRequest(delay)).exit_code
            if code is not None:
                return code
            time.sleep(0)
