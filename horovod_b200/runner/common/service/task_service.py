"""Task-side RPC service: runs commands for the driver and probes which of its interfaces reach the next task
(reference runner/common/service/task_service.py)."""
import threading

from horovod_b200.runner.common.util import network, safe_shell_exec


class RunCommandRequest(object):
    def __init__(self, command, env):
        self.command = command
        self.env = env


class CommandExitCodeRequest(object):
    pass


class CommandExitCodeResponse(object):
    def __init__(self, terminated, exit_code):
        self.terminated = terminated
        self.exit_code = exit_code


class NotifyInitialRegistrationCompleteRequest(object):
    pass


class RegisterCodeResultRequest(object):
    def __init__(self, result):
        self.result = result


class BasicTaskService(network.BasicService):
    def __init__(self, name, index, key, nic, service_env_keys=None):
        super(BasicTaskService, self).__init__(name, key, nic)
        self._index = index
        self._initial_registration_complete = False
        self._wait_cond = threading.Condition()
        self._service_env_keys = service_env_keys or []
        self._command_thread = None
        self._command_exit_code = None
        self._command_abort = threading.Event()
        self._fn_result = None

    def _run_command(self, command, env):
        def run():
            self._command_exit_code = safe_shell_exec.execute(command, env=env, events=[self._command_abort])
        self._command_thread = threading.Thread(target=run, daemon=True)
        self._command_thread.start()

    def _handle(self, req, client_address):
        if isinstance(req, RunCommandRequest):
            with self._wait_cond:
                if self._command_thread is None:
                    self._run_command(req.command, req.env)
                self._wait_cond.notify_all()
            return network.AckResponse()
        if isinstance(req, NotifyInitialRegistrationCompleteRequest):
            with self._wait_cond:
                self._initial_registration_complete = True
                self._wait_cond.notify_all()
            return network.AckResponse()
        if isinstance(req, CommandExitCodeRequest):
            with self._wait_cond:
                terminated = self._command_thread is not None and not self._command_thread.is_alive()
                return CommandExitCodeResponse(terminated, self._command_exit_code if terminated else None)
        if isinstance(req, RegisterCodeResultRequest):
            self._fn_result = req.result
            return network.AckResponse()
        return super(BasicTaskService, self)._handle(req, client_address)

    def fn_result(self):
        return self._fn_result

    def wait_for_initial_registration(self, timeout):
        with self._wait_cond:
            while not self._initial_registration_complete:
                self._wait_cond.wait(timeout.remaining())
                timeout.check_time_out_for('tasks to start')

    def wait_for_command_start(self, timeout):
        with self._wait_cond:
            while self._command_thread is None:
                self._wait_cond.wait(timeout.remaining())
                timeout.check_time_out_for('command to run')

    def wait_for_command_termination(self):
        self._command_thread.join()

    def command_exit_code(self):
        return self._command_exit_code


class BasicTaskClient(network.BasicClient):
    def __init__(self, service_name, task_addresses, key, verbose, match_intf=False, attempts=3):
        super(BasicTaskClient, self).__init__(service_name, task_addresses, key, verbose, match_intf=match_intf, attempts=attempts)

    def run_command(self, command, env):
        self._send(RunCommandRequest(command, env))

    def notify_initial_registration_complete(self):
        self._send(NotifyInitialRegistrationCompleteRequest())

    def command_terminated(self):
        return self._send(CommandExitCodeRequest()).terminated

    def command_exit_code(self):
        return self._send(CommandExitCodeRequest()).exit_code

    def register_code_result(self, result):
        self._send(RegisterCodeResultRequest(result))

    def wait_for_command_termination(self, delay=1.0):
        import time
        try:
            while not self.command_terminated():
                time.sleep(delay)
        except Exception:
            pass
