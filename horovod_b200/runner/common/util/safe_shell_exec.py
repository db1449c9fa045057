"""Run a shell command in its own process group, forward its output with optional per-rank prefixes, and make
sure the whole process tree dies when asked to (event) or when the launcher itself dies.

Role parity: horovod/runner/common/util/safe_shell_exec.py.  Implementation: subprocess with start_new_session
instead of a fork middleman; termination = SIGTERM to the group, SIGKILL after a grace period.
"""
import os
import signal
import subprocess
import sys
import threading
import time

GRACEFUL_TERMINATION_TIME_S = 5


def terminate_executor_shell_and_children(pid):
    """SIGTERM the process group, then SIGKILL whatever is left after the grace period."""
    try:
        pgid = os.getpgid(pid)
    except ProcessLookupError:
        return
    try:
        os.killpg(pgid, signal.SIGTERM)
    except ProcessLookupError:
        return
    deadline = time.time() + GRACEFUL_TERMINATION_TIME_S
    while time.time() < deadline:
        try:
            os.killpg(pgid, 0)
        except ProcessLookupError:
            return
        time.sleep(0.05)
    try:
        os.killpg(pgid, signal.SIGKILL)
    except ProcessLookupError:
        pass


def prefix_connection(src, dst, prefix, index, prefix_output_with_timestamp):
    """Copies lines from src to dst, optionally prefixed with `[index]<prefix>:` and a timestamp."""
    def fmt(line):
        if prefix is None:
            return line
        ts = time.strftime('%a %b %d %H:%M:%S %Y') if prefix_output_with_timestamp else ''
        tag = f'[{index}]<{prefix}>' if index is not None else f'<{prefix}>'
        return f'{ts}{tag}:{line}'
    for raw in iter(src.readline, b''):
        text = raw.decode('utf-8', errors='replace')
        try:
            dst.write(fmt(text))
            dst.flush()
        except ValueError:
            break
    src.close()


def execute(command, env=None, stdout=None, stderr=None, index=None, events=None, prefix_output_with_timestamp=False):
    """Runs `command` through the shell; returns its exit code. Any event in `events` being set kills the tree."""
    stdout = stdout if stdout is not None else sys.stdout
    stderr = stderr if stderr is not None else sys.stderr
    proc = subprocess.Popen(command, shell=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            start_new_session=True, executable='/bin/bash')
    threads = [
        threading.Thread(target=prefix_connection, args=(proc.stdout, stdout, 'stdout' if index is not None else None,
                                                         index, prefix_output_with_timestamp), daemon=True),
        threading.Thread(target=prefix_connection, args=(proc.stderr, stderr, 'stderr' if index is not None else None,
                                                         index, prefix_output_with_timestamp), daemon=True),
    ]
    for t in threads:
        t.start()
    stop = threading.Event()

    def watch():
        while not stop.is_set():
            if any(e.is_set() for e in (events or [])):
                terminate_executor_shell_and_children(proc.pid)
                return
            time.sleep(0.1)

    watcher = threading.Thread(target=watch, daemon=True)
    watcher.start()
    try:
        rc = proc.wait()
    except KeyboardInterrupt:
        terminate_executor_shell_and_children(proc.pid)
        rc = proc.wait()
    stop.set()
    for t in threads:
        t.join(timeout=2)
    return rc
