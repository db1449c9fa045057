"""Run a shell command in its own process group, forward its output with optional per-rank prefixes, and make
sure the whole process tree dies when asked to (event) or when the launcher itself dies.

Role parity: horovod/runner/common/util/safe_shell_exec.py.  Implementation: the command runs under a tiny
supervisor process (an inline `python -c` script, no package import) that is the leader of a new session, watches its
parent and takes the whole process group down when the launcher disappears (even by SIGKILL) — the job the reference
gives to a forked "middleman"; no fork() of the multi-threaded launcher is needed.  Termination on request = SIGTERM to
the group, SIGKILL after a grace period.
"""
import os
import signal
import subprocess
import sys
import threading
import time

GRACEFUL_TERMINATION_TIME_S = 5

# argv: <launcher pid> <shell command>.  Exit status = the command's (128 + signal when it was killed).
_SUPERVISOR = r"""
import os, signal, subprocess, sys, time
ppid, cmd = int(sys.argv[1]), sys.argv[2]
os.setsid()
p = subprocess.Popen(['/bin/bash', '-c', cmd])
def down(signum=None, frame=None):
    signal.signal(signal.SIGTERM, signal.SIG_IGN)
    pg = os.getpgid(0)
    try:
        os.killpg(pg, signal.SIGTERM)
    except ProcessLookupError:
        pass
    t0 = time.time()
    while time.time() - t0 < %d:
        if p.poll() is not None and not [x for x in os.listdir('/proc') if x.isdigit() and x != str(os.getpid()) and _pgid(x) == pg]:
            os._exit(143)
        time.sleep(0.05)
    os.killpg(pg, signal.SIGKILL)
def _pgid(pid):
    try:
        return os.getpgid(int(pid))
    except (ProcessLookupError, PermissionError):
        return -1
signal.signal(signal.SIGTERM, down)
while True:
    rc = p.poll()
    if rc is not None:
        os._exit(rc if rc >= 0 else 128 - rc)
    if os.getppid() != ppid:
        down()
    time.sleep(0.1)
""" % GRACEFUL_TERMINATION_TIME_S


def terminate_executor_shell_and_children(pid):
    """Asks the supervisor `pid` to take its process group down (SIGTERM; it escalates to SIGKILL after the grace period);
    if the supervisor itself does not go away, its group is killed from here.  Never signals a group that `pid` does not
    lead (right after spawn the supervisor is still in the launcher's group)."""
    try:
        os.kill(pid, signal.SIGTERM)
    except ProcessLookupError:
        return
    def gone():
        try:
            os.kill(pid, 0)
        except ProcessLookupError:
            return True
        try:  # exited but not reaped yet (the thread that owns the Popen object reaps it)
            with open('/proc/%d/stat' % pid) as f:
                return f.read().rsplit(')', 1)[1].split()[0] == 'Z'
        except (OSError, IndexError):
            return True

    deadline = time.time() + GRACEFUL_TERMINATION_TIME_S + 2
    while time.time() < deadline:
        if gone():
            return
        time.sleep(0.05)
    try:
        if os.getpgid(pid) == pid:
            os.killpg(pid, signal.SIGKILL)
        else:
            os.kill(pid, signal.SIGKILL)
    except ProcessLookupError:
        pass


def prefix_connection(src, dst, prefix, index, prefix_output_with_timestamp):
    """Copies src to dst as it arrives.  With a prefix every line — and every carriage-return-separated segment of a
    line, so that progress bars keep their tag — starts with `[index]<prefix>:` (optionally after a timestamp); a last
    line without a newline gets one, so outputs of different ranks never run into each other.  Bytes are decoded
    incrementally: a multi-byte UTF-8 character split across two reads is never mangled."""
    import codecs
    decoder = codecs.getincrementaldecoder('utf-8')(errors='replace')

    def tag():
        if prefix is None:
            return ''
        ts = time.strftime('%a %b %d %H:%M:%S %Y') if prefix_output_with_timestamp else ''
        return f'{ts}[{index}]<{prefix}>:' if index is not None else f'{ts}<{prefix}>:'

    at_line_start = True
    last = ''
    fd = src.fileno()
    try:
        while True:
            raw = os.read(fd, 65536)
            text = decoder.decode(raw, final=not raw)
            out = []
            for ch in text:
                if at_line_start and prefix is not None:
                    out.append(tag())
                out.append(ch)
                at_line_start = ch in ('\n', '\r')
            if text:
                last = text[-1]
                dst.write(''.join(out))
                dst.flush()
            if not raw:
                break
        if prefix is not None and last not in ('', '\n'):
            dst.write('\n')
            dst.flush()
    except ValueError:  # dst closed
        pass
    finally:
        src.close()


def execute(command, env=None, stdout=None, stderr=None, index=None, events=None, prefix_output_with_timestamp=False):
    """Runs `command` through the shell; returns its exit code. Any event in `events` being set kills the tree."""
    stdout = stdout if stdout is not None else sys.stdout
    stderr = stderr if stderr is not None else sys.stderr
    proc = subprocess.Popen([sys.executable, '-S', '-c', _SUPERVISOR, str(os.getpid()), command], env=env, stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE)
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
