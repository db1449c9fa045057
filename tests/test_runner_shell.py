"""safe_shell_exec / prefix_connection / thread helpers — the case families of the reference's test/single/test_run.py
(test_prefix_connection_*, test_safe_shell_exec_*, test_in_thread_args, test_on_event)."""
import io
import os
import signal
import subprocess
import sys
import threading
import time

import pytest

from conftest import REPO
from horovod_b200.runner.common.util import safe_shell_exec
from horovod_b200.runner.util.threads import in_thread, on_event


def _pc(data, prefix='stdout', index=0, ts=False, chunks=None):
    r, w = os.pipe()
    dst = io.StringIO()
    t = threading.Thread(target=safe_shell_exec.prefix_connection, args=(os.fdopen(r, 'rb', 0), dst, prefix, index, ts))
    t.start()
    for c in (chunks or [data]):
        os.write(w, c)
        time.sleep(0.02)
    os.close(w)
    t.join(5)
    return dst.getvalue()


def test_prefix_connection_basic_and_without_trailing_newline():
    assert _pc(b'first line\nsecond line\n') == '[0]<stdout>:first line\n[0]<stdout>:second line\n'
    assert _pc(b'first line\nlast') == '[0]<stdout>:first line\n[0]<stdout>:last\n'
    assert _pc(b'') == ''


def test_prefix_connection_without_index_or_prefix():
    assert _pc(b'a\nb\n', index=None) == '<stdout>:a\n<stdout>:b\n'
    assert _pc(b'a\nb', prefix=None, index=None) == 'a\nb'


def test_prefix_connection_unicode_split_across_reads():
    text = 'héllo wörld ✓ 漢字\n'
    raw = text.encode('utf-8')
    cut = raw.index('✓'.encode('utf-8')) + 1  # in the middle of a 3-byte character
    assert _pc(raw, chunks=[raw[:cut], raw[cut:]]) == '[0]<stdout>:' + text


def test_prefix_connection_carriage_returns_keep_their_tag():
    assert _pc(b'10%\r50%\r100%\ndone\n') == '[0]<stdout>:10%\r[0]<stdout>:50%\r[0]<stdout>:100%\n[0]<stdout>:done\n'
    assert _pc(b'10%\r50%\n', prefix=None, index=None) == '10%\r50%\n'


def test_prefix_connection_with_timestamp():
    out = _pc(b'x\n', ts=True)
    assert out.endswith('[0]<stdout>:x\n') and len(out) > len('[0]<stdout>:x\n') + 10
    assert time.strftime('%Y') in out


def test_prefix_connection_streams():
    r, w = os.pipe()

    class Dst:
        def __init__(self):
            self.parts, self.ev = [], threading.Event()

        def write(self, s):
            self.parts.append(s)
            self.ev.set()

        def flush(self):
            pass
    dst = Dst()
    t = threading.Thread(target=safe_shell_exec.prefix_connection, args=(os.fdopen(r, 'rb', 0), dst, 'stdout', 1, False))
    t.start()
    os.write(w, b'early\n')
    assert dst.ev.wait(2), 'output must be forwarded before the writer closes the pipe'
    os.close(w)
    t.join(5)
    assert ''.join(dst.parts) == '[1]<stdout>:early\n'


def test_execute_captures_streams_and_exit_code():
    out, err = io.StringIO(), io.StringIO()
    rc = safe_shell_exec.execute('echo to-out; echo to-err 1>&2; printf last-no-eol; exit 7', stdout=out, stderr=err, index=3)
    assert rc == 7
    assert out.getvalue() == '[3]<stdout>:to-out\n[3]<stdout>:last-no-eol\n' and err.getvalue() == '[3]<stderr>:to-err\n'
    out = io.StringIO()
    assert safe_shell_exec.execute('echo $HVD_T_VAR', env=dict(os.environ, HVD_T_VAR='xyz'), stdout=out, stderr=io.StringIO()) == 0
    assert out.getvalue() == 'xyz\n'


def test_execute_interrupts_on_event_and_kills_grandchildren(tmp_path):
    pidfile = tmp_path / 'pids'
    ev = threading.Event()
    cmd = f'(sleep 300 & echo $! >> {pidfile}; wait) & echo $! >> {pidfile}; sleep 300'
    res = {}
    t = threading.Thread(target=lambda: res.setdefault('rc', safe_shell_exec.execute(cmd, events=[ev], stdout=io.StringIO(), stderr=io.StringIO())))
    t0 = time.time()
    t.start()
    for _ in range(100):
        if pidfile.exists() and len(pidfile.read_text().split()) >= 2:
            break
        time.sleep(0.05)
    ev.set()
    t.join(20)
    assert not t.is_alive() and res['rc'] != 0 and time.time() - t0 < 15
    time.sleep(0.3)
    for pid in map(int, pidfile.read_text().split()):
        with pytest.raises(ProcessLookupError):
            os.kill(pid, 0)


def test_execute_dies_with_the_launcher(tmp_path):
    """kill -9 of the launcher process must not leave the workers behind (the reference's middleman test)."""
    pidfile = tmp_path / 'pid'
    code = ('import sys; sys.path.insert(0, %r)\n'
            'from horovod_b200.runner.common.util import safe_shell_exec\n'
            'safe_shell_exec.execute("echo $$ > %s; sleep 300")\n' % (REPO, pidfile))
    launcher = subprocess.Popen([sys.executable, '-c', code])
    for _ in range(200):
        if pidfile.exists() and pidfile.read_text().strip():
            break
        time.sleep(0.05)
    worker = int(pidfile.read_text())
    os.kill(worker, 0)
    launcher.send_signal(signal.SIGKILL)
    launcher.wait()
    for _ in range(200):
        try:
            os.kill(worker, 0)
        except ProcessLookupError:
            break
        time.sleep(0.05)
    else:
        os.kill(worker, signal.SIGKILL)
        raise AssertionError('worker survived the death of its launcher')


def test_in_thread_and_on_event():
    got = []
    t = in_thread(lambda a, b: got.append(a + b), args=(1, 2))
    t.join(2)
    assert got == [3] and t.daemon
    with pytest.raises(ValueError):
        in_thread(lambda: None, args=5)
    ev, stop = threading.Event(), threading.Event()
    fired = []
    th = on_event(ev, lambda x: fired.append(x), args=('go',), stop=stop, check_stop_interval_s=0.05)
    time.sleep(0.1)
    assert fired == []
    ev.set()
    th.join(2)
    assert fired == ['go']
    ev2 = threading.Event()
    th2 = on_event(ev2, lambda: fired.append('never'), stop=stop, check_stop_interval_s=0.05)
    stop.set()
    th2.join(2)
    assert not th2.is_alive() and 'never' not in fired
