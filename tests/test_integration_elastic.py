"""Fault-injection integration tests: a real `hvdrun` elastic job on this machine, with `localhost` and `127.0.0.1`
treated as two distinct hosts (the trick of the reference's test/integration/elastic_common.py:128-153), workers that
kill themselves or raise on a schedule, and a discovery script whose output changes with the training epoch."""
import json
import os
import stat
import subprocess
import sys

import pytest

from conftest import REPO

MAIN = os.path.join(REPO, 'tests', 'integration', 'elastic_torch_main.py')


def _run_elastic(tmp_path, discovery_lines_by_epoch, np_, min_np, max_np, exit_schedule=None, exit_mode='exception', extra=(),
                 timeout=300, expect_fail=False, batch_sleep=0.0, main_args=(), env_extra=None):
    logfile = str(tmp_path / 'log.jsonl')
    epoch_file = str(tmp_path / 'epoch')
    with open(epoch_file, 'w') as f:
        f.write('0')
    script = tmp_path / 'discover.sh'
    # the discovery script reads the current epoch and prints the host set scheduled for it
    body = ['#!/bin/bash', f'epoch=$(cat {epoch_file} 2>/dev/null || echo 0)']
    for i, (upto, lines) in enumerate(discovery_lines_by_epoch):
        cond = 'if' if i == 0 else 'elif'
        test = f'[ "$epoch" -le {upto} ]' if upto is not None else 'true'
        body.append(f'{cond} {test}; then')
        body += [f'  echo "{l}"' for l in lines]
    body.append('fi')
    script.write_text('\n'.join(body) + '\n')
    script.chmod(script.stat().st_mode | stat.S_IEXEC)
    env = dict(os.environ, PYTHONPATH=REPO, OMP_NUM_THREADS='1', HOROVOD_LOG_LEVEL='warning')
    env.update(env_extra or {})
    cmd = [sys.executable, '-m', 'horovod_b200.runner.launch', '-np', str(np_), '--min-np', str(min_np), '--max-np', str(max_np),
           '--host-discovery-script', str(script), *extra, sys.executable, MAIN, '--logfile', logfile,
           '--discovery-schedule-epoch-file', epoch_file, '--exit-mode', exit_mode,
           '--exit-schedule', json.dumps(exit_schedule or {}), '--batch-sleep', str(batch_sleep), *main_args]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=timeout, cwd=REPO)
    out = p.stdout.decode(errors='replace')
    if not expect_fail:
        assert p.returncode == 0, out[-5000:]
    recs = [json.loads(l) for l in open(logfile)] if os.path.exists(logfile) else []
    return p.returncode, out, recs


def test_elastic_static_hosts_no_failure(native_built, tmp_path):
    rc, out, recs = _run_elastic(tmp_path, [(None, ['localhost:2'])], 2, 2, 2)
    assert recs[-1].get('done') and recs[-1]['size'] == 2
    assert sorted({(r['epoch'], r['rank']) for r in recs if 'epoch' in r}) == [(e, r) for e in range(3) for r in range(2)]


def test_elastic_single_rank_failure_recovers(native_built, tmp_path):
    """Rank 1 (on the second 'host') dies in epoch 1: its host is blacklisted, the survivor restores the last commit and
    finishes alone (min_np = 1)."""
    rc, out, recs = _run_elastic(tmp_path, [(None, ['localhost:1', '127.0.0.1:1'])], 2, 1, 2,
                                 exit_schedule={'1,2': [1]}, exit_mode='kill')
    done = recs[-1]
    assert done.get('done') and done['size'] == 1, (done, out[-3000:])
    sizes_by_epoch = {}
    for r in recs:
        if 'epoch' in r:
            sizes_by_epoch.setdefault(r['epoch'], set()).add(r['size'])
    assert sizes_by_epoch[0] == {2} and sizes_by_epoch[2] == {1}, sizes_by_epoch


def test_elastic_hosts_added(native_built, tmp_path):
    """Starts on one host with 1 slot; after epoch 0 a second host appears: the job grows to 2 ranks without restart."""
    rc, out, recs = _run_elastic(tmp_path, [(0, ['localhost:1']), (None, ['localhost:1', '127.0.0.1:1'])], 1, 1, 2, batch_sleep=0.4)
    done = recs[-1]
    assert done.get('done'), out[-3000:]
    assert done['size'] == 2, (done, out[-3000:])
    assert any(r.get('size') == 1 for r in recs) and any(r.get('size') == 2 for r in recs)


def test_elastic_all_ranks_fail(native_built, tmp_path):
    rc, out, recs = _run_elastic(tmp_path, [(None, ['localhost:2'])], 2, 2, 2, exit_schedule={'0,1': [0, 1]},
                                 exit_mode='exception', expect_fail=True)
    assert rc != 0


def test_elastic_reset_limit(native_built, tmp_path):
    rc, out, recs = _run_elastic(tmp_path, [(None, ['localhost:1', '127.0.0.1:1'])], 2, 1, 2, exit_schedule={'0,2': [1]},
                                 exit_mode='kill', extra=['--reset-limit', '0'], expect_fail=True)
    assert rc != 0 and 'reset' in out.lower(), out[-2000:]
