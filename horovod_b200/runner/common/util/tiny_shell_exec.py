"""Run a short command and capture (output, exit_code) (reference runner/common/util/tiny_shell_exec.py)."""
import subprocess


def execute(command):
    try:
        r = subprocess.run(command, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
        return r.stdout.decode('utf-8', errors='replace'), r.returncode
    except Exception:
        return None
