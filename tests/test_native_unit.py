"""C++ unit tests of the native runtime (csrc/common/selftest.cc): wire codec, fusion planner, cross-rank validation,
response cache LRU, CPU collectives, CPU Adasum VHDD, Bayesian optimiser, autotuner state machine and 4 complete
Engine instances negotiating over the in-process loopback transport."""
import ctypes


def test_native_selftest():
    from horovod_b200.common.basics import load_library
    lib = load_library()
    lib.hvd_selftest.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(1 << 16)
    failures = lib.hvd_selftest(4, buf, 1 << 16)
    assert failures == 0, buf.value.decode()


def test_native_selftest_with_log_depth_bit_reduction():
    """The same binary in a fresh process with HVD_BITS_TREE_MIN_RANKS=2 (read once per process): every bit-vector reduction of
    the self-test — TestBitsAmong over 2..11 ranks, and the negotiation of the 4 loopback engines — takes the recursive-doubling
    path instead of the star."""
    import os
    import subprocess
    import sys
    code = ("import ctypes\n"
            "from horovod_b200.common.basics import load_library\n"
            "lib = load_library()\n"
            "lib.hvd_selftest.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]\n"
            "buf = ctypes.create_string_buffer(1 << 16)\n"
            "n = lib.hvd_selftest(4, buf, 1 << 16)\n"
            "print(buf.value.decode())\n"
            "raise SystemExit(1 if n else 0)\n")
    env = dict(os.environ, HVD_BITS_TREE_MIN_RANKS='2', PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
