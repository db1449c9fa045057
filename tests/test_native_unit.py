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
