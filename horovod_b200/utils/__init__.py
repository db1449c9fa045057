"""Measurement and housekeeping helpers shared by bench.py, the benches under bench/ and user scripts."""
from horovod_b200.utils.timing import ClockSampler, device_timer, max_over_ranks, measured_peaks  # noqa: F401
