"""Timing done the way the profiling recipe asks for: CUDA events on the launching stream with a synchronize on both
sides, the maximum over ranks, and the SM clock / throttle reasons sampled while the timed region runs."""
import json
import os
import subprocess
import threading
import time
from contextlib import contextmanager


class _Elapsed:
    ms = None


@contextmanager
def device_timer(sync_ranks=None):
    """`with device_timer(hvd.barrier) as t: ...; t.ms` — device milliseconds of the block on the current stream."""
    import torch
    t = _Elapsed()
    if sync_ranks:
        sync_ranks()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield t
        e1.record()
        torch.cuda.synchronize()
        t.ms = e0.elapsed_time(e1)
    else:
        t0 = time.perf_counter()
        yield t
        t.ms = (time.perf_counter() - t0) * 1e3
    if sync_ranks:
        sync_ranks()


def max_over_ranks(value, name='utils.max'):
    """The slowest rank decides (a multi-GPU number is never the rank-0 number)."""
    import torch
    import horovod_b200.torch as hvd
    return hvd.allreduce(torch.tensor([float(value)], dtype=torch.float64), op=hvd.Max, name=name).item()


def measured_peaks(root=None):
    """MEASURED_PEAKS.json of the repo (driver-written): measured copy bandwidth / cuBLAS bf16 rate used as roofline
    denominators; {} when absent."""
    root = root or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        with open(os.path.join(root, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi every 200 ms between start() and stop()."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    REASONS = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index=0):
        self.index, self.proc, self.lines, self.thread = index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self
        self.thread = threading.Thread(target=lambda: self.lines.extend(l.strip() for l in self.proc.stdout), daemon=True)
        self.thread.start()
        return self

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        return self.summarise(self.lines)

    @classmethod
    def summarise(cls, lines):
        sm, mx, reasons = [], None, set()
        for line in lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            reasons.update(n for n, v in zip(cls.REASONS, f[3:7]) if v.lower().startswith('active'))
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'samples': len(sm), 'reasons': sorted(reasons)}
