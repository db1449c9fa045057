"""Which hosts can the elastic job use right now?

`HostDiscoveryScript` runs the user's executable (one `host[:slots]` per line), `FixedHosts` serves a static (test-
settable) table.  `HostManager` polls one of them, classifies the difference to the previous poll (hosts/slots added,
removed, both), keeps a cooling-down blacklist of hosts whose workers failed, and keeps hosts in *seniority order* so
that ranks are handed out to the longest-serving hosts first (rank 0 then sits on a host that holds committed state).

Capability parity: horovod/runner/elastic/discovery.py (HostDiscoveryScript, FixedHosts, HostManager + blacklist
cooldown with exponential back-off, host ordering).
"""
import io
import logging
import random
import threading
import time

from horovod_b200.runner.common.util import safe_shell_exec
from horovod_b200.runner.elastic.worker import HostUpdateResult

COOLDOWN_FLOOR_S = 1
COOLDOWN_CEIL_S = 3600


class HostDiscovery(object):
    def find_available_hosts_and_slots(self):
        """Returns {hostname: slots}."""
        raise NotImplementedError()


class HostDiscoveryScript(HostDiscovery):
    def __init__(self, discovery_script, slots):
        self._script = discovery_script
        self._default_slots = slots

    def find_available_hosts_and_slots(self):
        sink = io.StringIO()
        rc = safe_shell_exec.execute(self._script, stdout=sink)
        if rc != 0:
            raise RuntimeError('Failed to execute discovery script: {}. Exit code: {}'.format(self._script, rc))
        table = {}
        for raw in sink.getvalue().splitlines():
            entry = raw.strip()
            if not entry:
                continue
            name, sep, count = entry.rpartition(':')
            if sep and count.isdigit():
                table[name] = int(count)
            elif self._default_slots is not None:
                table[entry] = self._default_slots
            else:
                raise ValueError('host discovery script printed "%s" without `:slots` and --slots-per-host was not given' % entry)
        return table


class FixedHosts(HostDiscovery):
    def __init__(self, host_slots):
        self._table = dict(host_slots)

    def find_available_hosts_and_slots(self):
        return dict(self._table)

    def set(self, host_slots):
        self._table = dict(host_slots)


class _Penalty(object):
    """Blacklist record of one host: strikes grow the cool-down exponentially inside [floor, ceil]."""

    def __init__(self, cooldown_range):
        self.range = cooldown_range
        self.strikes = 0
        self.banned = False
        self.until = 0.0
        self.signal = threading.Event()   # set when the host gets banned: its workers are torn down

    def ban(self):
        self.banned = True
        self.signal.set()
        if self.range is None:
            return                       # no cool-down configured: banned for good
        if self.until > time.time():
            return                       # already cooling down
        self.strikes += 1
        lo, hi = self.range
        delay = lo * (1 << self.strikes) + random.random() * lo
        self.until = time.time() + max(lo, min(hi, delay))

    def cooled_down(self):
        return self.banned and self.range is not None and 0 < self.until <= time.time()

    def pardon(self):
        self.banned = False
        self.until = 0.0
        self.signal = threading.Event()


class HostSnapshot(object):
    """Usable hosts at one point in time."""

    def __init__(self, slots, order):
        self.host_slots = slots
        self.host_assignment_order = order

    @property
    def available_hosts(self):
        return set(self.host_assignment_order)

    def get_slots(self, host):
        return self.host_slots.get(host, 0)

    def count_available_slots(self):
        return sum(self.host_slots.get(h, 0) for h in self.host_assignment_order)


class HostManager(object):
    def __init__(self, discovery, cooldown_range=None):
        if cooldown_range is not None:
            lo, hi = cooldown_range
            if lo < COOLDOWN_FLOOR_S or hi > COOLDOWN_CEIL_S or lo > hi:
                raise ValueError(f'blacklist cooldown range must lie within [{COOLDOWN_FLOOR_S}, {COOLDOWN_CEIL_S}] seconds')
        self._discovery = discovery
        self._cooldown_range = tuple(cooldown_range) if cooldown_range is not None else None
        self._lock = threading.Lock()
        self._penalties = {}
        self._slots = {}
        self._order = []

    def _penalty(self, host):
        p = self._penalties.get(host)
        if p is None:
            p = self._penalties[host] = _Penalty(self._cooldown_range)
        return p

    def update_available_hosts(self):
        """Polls discovery; returns a HostUpdateResult bit mask describing what changed."""
        found = self._discovery.find_available_hosts_and_slots()
        with self._lock:
            change = HostUpdateResult.no_update
            back = [h for h in found if h in self._penalties and self._penalties[h].cooled_down()]
            if found == self._slots and not back:
                return change
            for h in self._slots:
                if h not in found or found[h] < self._slots[h]:
                    change |= HostUpdateResult.removed
            for h in found:
                if h not in self._slots or found[h] > self._slots[h] or h in back:
                    change |= HostUpdateResult.added
            for h in back:
                self._penalties[h].pardon()
            usable = {h for h in found if not self._penalty(h).banned}
            # seniority: keep the relative order of hosts we already use, append newcomers alphabetically
            self._order = [h for h in self._order if h in usable] + sorted(h for h in usable if h not in self._order)
            self._slots = dict(found)
            return change

    @property
    def current_hosts(self):
        with self._lock:
            return HostSnapshot(dict(self._slots), [h for h in self._order if not self._penalty(h).banned])

    def blacklist(self, host):
        with self._lock:
            p = self._penalty(host)
            if not p.banned:
                logging.info('blacklist failing host: %s', host)
            p.ban()

    def is_blacklisted(self, host):
        with self._lock:
            return self._penalty(host).banned

    def get_host_event(self, host):
        with self._lock:
            return self._penalty(host).signal
