"""Hosts, slots and the rank layout of a job.

Role parity: horovod/runner/common/util/hosts.py (`HostInfo`, `SlotInfo`, `INVALID_SLOT_INFO`, `parse_hosts`,
`parse_hosts_and_slots`, `parse_host_files`, `get_host_assignments`).  Layout rule: ranks fill host after host in the
order given; `local_rank` counts inside a host; `cross_rank` of a process is its host's position among the hosts that
have a process with the same `local_rank` (hosts may have different slot counts).
"""
import re

_HOST_SLOTS = re.compile(r'^(?P<host>[\w.\-\[\]:]+):(?P<slots>\d+)$')


class HostInfo:
    __slots__ = ('hostname', 'slots')

    def __init__(self, hostname, slots):
        self.hostname, self.slots = hostname, slots

    @staticmethod
    def from_string(host_string):
        host, slots = _split_host_slots(host_string)
        return HostInfo(host, slots)

    def __repr__(self):
        return 'HostInfo(%s:%d)' % (self.hostname, self.slots)


class SlotInfo:
    """One process of the job: where it runs and all six rank / size numbers."""
    _FIELDS = ('hostname', 'rank', 'local_rank', 'cross_rank', 'size', 'local_size', 'cross_size')

    def __init__(self, hostname, rank, local_rank, cross_rank, size=None, local_size=None, cross_size=None):
        self.hostname, self.rank, self.local_rank, self.cross_rank = hostname, rank, local_rank, cross_rank
        self.size, self.local_size, self.cross_size = size, local_size, cross_size

    def to_response_string(self):
        """`rank,size,local_rank,local_size,cross_rank,cross_size` — the elastic rendezvous reply."""
        return '%s,%s,%s,%s,%s,%s' % (self.rank, self.size, self.local_rank, self.local_size, self.cross_rank, self.cross_size)

    def _key(self):
        return tuple(getattr(self, f) for f in self._FIELDS)

    def __eq__(self, other):
        return isinstance(other, SlotInfo) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return 'SlotInfo(%s)' % ', '.join('%s=%s' % (f, getattr(self, f)) for f in self._FIELDS)


INVALID_SLOT_INFO = SlotInfo(hostname='', rank=-1, local_rank=-1, cross_rank=-1, size=-1, local_size=-1, cross_size=-1)


def _split_host_slots(text):
    m = _HOST_SLOTS.match(text.strip())
    if not m:
        raise ValueError('Invalid host input, please make sure it has format as : host1:2,host2:4,host3:1.')
    return m.group('host'), int(m.group('slots'))


def parse_hosts_and_slots(hosts):
    """'h1:2,h2:4' -> (['h1', 'h2'], {'h1': 2, 'h2': 4})"""
    pairs = [_split_host_slots(item) for item in hosts.split(',')]
    return [h for h, _ in pairs], dict(pairs)


def parse_hosts(hosts_string):
    """'h1:2,h2:4' -> [HostInfo, HostInfo]"""
    return [HostInfo.from_string(item) for item in hosts_string.split(',')]


def parse_host_files(filename):
    """A hostfile has one host per line: `name slots=N` (mpirun style), `name:N`, or a bare name (one slot); `#` starts a
    comment.  Returns the equivalent `-H` string."""
    entries = []
    with open(filename) as f:
        for raw in f:
            line = raw.split('#', 1)[0].strip()
            if not line:
                continue
            name = line.split()[0]
            m = re.search(r'\bslots\s*=\s*(\d+)', line)
            if m:
                entries.append('%s:%s' % (name, m.group(1)))
            elif ':' in name:
                entries.append(name)
            else:
                entries.append(name + ':1')
    return ','.join(entries)


def get_host_assignments(hosts, min_num_proc, max_num_proc=None):
    """[HostInfo] -> rank-ordered [SlotInfo] using at most `max_num_proc` slots; ValueError when fewer than
    `min_num_proc` slots are available."""
    budget = max_num_proc if max_num_proc is not None else sum(h.slots for h in hosts)
    used = []                                   # (hostname, processes placed on it)
    for h in hosts:
        take = max(0, min(h.slots, budget))
        budget -= take
        used.append((h.hostname, take))
    world = sum(n for _, n in used)
    if world < min_num_proc:
        raise ValueError('Requested more processes ({}) than there are available slots ({})'.format(min_num_proc, world))
    layout, rank = [], 0
    for host_pos, (name, count) in enumerate(used):
        for local_rank in range(count):
            peers = [i for i, (_, n) in enumerate(used) if n > local_rank]   # hosts that have this local rank
            layout.append(SlotInfo(hostname=name, rank=rank, local_rank=local_rank, cross_rank=peers.index(host_pos),
                                   size=world, local_size=count, cross_size=len(peers)))
            rank += 1
    return layout
