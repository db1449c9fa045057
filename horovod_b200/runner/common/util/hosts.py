"""Host/slot bookkeeping: parse `-H host:slots,...` / hostfiles and lay ranks out over hosts.

Role parity: horovod/runner/common/util/hosts.py (HostInfo, SlotInfo, parse_hosts, get_host_assignments).
"""
import collections
import re


class HostInfo:
    def __init__(self, hostname, slots):
        self.hostname = hostname
        self.slots = slots

    @staticmethod
    def from_string(host_string):
        hostname, slots = host_string.strip().split(':')
        return HostInfo(hostname, int(slots))

    def __repr__(self):
        return f'HostInfo({self.hostname}:{self.slots})'


class SlotInfo:
    def __init__(self, hostname, rank, local_rank, cross_rank, size=None, local_size=None, cross_size=None):
        self.hostname = hostname
        self.rank = rank
        self.size = size
        self.local_rank = local_rank
        self.local_size = local_size
        self.cross_rank = cross_rank
        self.cross_size = cross_size

    def to_response_string(self):
        return ','.join(str(v) for v in [self.rank, self.size, self.local_rank, self.local_size, self.cross_rank,
                                         self.cross_size])

    def __eq__(self, other):
        return isinstance(other, SlotInfo) and self.__dict__ == other.__dict__

    def __repr__(self):
        return 'SlotInfo(%s)' % ', '.join(f'{k}={v}' for k, v in self.__dict__.items())


INVALID_SLOT_INFO = SlotInfo(hostname='', rank=-1, local_rank=-1, cross_rank=-1, size=-1, local_size=-1, cross_size=-1)


def parse_host_files(filename):
    """Hostfile lines: `hostname slots=N` (mpirun style) or `hostname:N`."""
    hosts = []
    with open(filename, 'r') as f:
        for line in f.readlines():
            line = line.strip()
            if not line or line.startswith('#'):
                continue
            m = re.match(r'^(\S+)\s+slots\s*=\s*(\d+)', line)
            if m:
                hosts.append(f'{m.group(1)}:{m.group(2)}')
            elif ':' in line:
                hosts.append(line.split()[0])
            else:
                hosts.append(f'{line.split()[0]}:1')
    return ','.join(hosts)


def parse_hosts_and_slots(hosts):
    host_names = []
    host_to_slots = {}
    host_list = hosts.split(',')
    pattern = re.compile(r'^[\w.\-\[\]:]+:\d+$')
    for host in host_list:
        if not pattern.match(host.strip()):
            raise ValueError('Invalid host input, please make sure it has format as : host1:2,host2:4,host3:1.')
        hostname, slots = host.strip().rsplit(':', 1)
        host_names.append(hostname)
        host_to_slots[hostname] = int(slots)
    return host_names, host_to_slots


def parse_hosts(hosts_string):
    """'h1:2,h2:4' -> [HostInfo]"""
    return [HostInfo.from_string(s.strip().rsplit(':', 1)[0] + ':' + s.strip().rsplit(':', 1)[1]) for s in hosts_string.split(',')]


def get_host_assignments(hosts, min_num_proc, max_num_proc=None):
    """Assign ranks host by host (all slots of host 0 first). Returns a list of SlotInfo, rank-ordered.

    Raises ValueError when fewer than `min_num_proc` slots exist."""
    host_ranks = []
    cross_ranks = collections.defaultdict(dict)
    rank = 0
    for host_info in hosts:
        ranks = []
        for local_rank in range(host_info.slots):
            if rank == max_num_proc:
                break
            ranks.append(rank)
            rank += 1
            cross_ranks_at_local = cross_ranks[local_rank]
            cross_ranks_at_local[host_info.hostname] = len(cross_ranks_at_local)
        host_ranks.append((host_info, ranks))
    world_size = rank
    if world_size < min_num_proc:
        raise ValueError('Requested more processes ({}) than there are available slots ({})'.format(min_num_proc, world_size))
    alloc_list = []
    for host_info, ranks in host_ranks:
        local_size = len(ranks)
        for local_rank, rank in enumerate(ranks):
            cross_ranks_at_local = cross_ranks[local_rank]
            cross_rank = cross_ranks_at_local[host_info.hostname]
            cross_size = len(cross_ranks_at_local)
            alloc_list.append(SlotInfo(hostname=host_info.hostname, rank=rank, local_rank=local_rank, cross_rank=cross_rank,
                                       size=world_size, local_size=local_size, cross_size=cross_size))
    return alloc_list
