"""Stable identifier of the physical host a process runs on (distinguishes containers sharing a hostname)."""
import hashlib
import os
import socket


def _namespaces():
    try:
        hash_ = ''
        for ns in sorted(os.listdir('/proc/self/ns')):
            hash_ += ns + '-' + os.readlink(os.path.join('/proc/self/ns', ns)) + ' '
        return hash_
    except OSError:
        return ''


def host_hash(salt=None):
    hostname = socket.gethostname()
    host_info = '{hostname}-{ns}'.format(hostname=hostname, ns=_namespaces())
    if salt:
        host_info = '{}-{}'.format(host_info, salt)
    return '{hostname}-{hash}'.format(hostname=hostname.split('.')[0], hash=hashlib.md5(host_info.encode('ascii')).hexdigest())
