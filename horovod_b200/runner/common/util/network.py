"""HMAC-authenticated pickle RPC over TCP: BasicService (threaded server) and BasicClient.

Every message is `digest(32) | length(4) | cloudpickle payload`; a message whose digest does not match the
shared secret is dropped before unpickling.  Used by the driver / task / worker-notification / compute services.
Role parity: horovod/runner/common/util/network.py.
"""
import base64
import io
import queue
import random
import socket
import socketserver
import struct
import threading

import cloudpickle
import psutil

from horovod_b200.runner.common.util import secret


class PingRequest(object):
    pass


class NoValidAddressesFound(Exception):
    pass


class PingResponse(object):
    def __init__(self, service_name, source_address):
        self.service_name = service_name
        self.source_address = source_address


class AckResponse(object):
    """Used for situations when the response does not carry any data."""
    pass


class AckStreamResponse(object):
    pass


def dumps_base64(obj):
    return base64.b64encode(cloudpickle.dumps(obj)).decode('ascii')


def loads_base64(text):
    return cloudpickle.loads(base64.b64decode(text.encode('ascii')))


class Wire(object):
    """Wire format: digest + length-prefixed pickle."""

    def __init__(self, key):
        self._key = key

    def write(self, obj, wfile):
        message = cloudpickle.dumps(obj)
        digest = secret.compute_digest(self._key, message)
        wfile.write(digest)
        wfile.write(struct.pack('i', len(message)))
        wfile.write(message)
        wfile.flush()

    def read(self, rfile):
        digest = rfile.read(secret.DIGEST_LENGTH)
        raw = rfile.read(4)
        if len(digest) != secret.DIGEST_LENGTH or len(raw) != 4:
            raise EOFError('connection closed')
        message_len = struct.unpack('i', raw)[0]
        message = rfile.read(message_len)
        if not secret.check_digest(self._key, message, digest):
            raise Exception('Security error: digest did not match the message.')
        return cloudpickle.loads(message)


def get_local_host_addresses():
    """IPv4 addresses of this host's interfaces."""
    out = []
    for intf, addrs in psutil.net_if_addrs().items():
        for a in addrs:
            if a.family == socket.AF_INET:
                out.append(a.address)
    return out


class BasicService(object):
    def __init__(self, service_name, key, nic=None):
        self._service_name = service_name
        self._wire = Wire(key)
        self._nic = nic
        self._server, _ = find_port(lambda addr: socketserver.ThreadingTCPServer(addr, self._make_handler()))
        self._server.daemon_threads = True
        self._port = self._server.socket.getsockname()[1]
        self._addresses = self._get_local_addresses()
        self._thread = threading.Thread(target=self._server.serve_forever, kwargs={'poll_interval': 0.1}, daemon=True)
        self._thread.start()

    def _make_handler(self):
        server = self

        class _Handler(socketserver.StreamRequestHandler):
            def handle(self):
                try:
                    req = server._wire.read(self.rfile)
                    resp = server._handle(req, self.client_address)
                    if resp is None:
                        raise Exception('Handler did not return a response.')
                    server._wire.write(resp, self.wfile)
                except (EOFError, BrokenPipeError, ConnectionResetError):
                    pass  # happens when the client is probing for open ports

        return _Handler

    def _handle(self, req, client_address):
        if isinstance(req, PingRequest):
            return PingResponse(self._service_name, client_address[0])
        raise NotImplementedError(req)

    def _get_local_addresses(self):
        result = {}
        for intf, intf_addresses in psutil.net_if_addrs().items():
            if self._nic and intf != self._nic:
                continue
            for addr in intf_addresses:
                if addr.family == socket.AF_INET:
                    result.setdefault(intf, []).append((addr.address, self._port))
        if not result and self._nic:
            raise NoValidAddressesFound(f'No available network interface found matching user provided interface: {self._nic}')
        return result

    def addresses(self):
        return self._addresses.copy()

    def shutdown(self):
        self._server.shutdown()
        self._server.server_close()
        self._thread.join(timeout=2)

    def get_port(self):
        return self._port


class BasicClient(object):
    def __init__(self, service_name, addresses, key, verbose=0, match_intf=False, probe_timeout=20, attempts=3):
        # Note: because of retry logic, ALL RPC calls are REQUIRED to be idempotent.
        self._verbose = verbose
        self._service_name = service_name
        self._wire = Wire(key)
        self._match_intf = match_intf
        self._probe_timeout = probe_timeout
        self._attempts = attempts
        self._addresses = self._probe(addresses)
        if not self._addresses:
            raise NoValidAddressesFound(
                'Horovod was unable to connect to {service_name} on any of the following addresses: {addresses}.'.format(
                    service_name=service_name, addresses=addresses))

    def _probe(self, addresses):
        result_queue = queue.Queue()
        threads = []
        for intf, intf_addresses in addresses.items():
            for addr in intf_addresses:
                t = threading.Thread(target=self._probe_one, args=(intf, addr, result_queue), daemon=True)
                t.start()
                threads.append(t)
        for t in threads:
            t.join(self._probe_timeout)
        result = {}
        while not result_queue.empty():
            intf, addr = result_queue.get()
            result.setdefault(intf, []).append(addr)
        return result

    def _probe_one(self, intf, addr, result_queue):
        for _ in range(self._attempts):
            try:
                with socket.create_connection(addr, timeout=self._probe_timeout) as sock:
                    rfile = sock.makefile('rb')
                    wfile = sock.makefile('wb')
                    self._wire.write(PingRequest(), wfile)
                    resp = self._wire.read(rfile)
                    if resp.service_name != self._service_name:
                        return
                    if self._match_intf:
                        # Interface name of destination and source must match since `match_intf` is requested.
                        client_intf_addrs = [x.address for x in psutil.net_if_addrs().get(intf, []) if x.family == socket.AF_INET]
                        if resp.source_address not in client_intf_addrs:
                            return
                    result_queue.put((intf, addr))
                    return
            except Exception:
                continue

    def _send_one(self, addr, req):
        for attempt in range(self._attempts):
            try:
                with socket.create_connection(addr, timeout=60) as sock:
                    rfile = sock.makefile('rb')
                    wfile = sock.makefile('wb')
                    self._wire.write(req, wfile)
                    return self._wire.read(rfile)
            except Exception:
                if attempt == self._attempts - 1:
                    raise

    def _send(self, req):
        # Since all the addresses were vetted, use the first one.
        addr = list(self._addresses.values())[0][0]
        return self._send_one(addr, req)

    def addresses(self):
        return self._addresses


def find_port(server_factory):
    min_port, max_port = 1024, 65536
    num_ports = max_port - min_port
    start_port = random.randrange(0, num_ports)
    for port_offset in range(num_ports):
        try:
            port = min_port + (start_port + port_offset) % num_ports
            addr = ('', port)
            server = server_factory(addr)
            return server, port
        except Exception:
            pass
    raise Exception('Unable to find a port to bind to.')
