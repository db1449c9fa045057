"""HMAC-authenticated object RPC over TCP: `BasicService` (server) and `BasicClient`.

Role parity: horovod/runner/common/util/network.py (Wire, BasicService, BasicClient, find_port).  Used by the driver / task /
worker-notification / compute services.

Frame on the wire:  magic "HVB2" | payload length (u32, network order) | HMAC-SHA256(key, payload) | cloudpickle payload.
A frame is verified BEFORE it is unpickled; a bad magic, an oversized length or a wrong digest closes the connection.

Differences from the reference's socketserver-based design: the server is a plain accept loop with one daemon thread per
connection, a connection may carry any number of request / response pairs (clients still open one per call, which keeps
every call idempotent and retryable), listening sockets are bound to port 0 so the kernel picks a free port in one
attempt, and a client fails over to the next vetted address when its first choice stops answering.
"""
import base64
import concurrent.futures
import socket
import struct
import threading

import cloudpickle
import psutil

from horovod_b200.runner.common.util import secret

_MAGIC = b'HVB2'
_HEADER = struct.Struct('!4sI')
_MAX_FRAME = 1 << 30


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


def _read_exact(rfile, n):
    chunks, left = [], n
    while left:
        part = rfile.read(left)
        if not part:
            raise EOFError('connection closed with %d of %d bytes outstanding' % (left, n))
        chunks.append(part)
        left -= len(part)
    return b''.join(chunks)


class Wire(object):
    """Frames objects onto / off a file-like pair (see the module docstring for the layout)."""

    def __init__(self, key):
        self._key = key

    def write(self, obj, wfile):
        payload = cloudpickle.dumps(obj)
        wfile.write(_HEADER.pack(_MAGIC, len(payload)) + secret.compute_digest(self._key, payload) + payload)
        wfile.flush()

    def read(self, rfile):
        magic, length = _HEADER.unpack(_read_exact(rfile, _HEADER.size))
        if magic != _MAGIC or length > _MAX_FRAME:
            raise Exception('Security error: not a frame of this protocol.')
        digest = _read_exact(rfile, secret.DIGEST_LENGTH)
        payload = _read_exact(rfile, length)
        if not secret.check_digest(self._key, payload, digest):
            raise Exception('Security error: digest did not match the message.')
        return cloudpickle.loads(payload)


def get_local_host_addresses():
    """IPv4 addresses of this host's interfaces."""
    out = []
    for intf, addrs in psutil.net_if_addrs().items():
        for a in addrs:
            if a.family == socket.AF_INET:
                out.append(a.address)
    return out


class _Listener(object):
    """Accept loop + one daemon thread per connection.  Quacks enough like a socketserver for `find_port` callers
    (`.socket`, `.shutdown()`, `.server_close()`)."""

    def __init__(self, address, serve_connection):
        self.socket = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.socket.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.socket.bind(address)
        self.socket.listen(128)
        self.socket.settimeout(0.2)
        self._serve_connection = serve_connection
        self._closing = threading.Event()
        self._thread = threading.Thread(target=self._accept_loop, name='hvd-rpc-accept', daemon=True)

    def start(self):
        self._thread.start()

    def _accept_loop(self):
        while not self._closing.is_set():
            try:
                conn, peer = self.socket.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            threading.Thread(target=self._run, args=(conn, peer), name='hvd-rpc-conn', daemon=True).start()

    def _run(self, conn, peer):
        with conn:
            conn.settimeout(None)
            self._serve_connection(conn.makefile('rb'), conn.makefile('wb'), peer)

    def shutdown(self):
        self._closing.set()
        if self._thread.is_alive():
            self._thread.join(timeout=2)

    def server_close(self):
        try:
            self.socket.close()
        except OSError:
            pass


class BasicService(object):
    """Subclasses override `_handle(req, client_address)` and call `super()._handle` for requests they do not know."""

    def __init__(self, service_name, key, nic=None):
        self._service_name = service_name
        self._wire = Wire(key)
        self._nic = nic
        self._server, self._port = find_port(lambda addr: _Listener(addr, self._serve_connection))
        self._addresses = self._get_local_addresses()
        self._server.start()

    def _serve_connection(self, rfile, wfile, peer):
        while True:
            try:
                req = self._wire.read(rfile)
            except Exception:  # noqa: BLE001 - EOF (also port probes), foreign protocol, bad digest: drop the connection
                return
            try:
                resp = self._handle(req, peer)
                if resp is None:
                    raise Exception('Handler did not return a response.')
            except Exception as e:  # noqa: BLE001 - the caller gets the failure instead of a hung socket
                resp = e
            try:
                self._wire.write(resp, wfile)
            except (BrokenPipeError, ConnectionResetError, OSError):
                return

    def _handle(self, req, client_address):
        if isinstance(req, PingRequest):
            return PingResponse(self._service_name, client_address[0])
        raise NotImplementedError(req)

    def _get_local_addresses(self):
        found = {}
        for intf, entries in psutil.net_if_addrs().items():
            if self._nic and intf != self._nic:
                continue
            v4 = [(e.address, self._port) for e in entries if e.family == socket.AF_INET]
            if v4:
                found[intf] = v4
        if self._nic and not found:
            raise NoValidAddressesFound('No available network interface found matching user provided interface: %s' % self._nic)
        return found

    def addresses(self):
        return {intf: list(addrs) for intf, addrs in self._addresses.items()}

    def shutdown(self):
        self._server.shutdown()
        self._server.server_close()

    def get_port(self):
        return self._port


class BasicClient(object):
    """Vets the advertised addresses with a ping at construction; every call opens a connection, sends one request and
    reads one response, so ALL RPCs must be idempotent (they are retried, and re-routed to another vetted address)."""

    def __init__(self, service_name, addresses, key, verbose=0, match_intf=False, probe_timeout=20, attempts=3):
        self._verbose = verbose
        self._service_name = service_name
        self._wire = Wire(key)
        self._match_intf = match_intf
        self._probe_timeout = probe_timeout
        self._attempts = attempts
        self._addresses = self._probe(addresses)
        if not self._addresses:
            raise NoValidAddressesFound('Horovod was unable to connect to {service_name} on any of the following addresses: '
                                        '{addresses}.'.format(service_name=service_name, addresses=addresses))

    def _exchange(self, addr, req, timeout):
        with socket.create_connection(tuple(addr), timeout=timeout) as sock:
            self._wire.write(req, sock.makefile('wb'))
            return self._wire.read(sock.makefile('rb'))

    def _vet(self, intf, addr):
        for _ in range(self._attempts):
            try:
                pong = self._exchange(addr, PingRequest(), self._probe_timeout)
            except Exception:  # noqa: BLE001 - unreachable or not ours
                continue
            if getattr(pong, 'service_name', None) != self._service_name:
                return False
            if self._match_intf:  # the service must have seen us coming from the interface of the same name
                mine = [e.address for e in psutil.net_if_addrs().get(intf, []) if e.family == socket.AF_INET]
                return pong.source_address in mine
            return True
        return False

    def _probe(self, addresses):
        candidates = [(intf, addr) for intf, addrs in addresses.items() for addr in addrs]
        if not candidates:
            return {}
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(32, len(candidates))) as pool:
            verdicts = list(pool.map(lambda c: self._vet(*c), candidates))
        vetted = {}
        for (intf, addr), ok in zip(candidates, verdicts):
            if ok:
                vetted.setdefault(intf, []).append(addr)
        return vetted

    def _send_one(self, addr, req):
        error = None
        for _ in range(self._attempts):
            try:
                resp = self._exchange(addr, req, 60)
            except Exception as e:  # noqa: BLE001
                error = e
                continue
            if isinstance(resp, Exception):
                raise resp
            return resp
        raise error

    def _send(self, req):
        routes = [a for addrs in self._addresses.values() for a in addrs]
        for i, addr in enumerate(routes):
            try:
                return self._send_one(addr, req)
            except (OSError, EOFError):
                if i == len(routes) - 1:
                    raise

    def addresses(self):
        return self._addresses


def find_port(server_factory):
    """-> (server, port): `server_factory((host, port))` on a port chosen by the kernel (bind to 0)."""
    last = None
    for _ in range(8):
        try:
            server = server_factory(('', 0))
            return server, server.socket.getsockname()[1]
        except OSError as e:
            last = e
    raise Exception('Unable to find a port to bind to: %s' % last)
