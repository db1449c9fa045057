"""The HMAC object RPC (runner/common/util/network.py): framing, authentication, failover, handler errors.
Reference coverage model: test/single/test_service.py."""
import io
import socket
import struct

import pytest

from horovod_b200.runner.common.util import network, secret


class Echo(network.BasicService):
    def __init__(self, key, name='echo'):
        self.calls = 0
        super().__init__(name, key)

    def _handle(self, req, client_address):
        if isinstance(req, dict):
            self.calls += 1
            if req.get('boom'):
                raise ValueError('handler failed on purpose')
            return {'echo': req, 'peer': client_address[0]}
        return super()._handle(req, client_address)


def test_wire_round_trip_and_tamper_detection():
    key = secret.make_secret_key()
    buf = io.BytesIO()
    network.Wire(key).write({'a': [1, 2, 3]}, buf)
    raw = buf.getvalue()
    assert raw[:4] == b'HVB2' and struct.unpack('!I', raw[4:8])[0] == len(raw) - 8 - secret.DIGEST_LENGTH
    assert network.Wire(key).read(io.BytesIO(raw)) == {'a': [1, 2, 3]}
    with pytest.raises(Exception, match='digest'):
        network.Wire(secret.make_secret_key()).read(io.BytesIO(raw))          # other key
    flipped = bytearray(raw)
    flipped[-1] ^= 1
    with pytest.raises(Exception, match='digest'):
        network.Wire(key).read(io.BytesIO(bytes(flipped)))                   # payload modified
    with pytest.raises(Exception, match='not a frame'):
        network.Wire(key).read(io.BytesIO(b'GET / HTTP/1.1\r\n' + raw))
    with pytest.raises(EOFError):
        network.Wire(key).read(io.BytesIO(raw[:-3]))                         # truncated


def test_service_client_calls_errors_and_wrong_key():
    key = secret.make_secret_key()
    svc = Echo(key)
    try:
        client = network.BasicClient('echo', svc.addresses(), key, probe_timeout=5)
        assert client._send({'x': 1})['echo'] == {'x': 1}
        with pytest.raises(ValueError, match='on purpose'):                   # handler exceptions travel to the caller
            client._send({'boom': True})
        assert client._send({'x': 2})['echo'] == {'x': 2}                     # and the service keeps serving
        with pytest.raises(network.NoValidAddressesFound):
            network.BasicClient('echo', svc.addresses(), secret.make_secret_key(), probe_timeout=2, attempts=1)
        with pytest.raises(network.NoValidAddressesFound):
            network.BasicClient('other-service', svc.addresses(), key, probe_timeout=2, attempts=1)
        # several requests on ONE connection
        addr = ('127.0.0.1', svc.get_port())
        with socket.create_connection(addr, timeout=5) as s:
            r, w = s.makefile('rb'), s.makefile('wb')
            wire = network.Wire(key)
            for i in range(3):
                wire.write({'n': i}, w)
                assert wire.read(r)['echo'] == {'n': i}
        # garbage does not take the service down
        with socket.create_connection(addr, timeout=5) as s:
            s.sendall(b'\x00' * 64)
        assert client._send({'x': 3})['echo'] == {'x': 3}
    finally:
        svc.shutdown()


def test_client_fails_over_to_second_address():
    key = secret.make_secret_key()
    a, b = Echo(key), Echo(key)
    try:
        addrs = {'lo': [('127.0.0.1', a.get_port()), ('127.0.0.1', b.get_port())]}
        client = network.BasicClient('echo', addrs, key, probe_timeout=5, attempts=1)
        assert client._send({'k': 1}) and a.calls == 1 and b.calls == 0
        a.shutdown()
        assert client._send({'k': 2})['echo'] == {'k': 2} and b.calls == 1
    finally:
        b.shutdown()


def test_find_port_gives_distinct_bound_servers():
    made = [network.find_port(lambda addr: network._Listener(addr, lambda r, w, p: None)) for _ in range(3)]
    try:
        assert len({port for _, port in made}) == 3 and all(s.socket.getsockname()[1] == port for s, port in made)
    finally:
        for s, _ in made:
            s.server_close()
