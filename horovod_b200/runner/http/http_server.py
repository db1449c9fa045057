"""Threaded HTTP key-value store used for rendezvous and for run-func result passing.

`GET /scope/key` -> 200 + value | 404;  `PUT /scope/key` stores the body;
`DELETE /scope/` marks one participant of the scope finished (the scope is dropped
once every participant did).  The C++ runtime talks to it through
csrc/transport/tcp_transport.cc:HttpKVStore.

Role parity: horovod/runner/http/http_server.py (KVStoreHandler, RendezvousHandler,
RendezvousServer, KVStoreServer).
"""
import collections
import logging
import socketserver
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer

OK = 200
NOT_FOUND = 404
BAD_REQUEST = 400


class KVStoreHandler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.0"

    def _split(self):
        parts = self.path.split('/')
        if len(parts) < 3:
            return None, None
        return parts[1], '/'.join(parts[2:])

    def do_GET(self):
        scope, key = self._split()
        if scope is None:
            return self._reply(BAD_REQUEST)
        value = self.server.handle_get(scope, key, self)
        if value is None:
            return self._reply(NOT_FOUND)
        self._reply(OK, value)

    def do_PUT(self):
        scope, key = self._split()
        if scope is None:
            return self._reply(BAD_REQUEST)
        n = int(self.headers.get('Content-Length', 0))
        value = self.rfile.read(n) if n else b''
        self.server.handle_put(scope, key, value, self)
        self._reply(OK)

    def do_DELETE(self):
        scope, key = self._split()
        if scope is None:
            return self._reply(BAD_REQUEST)
        self.server.handle_delete(scope, key)
        self._reply(OK)

    def _reply(self, code, body=b''):
        self.send_response(code)
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        if body:
            self.wfile.write(body)

    def log_message(self, fmt, *args):  # silence per-request logging
        logging.debug(fmt, *args)


class _ThreadedHTTPServer(socketserver.ThreadingMixIn, HTTPServer):
    daemon_threads = True
    allow_reuse_address = True
    request_queue_size = 128


class KVStoreHTTPServer(_ThreadedHTTPServer):
    def __init__(self, addr, handler_cls=KVStoreHandler):
        super().__init__(addr, handler_cls)
        self.cache_lock = threading.Lock()
        self.cache = collections.defaultdict(dict)
        self.get_hooks = {}   # scope -> fn(key, handler) -> bytes | None   (elastic rendezvous)
        self.put_hooks = {}   # scope -> fn(key, value)

    def handle_get(self, scope, key, handler):
        hook = self.get_hooks.get(scope)
        if hook is not None:
            return hook(key, handler)
        with self.cache_lock:
            return self.cache.get(scope, {}).get(key)

    def handle_put(self, scope, key, value, handler):
        with self.cache_lock:
            self.cache[scope][key] = value
        hook = self.put_hooks.get(scope)
        if hook is not None:
            hook(key, value)

    def handle_delete(self, scope, key):
        with self.cache_lock:
            self.cache.pop(scope, None)


class KVStoreServer:
    """Plain KV server (run-func mode results, generic rendezvous)."""

    def __init__(self, verbose=0):
        self.httpd = None
        self.thread = None
        self.verbose = verbose

    def start_server(self, port=0, host=''):
        self.httpd = KVStoreHTTPServer((host, port))
        self.thread = threading.Thread(target=self.httpd.serve_forever, kwargs={'poll_interval': 0.1}, daemon=True)
        self.thread.start()
        return self.httpd.server_address[1]

    @property
    def port(self):
        return self.httpd.server_address[1]

    def get(self, scope, key):
        with self.httpd.cache_lock:
            return self.httpd.cache.get(scope, {}).get(key)

    def put(self, scope, key, value):
        with self.httpd.cache_lock:
            self.httpd.cache[scope][key] = value

    def clear(self, scope_prefix=None):
        with self.httpd.cache_lock:
            if scope_prefix is None:
                self.httpd.cache.clear()
            else:
                for s in [s for s in self.httpd.cache if s.startswith(scope_prefix)]:
                    del self.httpd.cache[s]

    def shutdown_server(self):
        if self.httpd is not None:
            self.httpd.shutdown()
            self.httpd.server_close()
            self.httpd = None
        if self.thread is not None:
            self.thread.join(timeout=2)
            self.thread = None


class RendezvousServer(KVStoreServer):
    """KV server + the host/slot plan of the job (reference RendezvousServer.init(host_alloc_plan))."""

    def __init__(self, verbose=0):
        super().__init__(verbose)
        self.slot_info = []
        self.round = 0

    def init(self, host_alloc_plan):
        """Publishes a new allocation plan; every call starts a new rendezvous round."""
        self.slot_info = list(host_alloc_plan)
        self.round += 1
        if self.httpd is not None:
            self.clear('mesh.')
        return self.round

    def stop(self):
        self.shutdown_server()
