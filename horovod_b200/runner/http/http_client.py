"""Tiny HTTP KV client for Python-side rendezvous traffic (reference runner/http/http_client.py)."""
import time
import urllib.error
import urllib.request


def put_data_into_kvstore(addr, port, scope, key, value, timeout=10):
    if isinstance(value, str):
        value = value.encode()
    req = urllib.request.Request(f"http://{addr}:{port}/{scope}/{key}", data=value, method='PUT')
    with urllib.request.urlopen(req, timeout=timeout) as r:
        return r.status


def read_data_from_kvstore(addr, port, scope, key, timeout=60.0, poll=0.02, request_timeout=None):
    deadline = time.time() + timeout
    url = f"http://{addr}:{port}/{scope}/{key}"
    while True:
        try:
            with urllib.request.urlopen(url, timeout=request_timeout or 10) as r:
                return r.read()
        except urllib.error.HTTPError as e:
            if e.code != 404:
                raise
        except (urllib.error.URLError, ConnectionError, TimeoutError):
            pass
        if time.time() > deadline:
            raise TimeoutError(f"timed out reading {scope}/{key} from the KV store at {addr}:{port}")
        time.sleep(poll)
