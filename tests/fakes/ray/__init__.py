"""An in-process stand-in for the few Ray calls RayBackend makes (Ray is not installed in this image): actors are plain
objects, `.remote()` calls run on a thread pool and return futures, placement groups record what was asked for."""
import concurrent.futures
import threading

_pool = concurrent.futures.ThreadPoolExecutor(max_workers=16)
_state = {'cluster': {'CPU': 8, 'GPU': 0}, 'groups': [], 'actors': [], 'killed': [], 'current_pg': None}


class ObjectRef:
    def __init__(self, fut):
        self.fut = fut


class _Method:
    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    def remote(self, *a, **k):
        return ObjectRef(_pool.submit(getattr(self.obj, self.name), *a, **k))


class ActorHandle:
    def __init__(self, obj, options):
        self._obj, self.options = obj, options

    def __getattr__(self, name):
        return _Method(self._obj, name)


class _RemoteClass:
    def __init__(self, cls, options=None):
        self.cls, self._options = cls, dict(options or {})

    def options(self, **kw):
        return _RemoteClass(self.cls, {**self._options, **kw})

    def remote(self, *a, **k):
        h = ActorHandle(self.cls(*a, **k), self._options)
        _state['actors'].append(h)
        return h


def remote(*args, **kwargs):
    if len(args) == 1 and isinstance(args[0], type) and not kwargs:
        return _RemoteClass(args[0])
    return lambda cls: _RemoteClass(cls, kwargs)


def get(refs, timeout=None):
    if isinstance(refs, (list, tuple)):
        return [r.fut.result(timeout) for r in refs]
    return refs.fut.result(timeout)


def wait(refs, timeout=None, num_returns=1):
    done, _ = concurrent.futures.wait([r.fut for r in refs], timeout=timeout)
    ready = [r for r in refs if r.fut in done]
    return ready[:num_returns], [r for r in refs if r not in ready[:num_returns]]


def kill(handle):
    _state['killed'].append(handle)


def available_resources():
    return dict(_state['cluster'])


def nodes():
    return [{'alive': True, 'NodeManagerAddress': '10.0.0.1', 'Resources': dict(_state['cluster'])}]
