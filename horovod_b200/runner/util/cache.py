"""On-disk cache of launcher pre-flight checks (ssh reachability, common NICs) — reference runner/util/cache.py."""
import datetime
import os
import pickle
import threading


class Cache(object):
    """Results are stale after `cache_staleness_threshold_in_minutes` or when the launch parameters change."""

    def __init__(self, cache_folder, cache_staleness_threshold_in_minutes, parameters_hash):
        self._cache_file = os.path.join(cache_folder, 'cache.bin')
        os.makedirs(cache_folder, exist_ok=True)
        if not os.path.isfile(self._cache_file) or os.path.getsize(self._cache_file) == 0:
            self._content = {'parameters_hash': parameters_hash}
            self._dump()
        else:
            try:
                with open(self._cache_file, 'rb') as cf:
                    content = pickle.load(cf)
            except Exception:
                content = {}
            if content.get('parameters_hash', None) == parameters_hash:
                self._content = content
            else:
                self._content = {'parameters_hash': parameters_hash}
                self._dump()
        self._cache_staleness_threshold = datetime.timedelta(minutes=cache_staleness_threshold_in_minutes)
        self._lock = threading.Lock()

    def _dump(self):
        with open(self._cache_file, 'wb') as cf:
            pickle.dump(self._content, cf)

    def get(self, key):
        with self._lock:
            timestamp, val = self._content.get(key, (None, None))
        if timestamp and timestamp >= datetime.datetime.now() - self._cache_staleness_threshold:
            return val
        return None

    def put(self, key, val):
        with self._lock:
            self._content[key] = (datetime.datetime.now(), val)
            try:
                self._dump()
            except Exception as e:
                print('There is an error with writing to cache file: {}'.format(e))

    def use_cache(self):
        """Decorator: memoises a function whose positional args are hashable-by-repr."""
        def wrap(func):
            def wrap_f(*args, **kwargs):
                key = (func.__name__, repr(args[:-1] if args and isinstance(args[-1], Cache) else args), repr(sorted(kwargs.items())))
                cached_result = self.get(key)
                if cached_result is not None:
                    return cached_result
                result = func(*args, **kwargs)
                if result:  # only cache successes
                    self.put(key, result)
                return result
            return wrap_f
        return wrap
