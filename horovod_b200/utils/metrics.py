"""Exporters for the runtime counters (`hvd.metrics()`).

The reference has no metrics surface ("No Prometheus/StatsD-style metrics"); for a production deployment the
questions "how many bytes did this rank reduce, on which path, how many negotiation cycles were idle, did anything error"
should not need a timeline file.

    from horovod_b200.utils import metrics
    metrics.start_prometheus(port=9400 + hvd.local_rank())      # /metrics, one endpoint per rank
    with metrics.Interval() as m: train_epoch(); print(m.delta['allreduce']['bytes'])
    metrics.log_every(60)                                         # one line per minute through `logging`
"""
import logging
import threading
import time

_FIELDS = ('responses', 'tensors', 'bytes', 'on_gpu', 'errors')


def snapshot():
    from horovod_b200.common.basics import HorovodBasics  # noqa: F401  (import cost only when used)
    import horovod_b200.torch as hvd
    return hvd.metrics()


def flatten(snap, prefix='hvd'):
    """{'allreduce': {'bytes': 3}} -> {'hvd_allreduce_bytes': 3}"""
    return {'%s_%s_%s' % (prefix, group, key): value for group, vals in snap.items() for key, value in vals.items()}


def diff(new, old):
    out = {}
    for group, vals in new.items():
        base = old.get(group, {})
        out[group] = {k: v - base.get(k, 0) for k, v in vals.items()}
    return out


class Interval:
    """Context manager: `.delta` holds the counters accumulated inside the block, `.seconds` its duration."""

    def __init__(self, snapshot_fn=None):
        self._snap = snapshot_fn or snapshot
        self.delta, self.seconds = {}, 0.0

    def __enter__(self):
        self._t0, self._before = time.perf_counter(), self._snap()
        return self

    def __exit__(self, *exc):
        self.delta, self.seconds = diff(self._snap(), self._before), time.perf_counter() - self._t0
        return False

    def rate(self, group='allreduce', key='bytes'):
        return self.delta.get(group, {}).get(key, 0) / self.seconds if self.seconds > 0 else 0.0


def prometheus_text(snap=None, labels=None):
    """The counters in the Prometheus text exposition format (no dependency on prometheus_client)."""
    snap = snapshot() if snap is None else snap
    label = ''
    if labels:
        label = '{' + ','.join('%s="%s"' % (k, str(v).replace('"', '\\"')) for k, v in sorted(labels.items())) + '}'
    lines = []
    for name, value in sorted(flatten(snap).items()):
        lines.append('# TYPE %s counter' % name)
        lines.append('%s%s %d' % (name, label, value))
    return '\n'.join(lines) + '\n'


def start_prometheus(port, addr='', labels=None, snapshot_fn=None):
    """Serves `prometheus_text()` on http://addr:port/metrics from a daemon thread; returns the server (`.shutdown()` stops
    it).  `labels` default to this rank's coordinates."""
    import http.server
    snap_fn = snapshot_fn or snapshot
    if labels is None:
        try:
            import horovod_b200.torch as hvd
            labels = {'rank': hvd.rank(), 'local_rank': hvd.local_rank(), 'size': hvd.size()}
        except Exception:  # noqa: BLE001 - not initialised yet: serve without labels
            labels = {}

    class Handler(http.server.BaseHTTPRequestHandler):
        def do_GET(self):  # noqa: N802
            if self.path.split('?')[0] not in ('/metrics', '/'):
                self.send_error(404)
                return
            try:
                body = prometheus_text(snap_fn(), labels).encode()
            except Exception as e:  # noqa: BLE001 - e.g. scraped after hvd.shutdown()
                self.send_error(503, str(e))
                return
            self.send_response(200)
            self.send_header('Content-Type', 'text/plain; version=0.0.4')
            self.send_header('Content-Length', str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *args):
            pass
    server = http.server.ThreadingHTTPServer((addr, port), Handler)
    server.daemon_threads = True
    threading.Thread(target=server.serve_forever, name='hvd-metrics', daemon=True).start()
    return server


def log_every(seconds, logger=None, snapshot_fn=None):
    """Logs the counters accumulated in each period; returns a `threading.Event` — set it to stop."""
    log = logger or logging.getLogger('horovod_b200.metrics')
    snap_fn = snapshot_fn or snapshot
    stop = threading.Event()

    def loop():
        before = snap_fn()
        while not stop.wait(seconds):
            try:
                now = snap_fn()
            except Exception:  # noqa: BLE001 - runtime gone
                return
            d = diff(now, before)
            before = now
            parts = ['%s: %d ops / %d tensors / %.1f MB' % (g, v.get('responses', 0), v.get('tensors', 0), v.get('bytes', 0) / 1e6)
                     for g, v in sorted(d.items()) if g != 'runtime' and v.get('responses')]
            rt = d.get('runtime', {})
            log.info('last %.0f s: %s; cycles %d (%d idle), kernel launches %d', seconds, '; '.join(parts) or 'no collectives',
                     rt.get('cycles', 0), rt.get('idle_cycles', 0), rt.get('kernel_launches', 0))
    threading.Thread(target=loop, name='hvd-metrics-log', daemon=True).start()
    return stop
