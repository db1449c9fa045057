"""Stall inspector, autotuner log and runtime timeline through real np=2 jobs (reference coverage:
test/parallel/test_torch.py::test_timeline_api, test/integration stall tests, docs/autotune)."""
import json
import os

from conftest import run_parallel


def test_stall_inspector_warns_and_names_missing_rank(native_built):
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["stall_warning"],
                           env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1", "HOROVOD_LOG_LEVEL": "warning"})
    assert out.count("STALL WARNING DONE") == 2, out[-3000:]
    assert "waiting for remainder of ranks" in out and "late.tensor: [1]" in out, out[-3000:]


def test_stall_inspector_reports_cached_tensors(native_built):
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["stall_cached"],
                           env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1", "HOROVOD_LOG_LEVEL": "warning"})
    assert out.count("STALL CACHED DONE") == 2, out[-3000:]
    assert "waiting for remainder of ranks" in out and "cached.tensor: [1]" in out, out[-3000:]


def test_static_timeline_env(native_built, tmp_path):
    path = tmp_path / "static_timeline.json"
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["stall_warning"],
                           env={"HOROVOD_TIMELINE": str(path), "HOROVOD_STALL_CHECK_DISABLE": "1"})
    assert out.count("STALL WARNING DONE") == 2
    raw = path.read_text().strip()
    events = json.loads(raw if raw.endswith("]") else raw.rstrip(",") + "]")
    assert any(isinstance(e, dict) and "ALLREDUCE" in str(e.get("name")) for e in events)


def test_stall_inspector_shutdown(native_built):
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["stall_shutdown"], expect_fail=True,
                           env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1", "HOROVOD_STALL_SHUTDOWN_TIME_SECONDS": "3",
                                "HOROVOD_LOG_LEVEL": "warning"})
    assert "STALL SHUTDOWN RAISED" in out and "Will shutdown" in out, out[-3000:]


def test_stall_check_can_be_disabled(native_built):
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["stall_warning"],
                           env={"HOROVOD_STALL_CHECK_TIME_SECONDS": "1", "HOROVOD_STALL_CHECK_DISABLE": "1", "HOROVOD_LOG_LEVEL": "warning"})
    assert out.count("STALL WARNING DONE") == 2 and "waiting for remainder of ranks" not in out, out[-3000:]


def test_autotune_writes_log_and_keeps_ranks_consistent(native_built, tmp_path):
    log = tmp_path / "autotune.csv"
    rc, out = run_parallel("aux_worker.py", np=2, timeout=300, args=["autotune"],
                           env={"HOROVOD_AUTOTUNE": "1", "HOROVOD_AUTOTUNE_LOG": str(log), "HOROVOD_AUTOTUNE_WARMUP_SAMPLES": "1",
                                "HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE": "5", "HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES": "6"})
    params = [json.loads(l.split("AUTOTUNE PARAMS", 1)[1]) for l in out.splitlines() if "AUTOTUNE PARAMS" in l]
    assert len(params) == 2 and params[0] == params[1], out[-2000:]      # every rank ends with rank 0's parameters
    text = log.read_text().strip().splitlines()
    assert len(text) >= 3 and "fusion" in text[0].lower() and "score" in text[0].lower(), text[:3]


def test_runtime_timeline_is_valid_chrome_trace(native_built, tmp_path):
    path = tmp_path / "timeline.json"
    rc, out = run_parallel("aux_worker.py", np=2, timeout=120, args=["timeline", str(path)])
    assert out.count("TIMELINE DONE") == 2, out[-2000:]
    assert "TIMELINE LIVE JSON OK" in out, out[-2000:]
    events = json.loads(path.read_text())  # strict: a complete document, no repair
    names = {e.get("name") for e in events if isinstance(e, dict)}
    assert any(n and "NEGOTIATE" in n for n in names), sorted(n for n in names if n)[:20]
    assert any(n and "ALLREDUCE" in n for n in names) and any(n and "ALLGATHER" in n for n in names)
    assert any(n and "CYCLE_START" in n for n in names)
    tids = {e.get("args", {}).get("name") for e in events if isinstance(e, dict) and e.get("ph") == "M"}
    assert any(t and "tl.ar" in t for t in tids), tids


def test_metrics_counters_and_exporters(native_built):
    """hvd.metrics(): per-collective counters (fusion visible as responses < tensors, error responses counted), Interval deltas,
    Prometheus text endpoint with rank labels, periodic log line."""
    rc, out = run_parallel("metrics_worker.py", np=2, timeout=200)
    assert "METRICS OK" in out, out[-3000:]
