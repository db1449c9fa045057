"""np=2 (and 3, 4) CPU runs of the distributed op matrix through our own launcher — the "plumbing" config of
BASELINE.json (`synthetic allreduce correctness world_size=2 on CPU`)."""
import pytest

from conftest import run_parallel


@pytest.mark.parametrize("np_", [2])
def test_ops_matrix_np2(native_built, np_):
    rc, out = run_parallel("ops_worker.py", np=np_, timeout=400)
    assert "ALL OK" in out, out[-3000:]


def test_ops_matrix_np3_non_power_of_two(native_built):
    rc, out = run_parallel("ops_worker.py", np=3, timeout=400,
                           args=["--only", "rank_size,allreduce_sum_avg,allreduce_async_fused,allgather,broadcast,alltoall,"
                                 "reducescatter,process_sets,errors,barrier_join,optimizer"])
    assert "ALL OK" in out, out[-3000:]


def test_tcp_control_plane(native_built):
    """Same matrix subset with the shared-memory control plane disabled (pure TCP negotiation)."""
    rc, out = run_parallel("ops_worker.py", np=2, timeout=400, env={"HVD_CONTROL_PLANE": "tcp"},
                           args=["--only", "allreduce_sum_avg,allreduce_async_fused,cache_invalidation,barrier_join"])
    assert "ALL OK" in out, out[-3000:]


def test_cache_disabled_and_no_fusion(native_built):
    rc, out = run_parallel("ops_worker.py", np=2, timeout=400,
                           env={"HOROVOD_CACHE_CAPACITY": "0", "HOROVOD_FUSION_THRESHOLD": "0"},
                           args=["--only", "allreduce_sum_avg,allreduce_async_fused,grouped_allreduce,barrier_join"])
    assert "ALL OK" in out, out[-3000:]


def test_torchrun_env_bootstrap(native_built, tmp_path):
    """hvd.init() under a torchrun-style environment (RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT, no hvdrun)."""
    import os, socket, subprocess, sys
    from conftest import REPO
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text("import torch\nimport horovod_b200.torch as hvd\nhvd.init()\n"
                      "out = hvd.allreduce(torch.ones(4) * (hvd.rank() + 1), op=hvd.Sum)\n"
                      "assert out.tolist() == [3.0] * 4, out\nassert hvd.local_size() == 2\nprint('OK', hvd.rank())\nhvd.shutdown()\n")
    procs = []
    for r in range(2):
        e = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), LOCAL_WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                 MASTER_PORT=str(port), PYTHONPATH=REPO)
        for k in list(e):
            if k.startswith("HOROVOD_"):
                del e[k]
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_fake_two_hosts_np4(native_built):
    """4 ranks presented as 2 hosts x 2: TCP control plane, local/cross topology, multi-host CPU data plane."""
    rc, out = run_parallel("ops_worker.py", np=4, timeout=400, env={"HVD_TEST_FAKE_HOSTS": "2"},
                           args=["--only", "rank_size,fake_hosts_topology,allreduce_sum_avg,allreduce_async_fused,allgather,"
                                 "broadcast,process_sets,barrier_join"])
    assert "ALL OK" in out, out[-3000:]


@pytest.mark.parametrize("np_,hosts", [(3, "3"), (4, "2")])
def test_pipelined_tcp_ring_allreduce(native_built, np_, hosts):
    """Cross-host ring with a tiny HVD_RING_CHUNK_BYTES: every ring step is cut into many chunks and the reducer thread folds
    chunk k while chunk k+1 is on the wire (cpu_ops.cc:RingAllreduce).  3 ranks on 3 "hosts" = the plain ring with uneven
    segments; 4 ranks on 2 "hosts" = the cross-host rings of the two-level allreduce.  Exact integer-valued sums.  With three
    ranks the broadcasts of >= 4 chunks take the chunked chain instead of the binomial tree (cpu_ops.cc:TreeBroadcast)."""
    rc, out = run_parallel("ops_worker.py", np=np_, timeout=400, env={"HVD_TEST_FAKE_HOSTS": hosts, "HVD_RING_CHUNK_BYTES": "8192"},
                           args=["--only", "rank_size,allreduce_sum_avg,allreduce_min_max_product,allreduce_mixed_dtype_fusion,"
                                 "large_allreduce,reducescatter,broadcast,objects_and_state"])
    assert "ALL OK" in out, out[-3000:]


@pytest.mark.parametrize("np_,env", [(3, {"HVD_CONTROL_PLANE": "tcp"}), (4, {"HVD_CONTROL_PLANE": "tcp"}), (3, {"HVD_TEST_FAKE_HOSTS": "3"})])
def test_recursive_doubling_bit_reduction(native_built, np_, env):
    """HVD_BITS_TREE_MIN_RANKS=2 forces the log-depth exchange of the negotiation bit vectors (default: from 9 members on) for
    3 members (one extra rank folded into a power-of-two core) and 4 members, over the plain TCP plane and among host leaders."""
    rc, out = run_parallel("ops_worker.py", np=np_, timeout=400, env=dict(env, HVD_BITS_TREE_MIN_RANKS="2"),
                           args=["--only", "rank_size,allreduce_sum_avg,allreduce_async_fused,grouped_allreduce,allgather,broadcast,"
                                 "process_sets,errors,cache_invalidation,barrier_join"])
    assert "ALL OK" in out, out[-3000:]


def test_fake_three_hosts_np6(native_built):
    """6 ranks presented as 3 hosts x 2: two-level planes with an odd number of hosts (leader star of three, cross-host rings
    of three, chain broadcast over three column ranks), host-local process sets on every host, every host collective."""
    rc, out = run_parallel("ops_worker.py", np=6, timeout=400, env={"HVD_TEST_FAKE_HOSTS": "3"},
                           args=["--only", "rank_size,fake_hosts_topology,allreduce_sum_avg,allreduce_async_fused,allgather,broadcast,"
                                 "alltoall,reducescatter,process_sets,barrier_join,large_allreduce"])
    assert "ALL OK" in out, out[-3000:]


def test_numpy_frontend_np2(native_built):
    rc, out = run_parallel("numpy_worker.py", np=2, timeout=200)
    assert "NUMPY OK" in out, out[-3000:]


def test_tensorflow_frontend_against_fake_tf_np2(native_built):
    """Control flow of horovod_b200.tensorflow (+ keras callbacks) over a numpy-backed TensorFlow stand-in
    (tests/fakes/tensorflow): TensorFlow itself is not installed in this image."""
    rc, out = run_parallel("tf_fake_worker.py", np=2, timeout=200)
    assert "TF FAKE OK" in out, out[-3000:]


def test_tensorflow_frontend_import_error_without_tf():
    import importlib, sys
    if importlib.util.find_spec("tensorflow") is not None:
        pytest.skip("TensorFlow is installed")
    for m in [k for k in sys.modules if k.startswith("horovod_b200.tensorflow")]:
        del sys.modules[m]
    with pytest.raises(ImportError, match="TensorFlow"):
        importlib.import_module("horovod_b200.tensorflow")


@pytest.mark.parametrize("np_", [2, 3])
def test_extra_reference_cases(native_built, np_):
    """Second matrix: grad variants with process sets, grouped allgather / reducescatter, per-op error paths, sparse
    gradients, optimizer corner cases, join with non-allreduce ops, barriers mixed with collectives."""
    rc, out = run_parallel("ops_worker_extra.py", np=np_, timeout=300)
    assert "EXTRA ALL OK" in out, out[-3000:]


def test_parallel_package_np4(native_built):
    """horovod_b200.parallel: local / cross / 2-D mesh process sets (4 ranks shown as 2 hosts x 2), shard helpers,
    ShardedSGD == DistributedOptimizer(SGD)."""
    rc, out = run_parallel("parallel_pkg_worker.py", np=4, timeout=300)
    assert "PARALLEL PKG OK" in out, out[-3000:]
    assert "background loop failed" not in out, out[-3000:]   # shutdown with sub-sets registered must be clean


def test_mxnet_frontend_against_fake_mxnet_np2(native_built):
    """Control flow of horovod_b200.mxnet (ops, DistributedOptimizer, DistributedTrainer, broadcast_parameters incl.
    deferred initialisation) over a numpy-backed MXNet stand-in (tests/fakes/mxnet)."""
    rc, out = run_parallel("mx_fake_worker.py", np=2, timeout=200)
    assert "MX FAKE OK" in out, out[-3000:]


def test_init_with_rank_subset_np3(native_built):
    """hvd.init(comm=[2, 0]) on two of three launched ranks (renumbered in list order), hvd.init(comm=[1]) on the third;
    a non-member is rejected (reference: common/basics.py init(comm=<rank list>))."""
    rc, out = run_parallel("comm_subset_worker.py", np=3, timeout=200)
    assert out.count("COMM SUBSET OK") == 3, out[-3000:]


def test_peer_shutdown_semantics_np2(native_built):
    """One rank shuts the job down: the other rank's rank()/size() stay valid until ITS shutdown, collectives fail with the
    shut-down error (reference operations.cc: initialization_done survives the loop exit)."""
    rc, out = run_parallel("peer_shutdown_worker.py", np=2, timeout=120)
    assert "rank0 done" in out and "rank1 done" in out, out[-3000:]


@pytest.mark.parametrize("np_", [2, 3])
def test_op_api_edge_cases(native_built, np_):
    """Empty / 0-dim / bool tensors, 300 ops in flight, tensors above a 64 KiB fusion threshold, name reuse across op
    types, non-contiguous inputs, exact int64 sums."""
    rc, out = run_parallel("edge_worker.py", np=np_, timeout=200, env={"HOROVOD_FUSION_THRESHOLD": "65536"})
    assert "EDGE OK" in out, out[-3000:]


def test_join_with_cached_responses_np3(native_built):
    """Ranks that joined keep serving responses that live in the cache (bit-vector fast path): the active rank's sums contain
    only its own contribution, Average still divides by the full size, and the cache entry survives the join."""
    rc, out = run_parallel("join_cached_worker.py", np=3, timeout=120)
    assert "JOIN CACHED OK" in out, out[-3000:]


def test_static_process_sets_np4(native_built):
    """hvd.init(process_sets=[even, odd]): ids, membership, collectives inside a set, non-member rejection, duplicate
    rejection, re-init with the same static sets."""
    rc, out = run_parallel("static_sets_worker.py", np=4, timeout=200)
    assert "STATIC SETS OK" in out, out[-3000:]


def test_broadcast_optimizer_state_every_torch_optimizer(native_built):
    """All 13 optimizer classes torch ships (LBFGS / SparseAdam excluded): different hyper-parameters and state per rank,
    equal to the root's afterwards; state on the root only; wrapped optimizer."""
    rc, out = run_parallel("optim_state_worker.py", np=2, timeout=300)
    assert "OPTIM STATE OK" in out, out[-3000:]


@pytest.mark.parametrize("np_,env", [(2, {"HVD_SHM_SLOT_BYTES": "4096"}), (3, {"HVD_SHM_SLOT_BYTES": "8192"}),
                                     (3, {"HVD_SHM_DATA_PLANE": "0"}), (4, {"HVD_SHM_SLOT_BYTES": "16384"}),
                                     (2, {"HVD_SHM_SLOT_BYTES": "4194304", "HVD_CPU_THREADS": "3"}),
                                     (4, {"HVD_TEST_FAKE_HOSTS": "2", "HVD_SHM_SLOT_BYTES": "8192"}),
                                     (3, {"HVD_TEST_FAKE_HOST_MAP": "0,0,1"})])
def test_shared_memory_data_plane(native_built, np_, env):
    """Host-tensor collectives through the shm slots with a tiny slot size (many pieces, double buffering across different
    collectives), the same program with the data plane / the whole shm overlay switched off, and with the 4 ranks
    presented as 2 hosts (two-level control plane: shm inside a host, leaders over TCP; two-level data plane: shm inside a
    host + one cross-host ring per local rank, tiny slots), and 3 ranks on uneven hosts (2 + 1: two-level
    control, plain ring for the data)."""
    rc, out = run_parallel("shm_plane_worker.py", np=np_, timeout=400, env=env)
    assert "SHM PLANE OK" in out, out[-3000:]


def test_reference_behavioural_details(native_built):
    rc, out = run_parallel("appendix_a_worker.py", np=3, timeout=200)
    assert "APPENDIX A OK dup_err=True" in out, out[-3000:]


def test_half_precision_host_reductions_are_bit_exact(native_built):
    """fp16 (F16C) and bf16 (auto-vectorised) host reductions against torch's fp32-op-then-round result, bit for bit, including
    inf / NaN / subnormals / signed zeros and odd lengths (scalar tails)."""
    rc, out = run_parallel("half_exact_worker.py", np=2, timeout=300)
    assert "HALF EXACT OK" in out, out[-3000:]


def test_shm_wait_timeout_on_a_stopped_peer(native_built):
    """HVD_SHM_TIMEOUT_SECONDS: rank 1 SIGSTOPs itself (alive, but its cycle thread is frozen); rank 0 gets a HorovodInternalError
    after the timeout instead of spinning forever."""
    rc, out = run_parallel("shm_timeout_worker.py", np=2, timeout=120, env={"HVD_SHM_TIMEOUT_SECONDS": "3"}, expect_fail=True)
    assert "TIMEOUT RAISED" in out and "NO ERROR" not in out, out[-3000:]


def test_tcp_wait_timeout_on_a_stopped_peer(native_built):
    """HVD_TCP_TIMEOUT_SECONDS: the same frozen peer with every negotiation round over sockets (HVD_CONTROL_PLANE=tcp, what a
    multi-host job uses across hosts): a stopped process never closes its socket, the receive has to give up on its own."""
    rc, out = run_parallel("shm_timeout_worker.py", np=2, timeout=120, expect_fail=True,
                           env={"HVD_TCP_TIMEOUT_SECONDS": "3", "HVD_CONTROL_PLANE": "tcp"})
    assert "TIMEOUT RAISED" in out and "NO ERROR" not in out, out[-3000:]
