"""Parallelism building blocks on top of the collectives.

The reference implements data parallelism only; what it offers for everything else are the primitives:
process sets, alltoall(v), reducescatter + allgather.  This package keeps that scope and adds the bookkeeping people
otherwise rewrite in every project:

* `DataParallel` pieces re-exported in one place: `DistributedOptimizer`, `GraphedStep`, `broadcast_parameters`,
  `broadcast_optimizer_state`, `SyncBatchNorm`;
* `groups`: process sets for the usual rank grids — ranks of my host (`local`), same local rank across hosts (`cross`),
  and rows / columns of a 2-D (data x model) mesh;
* `sharded`: reduce-scatter / all-gather helpers for ZeRO-1 style sharded optimizer state and sequence / expert style
  exchanges (`shard_range`, `reduce_scatter_flat`, `all_gather_flat`, `alltoall_rows`).
"""
from horovod_b200.parallel import groups, sharded  # noqa: F401
from horovod_b200.parallel.groups import Mesh2D, cross_process_set, local_process_set, mesh_2d  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not pull torch-dependent modules unless they are used
    if name in ('DistributedOptimizer', 'GraphedStep', 'broadcast_parameters', 'broadcast_optimizer_state', 'SyncBatchNorm'):
        import horovod_b200.torch as hvd
        return getattr(hvd, name)
    raise AttributeError(name)
