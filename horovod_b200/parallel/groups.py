"""Process sets for common rank grids.  Every function here is COLLECTIVE (all ranks call it with the same arguments)
because registering a process set is (`hvd.add_process_set`); sets are created for every group — also the ones this rank
is not part of — so that the registration order is identical everywhere."""
from collections import OrderedDict


def _hvd():
    import horovod_b200.torch as hvd
    return hvd


def _gather_layout():
    """[(local_rank, cross_rank)] indexed by global rank."""
    import torch
    hvd = _hvd()
    mine = torch.tensor([[hvd.local_rank(), hvd.cross_rank()]], dtype=torch.int64)
    table = hvd.allgather(mine, name='parallel.groups.layout')
    return [tuple(int(v) for v in row) for row in table.tolist()]


def _existing(ranks):
    """An already registered set with exactly these ranks (registration is collective, so every rank finds the same)."""
    from horovod_b200.common import process_sets as ps_mod
    hvd = _hvd()
    if sorted(ranks) == list(range(hvd.size())):
        return hvd.global_process_set
    for ps in ps_mod._id_to_process_sets.values():
        if ps.process_set_id not in (None, 0) and sorted(ps.ranks) == sorted(ranks):
            return ps
    return None


def _register(groups):
    """groups: list of rank lists.  Returns (sets created here, the set containing this rank or None); an identical set
    that is already registered is reused (and not owned, i.e. not removed by `Mesh2D.release`)."""
    hvd = _hvd()
    created, mine = [], None
    for ranks in groups:
        ps = _existing(ranks)
        if ps is None:
            ps = hvd.add_process_set(sorted(ranks))
            created.append(ps)
        if hvd.rank() in ranks:
            mine = ps
    return created, mine


def local_process_set():
    """The ranks that share my host (the intra-node NVLink domain).  Returns my set."""
    by_host = OrderedDict()
    for r, (_, cross) in enumerate(_gather_layout()):
        by_host.setdefault(cross, []).append(r)
    if len(by_host) == 1:
        return _hvd().global_process_set
    return _register(list(by_host.values()))[1]


def cross_process_set():
    """The ranks with my local rank on every host (one GPU per node: the inter-node domain).  Returns my set."""
    by_local = OrderedDict()
    for r, (local, _) in enumerate(_gather_layout()):
        by_local.setdefault(local, []).append(r)
    if len(by_local) == 1:
        return _hvd().global_process_set
    return _register(list(by_local.values()))[1]


class Mesh2D:
    """`rows x cols` grid over the global ranks in row-major order: rank = row * cols + col.

    `.row_set` — the ranks of my row (vary the column: e.g. the model-parallel group when cols = model-parallel degree),
    `.col_set` — the ranks of my column (e.g. the data-parallel replicas of my model shard)."""

    def __init__(self, rows, cols, row_set, col_set, row, col, all_sets):
        self.rows, self.cols, self.row_set, self.col_set, self.row, self.col = rows, cols, row_set, col_set, row, col
        self._all = all_sets

    def release(self):
        """Collective: deregisters every set of the mesh."""
        hvd = _hvd()
        for ps in self._all:
            hvd.remove_process_set(ps)
        self._all = []


def mesh_2d(rows, cols):
    """Collective: registers the row and column process sets of a rows x cols grid over all ranks; returns a Mesh2D."""
    hvd = _hvd()
    if rows * cols != hvd.size():
        raise ValueError('mesh %dx%d does not cover %d ranks' % (rows, cols, hvd.size()))
    row_groups = [[r * cols + c for c in range(cols)] for r in range(rows)]
    col_groups = [[r * cols + c for r in range(rows)] for c in range(cols)]
    if cols == 1 or rows == 1:
        # one of the two directions is trivial: size-1 sets are still registered so that every rank's view is uniform
        pass
    row_sets, my_row = _register(row_groups)
    col_sets, my_col = _register(col_groups)
    return Mesh2D(rows, cols, my_row, my_col, hvd.rank() // cols, hvd.rank() % cols, row_sets + col_sets)
