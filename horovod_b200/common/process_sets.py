"""Process sets: collectives restricted to a subgroup of ranks.

Role parity: horovod/common/process_sets.py (ProcessSet, global_process_set,
add_process_set, remove_process_set).
"""
from typing import List, Optional, Sequence, Union


class ProcessSet(object):
    """Representation of a set of Horovod processes that will run collectives together.

    Initialize with a list of global ranks; hand to hvd.init(process_sets=[...]) or hvd.add_process_set()."""
    process_set_id = None
    ranks = None

    def __init__(self, ranks_or_comm: Sequence[int]):
        if not isinstance(ranks_or_comm, (list, tuple, range)):
            raise ValueError("ProcessSet takes a list of ranks (mpi4py communicators are not supported in this build)")
        self.ranks = sorted(set(int(r) for r in ranks_or_comm))
        self.process_set_id = None

    def _attach(self, process_set_id):
        self.process_set_id = process_set_id

    def _invalidate(self):
        self.process_set_id = None

    def _basics(self):
        from horovod_b200.torch.mpi_ops import _basics
        return _basics

    def size(self) -> Optional[int]:
        if self.process_set_id is None:
            return None
        return self._basics()._process_set_size(self.process_set_id)

    def rank(self) -> Optional[int]:
        """Rank relative to this set, -1 if the calling process is not a member."""
        if self.process_set_id is None:
            return None
        return self._basics()._process_set_rank(self.process_set_id)

    def included(self) -> Optional[bool]:
        if self.process_set_id is None:
            return None
        return self._basics()._process_set_included(self.process_set_id)

    def __str__(self):
        return f"ProcessSet(process_set_id={self.process_set_id}, ranks={self.ranks})"

    def __eq__(self, other):
        return isinstance(other, ProcessSet) and self.process_set_id == other.process_set_id and self.ranks == other.ranks

    def __hash__(self):
        return hash((self.process_set_id, tuple(self.ranks or ())))


global_process_set = ProcessSet([])
global_process_set.process_set_id = 0

_id_to_process_sets = {0: global_process_set}


def _setup(basics):
    """Called after hvd.init(): sync ids / ranks of statically registered sets."""
    global _id_to_process_sets
    table = basics._get_process_set_ids_and_ranks()
    global_process_set.ranks = table.get(0, [])
    global_process_set.process_set_id = 0
    fresh = {0: global_process_set}
    for ps_id, ranks in table.items():
        if ps_id == 0:
            continue
        existing = _id_to_process_sets.get(ps_id)
        if existing is not None and existing.ranks == ranks and existing.process_set_id == ps_id:
            fresh[ps_id] = existing
        else:
            ps = ProcessSet(ranks)
            ps._attach(ps_id)
            fresh[ps_id] = ps
    for ps_id, ps in _id_to_process_sets.items():
        if ps_id not in fresh and ps is not global_process_set:
            ps._invalidate()
    _id_to_process_sets = fresh


def is_process_set_included(process_set_id: int) -> bool:
    from horovod_b200.torch.mpi_ops import _basics
    return _basics._process_set_included(process_set_id)


def _reset_unnamed_op_numbering(ps_id):
    """Ids are recycled; unnamed ops of a new set must be numbered from zero on every member."""
    try:
        from horovod_b200.torch.mpi_ops import _native
        _native().reset_noname_counters(int(ps_id))
    except Exception:  # binding not loaded (front end without torch ops in use)
        pass


def add_process_set(process_set: Union[ProcessSet, Sequence[int]]) -> ProcessSet:
    """Collective: every rank must call it with the same ranks. Returns the registered ProcessSet."""
    from horovod_b200.torch.mpi_ops import _basics
    if not isinstance(process_set, ProcessSet):
        process_set = ProcessSet(process_set)
    if process_set.process_set_id is not None:
        raise ValueError("Attempted to register an already registered process set: " + str(process_set))
    ps_id = _basics._add_process_set_impl(process_set.ranks)
    _reset_unnamed_op_numbering(ps_id)
    process_set._attach(ps_id)
    _id_to_process_sets[ps_id] = process_set
    return process_set


def remove_process_set(process_set: ProcessSet) -> bool:
    """Collective: deregisters a previously added process set."""
    from horovod_b200.torch.mpi_ops import _basics
    ps_id = process_set.process_set_id
    if ps_id is None or ps_id == 0:
        return False
    _basics._remove_process_set_impl(ps_id)
    _reset_unnamed_op_numbering(ps_id)
    _id_to_process_sets.pop(ps_id, None)
    process_set._invalidate()
    return True


def process_set_ids() -> List[int]:
    return sorted(_id_to_process_sets.keys())


def get_process_set_ids_and_ranks():
    from horovod_b200.torch.mpi_ops import _basics
    return _basics._get_process_set_ids_and_ranks()
