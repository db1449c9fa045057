"""TorchState: elastic state for models, optimizers, samplers and plain values.

Public API parity with horovod/torch/elastic/state.py (`TorchState(model=..., optimizer=..., **values)`, the handler
registry, `StateHandler` subclasses).  What is different:

* a commit of a model / optimizer does not deep-copy tensors on the device (which doubles their HBM footprint and is what
  the reference's `copy.deepcopy(state_dict())` does): tensors are snapshotted into **pinned host buffers** that are
  allocated once and refilled with asynchronous D2H copies on every commit (`_HostSnapshot`); restore copies them back;
* one generic state-dict handler (`_StateDictHandler`) serves modules and optimizers; a handler is described by how to
  read / load its state dict and how to synchronise it across ranks.
"""
import copy

import torch

from horovod_b200.common.elastic import ObjectState
from horovod_b200.torch.elastic.sampler import ElasticSampler
from horovod_b200.torch.functions import (allgather_object, broadcast_object, broadcast_optimizer_state,
                                          broadcast_parameters)
from horovod_b200.torch.mpi_ops import rank


class _HostSnapshot:
    """A nested structure (dicts / lists / tuples / tensors / plain values) mirrored on the host.  Tensor leaves keep a
    reusable (pinned, when CUDA is present) host buffer; everything else is deep-copied."""

    def __init__(self):
        self._buffers = {}
        self._tree = None

    def _host_copy(self, path, t):
        buf = self._buffers.get(path)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=t.is_cuda and torch.cuda.is_available())
            self._buffers[path] = buf
        buf.copy_(t.detach(), non_blocking=t.is_cuda)
        return buf

    def _walk(self, obj, path, leaf):
        if isinstance(obj, torch.Tensor):
            return leaf(path, obj)
        if isinstance(obj, dict):
            return {k: self._walk(v, path + (('k', k),), leaf) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._walk(v, path + (('i', i),), leaf) for i, v in enumerate(obj))
        return copy.deepcopy(obj)

    def capture(self, tree):
        self._tree = self._walk(tree, (), self._host_copy)
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()  # the snapshot must be complete when commit() returns

    def materialise(self):
        """A copy of the snapshot whose tensors are fresh clones (load_state_dict copies them to the right device)."""
        return self._walk(self._tree, (), lambda path, t: t.clone())


class StateHandler:
    """Tracks one object of the training state.  `value` is the live object."""

    def __init__(self, value):
        self.value = value

    def save(self):
        raise NotImplementedError

    def restore(self):
        raise NotImplementedError

    def sync(self):
        raise NotImplementedError

    def set_value(self, value):
        self.value = value
        self.save()


class _StateDictHandler(StateHandler):
    """Anything with state_dict() / load_state_dict()."""

    def __init__(self, value):
        super().__init__(value)
        self._snapshot = _HostSnapshot()
        self.save()

    def save(self):
        self._snapshot.capture(self.value.state_dict())

    def restore(self):
        self.value.load_state_dict(self._snapshot.materialise())


class ModelStateHandler(_StateDictHandler):
    def sync(self):
        broadcast_parameters(self.value.state_dict(), root_rank=0)


class OptimizerStateHandler(_StateDictHandler):
    def sync(self):
        broadcast_optimizer_state(self.value, root_rank=0)


class SamplerStateHandler(StateHandler):
    """ElasticSampler: what matters is the union of the samples every rank already consumed this epoch."""

    def __init__(self, sampler):
        super().__init__(sampler)
        self._saved = None
        self.save()

    def save(self):
        self._saved = copy.deepcopy(self.value.state_dict())

    def restore(self):
        self.value.load_state_dict(copy.deepcopy(self._saved))

    def sync(self):
        consumed = set().union(*[set(part) for part in allgather_object(self.value.processed_indices)])
        merged = dict(self.value.state_dict(), processed_indices=consumed)
        self.value.load_state_dict(broadcast_object(merged))  # rank 0's epoch / seed view wins


# (type, handler class), first match wins; users can extend it (get_handler_registry / set_handler_registry)
_handler_registry = [
    (torch.nn.Module, ModelStateHandler),
    (torch.optim.Optimizer, OptimizerStateHandler),
    (ElasticSampler, SamplerStateHandler),
]


def get_handler_registry():
    return _handler_registry


def set_handler_registry(registry):
    global _handler_registry
    _handler_registry = registry


def _handler_for(value):
    for kind, cls in _handler_registry:
        if isinstance(value, kind):
            return cls(value)
    return None


class TorchState(ObjectState):
    """`TorchState(model=model, optimizer=opt, sampler=sampler, epoch=0, batch=0)`: objects with a registered handler are
    saved / restored / synchronised through it, everything else is treated as a plain (picklable) value."""

    def __init__(self, model=None, optimizer=None, **kwargs):
        named = dict(kwargs)
        if model is not None:
            named['model'] = model
        if optimizer is not None:
            named['optimizer'] = optimizer
        handlers, plain = {}, {}
        for name, value in named.items():
            if value is None:
                continue
            h = _handler_for(value)
            if h is not None:
                handlers[name] = h
            else:
                plain[name] = value
        object.__setattr__(self, '_handlers', handlers)
        for name, h in handlers.items():
            object.__setattr__(self, name, h.value)
        super().__init__(bcast_object=broadcast_object, get_rank=rank, **plain)

    def save(self):
        for h in self._handlers.values():
            h.save()
        super().save()

    def restore(self):
        for h in self._handlers.values():
            h.restore()
        super().restore()

    def sync(self):
        for h in self._handlers.values():
            h.sync()
        super().sync()

    def __setattr__(self, name, value):
        handlers = self.__dict__.get('_handlers', {})
        if name in handlers:
            handlers[name].set_value(value)  # e.g. state.model = new_model re-points the handler and snapshots it
        object.__setattr__(self, name, value)
