"""Drop-in namespace for programs written against horovod/horovod.

    import horovod.torch as hvd            # -> horovod_b200.torch
    import horovod.tensorflow.keras as hvd # -> horovod_b200.tensorflow.keras
    from horovod.runner.common.util import hosts
    horovod.run(fn, np=4)

Every `horovod.<x>` import is answered with the module object of `horovod_b200.<x>` (no copies: `horovod.torch is
horovod_b200.torch`), so a training script of the reference runs unchanged.  Nothing here
re-implements anything; if this directory is not on `sys.path` the framework is unaffected.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import horovod_b200 as _impl

_PREFIX, _TARGET = 'horovod.', 'horovod_b200.'


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """`horovod.a.b` -> the already-importable `horovod_b200.a.b`."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            found = importlib.util.find_spec(real)
        except (ImportError, ValueError, AttributeError):
            return None
        if found is None:
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=found.submodule_search_locations is not None)

    def create_module(self, spec):
        return importlib.import_module(_TARGET + spec.name[len(_PREFIX):])

    def exec_module(self, module):
        pass            # the real module is already initialised


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

__version__ = _impl.__version__
run = _impl.run
