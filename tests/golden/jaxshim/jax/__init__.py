"""`jax` stand-in (see `_core.py`): TEST INFRASTRUCTURE for generating golden
vectors from the unmodified reference sources.  Not JAX."""
import importlib.abc
import importlib.machinery
import sys
import types

from . import _core
from ._core import (Array, checkpoint, config, custom_jvp, debug, grad, hessian, jacfwd, jacobian,  # noqa: F401
                    jacrev, jit, jvp, value_and_grad, vmap, ShapeDtypeStruct, pure_callback)
from . import numpy, lax, tree_util, flatten_util, experimental  # noqa: F401,E402

remat = checkpoint
__version__ = "0.0-shim"


class _Missing:
    def __init__(self, name):
        self._n = name

    def __call__(self, *a, **k):
        # tolerated as a decorator factory at import time; anything else is an error at use
        raise NotImplementedError(f"jax shim: {self._n} is not implemented")

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Missing(f"{self._n}.{k}")

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, o):
        return self

    __ror__ = __or__

    def __getitem__(self, k):
        return self


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing(f"jax.{name}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any `jax.<something>` submodule that the shim does not define becomes an
    empty module whose attributes are `_Missing` placeholders, so that reference
    modules which merely import (but, on the paths exercised, never call) other
    JAX features still import."""

    def find_spec(self, fullname, path, target=None):
        if fullname.startswith("jax.") and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda k, _n=spec.name: (_ for _ in ()).throw(AttributeError(k)) \
            if k.startswith("__") else _Missing(f"{_n}.{k}")
        return m

    def exec_module(self, module):
        pass


sys.meta_path.append(_Finder())
