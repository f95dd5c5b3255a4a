"""Import the reference's pure-numpy simulators with jax / numpyro / funsor stubbed.

The reference (/root/reference, read-only, never copied) needs jax+numpyro+funsor to
*fit* models, but its data generators ``simulate``, ``simulate_rn``, ``simulate_cop``
(biolith/models/occu.py:245-430, occu_rn.py:225-358, occu_cop.py:258-396) are plain numpy.
Stubbing the missing third-party packages in ``sys.modules`` lets them run unchanged in this
container; this is only used by make_golden.py to *generate* fixtures -- nothing at test time
or on the GPU box reads /root/reference.
"""

import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

_STUB_ROOTS = ("jax", "numpyro", "funsor", "rpy2", "optax", "flax")


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference_models(reference_root="/root/reference"):
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import biolith.models as models  # noqa: E402

    return models


def remove_stubs():
    """Drop the stub modules again (scipy's array-API helpers probe sys.modules['jax'])."""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _StubFinder)]
    for name in list(sys.modules):
        if name.split(".")[0] in _STUB_ROOTS and isinstance(sys.modules[name], _StubModule):
            del sys.modules[name]
