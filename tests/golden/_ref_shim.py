"""Import shim for the UNMODIFIED reference (/root/reference) in the build container.

Only used by tests/golden/make_golden.py (fixture generation) -- never at test/bench time:
/root/reference does not exist on the GPU box.  Four third-party packages that the reference
imports at module top level but never executes on the dense-sample thermal-nerfacto path are
replaced by attribute-returning stubs.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("TN_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = MagicMock()
        setattr(self, k, m)
        return m


def install():
    sys.dont_write_bytecode = True
    for n in ["viser", "viser.transforms", "nerfacc", "matplotlib",
              "torchmetrics", "torchmetrics.functional", "torchmetrics.image", "torchmetrics.image.lpip"]:
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            sys.modules[n] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
