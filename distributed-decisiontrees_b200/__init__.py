"""B200-native decision-tree-ensemble inference engine (drop-in for the Core tree walk +
ResultsCombiner aggregation path of fpgasystems/Distributed-DecisionTrees).

The directory name carries a hyphen, so import it with
``importlib.import_module("distributed-decisiontrees_b200")`` or through the ``ddt_b200`` shim at
the repository root.  The product is ``libdte.so`` (CUDA, C ABI in ``include/dte.h``); this package
is only the thin host-side mirror of the reference's register/stream interface.
"""
from .engine import Engine, DteError, lib_path, build_library, load_library  # noqa: F401
from . import layout  # noqa: F401
from . import sharding  # noqa: F401

__all__ = ["Engine", "DteError", "layout", "sharding", "lib_path", "build_library", "load_library"]
