"""Importable alias of the product package, whose directory is named ``morphik-core_b200`` (not a valid
Python identifier).  ``import morphik_core_b200`` executes ``morphik-core_b200/__init__.py`` in this module's
namespace and points ``__path__`` at that directory, so ``morphik_core_b200.store`` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "morphik-core_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
