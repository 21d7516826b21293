"""Import shim: the package directory is named `divide-and-conquer-smc_b200/` (not a valid
Python identifier), so `import dcsmc_b200` loads it from there."""
import importlib.util as _u
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg = _os.path.join(_here, "divide-and-conquer-smc_b200")
_spec = _u.spec_from_file_location(
    "dcsmc_b200", _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg]
)
_mod = _u.module_from_spec(_spec)
_sys.modules["dcsmc_b200"] = _mod
_spec.loader.exec_module(_mod)
