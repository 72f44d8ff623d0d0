"""Import shim: loads the package directory ``directtrajopt.jl_b200/`` (whose name is not a valid
Python identifier) under the module name ``dto_b200``."""
import importlib.util as _u
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "directtrajopt.jl_b200")
_spec = _u.spec_from_file_location("dto_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["dto_b200"] = _mod
_spec.loader.exec_module(_mod)
