"""Import shim: the package directory is `amrvw.jl_b200/` (a dot is not importable), so load it
under the module name `amrvw_jl_b200`."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "amrvw.jl_b200")
_spec = importlib.util.spec_from_file_location("amrvw_jl_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["amrvw_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
