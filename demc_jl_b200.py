"""Import shim: the package directory is named `demc.jl_b200/` after the reference
(chrished/DEMC.jl); a dotted directory cannot be imported by name, so this module loads it and
registers it as `demc_jl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demc.jl_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
