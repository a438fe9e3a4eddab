"""Import shim: `import ahmc_b200` loads the package that lives in the directory `advancedhmc.jl_b200/`
(whose name, mandated by the repo layout, is not a legal Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "advancedhmc.jl_b200")
_spec = importlib.util.spec_from_file_location("ahmc_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ahmc_b200"] = _mod
_spec.loader.exec_module(_mod)
